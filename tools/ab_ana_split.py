#!/usr/bin/env python
"""A/B of the row-tile split of k_leg_ana_small (SHB200_ANA_SPLIT = 0 never / 1 always / unset: when the orders underfill the GPU) on ONE GPU: analysis of one nmax-2160
field on the 5' grid, unsharded and as the order shards (r, 8) an 8-GPU run gives each rank; device time per call and a
checksum of the result bits (the split must not change a bit).
    SHB200_ANA_SPLIT=0 python tools/ab_ana_split.py ; SHB200_ANA_SPLIT=1 python tools/ab_ana_split.py"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import shxarray_b200 as sb
from shxarray_b200 import capi, synthetic as S

nmax = int(os.environ.get("NMAX", "2160"))
d = 180.0 / nmax
h = np.arange(d / 2, 90, d)
lat = np.concatenate([-h[::-1], h])
lon = np.arange(0, 360, d)
w = (d * np.pi / 180) ** 2 / (4 * np.pi)
lw = np.cos(lat * (np.pi / 180))
eng = sb.SHEngine(device=0, flags=capi.FLAG_TIMING)
c = torch.from_numpy(S.random_coefficients(nmax, 1, kaula=True)).cuda()
f = eng.synthesis(c, lon=lon, lat=lat, nmax=nmax)
kw = dict(quadrature="weights", lat_weights=lw, weight=w)


def timed(fn, reps=5):
    out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / reps


res = {"SHB200_ANA_SPLIT": os.environ.get("SHB200_ANA_SPLIT", "unset"), "nmax": nmax}
cases = (("unsharded", None), ("shard_0_of_8", (0, 8)), ("shard_7_of_8", (7, 8)), ("shard_1_of_4", (1, 4)), ("shard_1_of_2", (1, 2)))
only = os.environ.get("CASES")
for tag, shard in cases:
    if only and tag not in only.split(","):
        continue
    a, t = timed(lambda: eng.analysis(f, nmax, lon, lat, order_shard=shard, **kw))
    plan = list(eng._plans.values())[-1]
    res[tag] = {"ms": t, "sha": hashlib.sha256(a.cpu().numpy().tobytes()).hexdigest()[:16], "stages_ms": {k: round(v, 4) for k, v in plan.stage_times().items() if v > 0 and k.startswith("ana")}, "kernels": plan.last_kernels()}
print(json.dumps(res))
