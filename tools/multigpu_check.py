#!/usr/bin/env python
"""Multi-GPU check on real hardware (torchrun, NCCL over NVLink): latitude-sharded synthesis and m-order-sharded analysis
of ONE nmax=2160 field (BASELINE config 4) against the single-GPU transform on rank 0; device timings, max over ranks.
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multigpu_check.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import shxarray_b200 as sb
from shxarray_b200 import synthetic as S
from shxarray_b200.distributed import ShardedSH

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

nmax = int(os.environ.get("NMAX", "2160"))
d = 180.0 / nmax
h = np.arange(d / 2, 90, d)
lat = np.concatenate([-h[::-1], h])
lon = np.arange(0, 360, d)
w = (d * np.pi / 180) ** 2 / (4 * np.pi)
lw = np.cos(lat * (np.pi / 180))
eng = sb.SHEngine(device=local)
sh = ShardedSH(eng)
c = torch.from_numpy(S.random_coefficients(nmax, 1, kaula=True)).cuda()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, float(t.item())


f_sh, t_syn = timed(lambda: sh.synthesis_latsharded(c, lon, lat, nmax=nmax))
a_sh, t_ana = timed(lambda: sh.analysis_msharded(f_sh, nmax, lon, lat, quadrature="weights", lat_weights=lw, weight=w))
res = {"world": world, "nmax": nmax, "grid": [len(lat), len(lon)], "synthesis_latsharded_ms": t_syn, "analysis_msharded_ms": t_ana}
if rank == 0:
    f1, t1s = None, None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f1 = eng.synthesis(c, lon=lon, lat=lat, nmax=nmax)
    a1 = eng.analysis(f1, nmax, lon, lat, quadrature="weights", lat_weights=lw, weight=w)
    torch.cuda.synchronize()
    e0.record(); f1 = eng.synthesis(c, lon=lon, lat=lat, nmax=nmax); e1.record(); torch.cuda.synchronize()
    res["synthesis_1gpu_ms"] = e0.elapsed_time(e1)
    e0.record(); a1 = eng.analysis(f1, nmax, lon, lat, quadrature="weights", lat_weights=lw, weight=w); e1.record(); torch.cuda.synchronize()
    res["analysis_1gpu_ms"] = e0.elapsed_time(e1)
    res["synthesis_max_rel_diff_vs_1gpu"] = float((f_sh - f1).abs().max() / f1.abs().max())
    res["analysis_bit_identical_to_1gpu"] = bool(torch.equal(a_sh, a1))
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()
