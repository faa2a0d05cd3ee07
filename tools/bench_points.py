#!/usr/bin/env python
"""Point-mode synthesis (synthesis.pyx:186-189; the pattern of isokernelbase.py:100-113: an isotropic kernel evaluated at many
positions) and dense-longitude grids (irregular longitudes: no FFT): device time per call and per-stage times.
    python tools/bench_points.py  -> gpurun_out/points_r02.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import shxarray_b200 as sb
from shxarray_b200 import capi, synthetic as S


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = []
rng = np.random.default_rng(0)
for npts, nmax, B in ((100000, 200, 20), (20000, 200, 20), (10000, 720, 4)):
    eng = sb.SHEngine(0, flags=capi.FLAG_TIMING)
    lon = rng.uniform(-180, 180, npts)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, npts)))
    c = torch.from_numpy(S.random_coefficients(nmax, B, kaula=True)).cuda()
    f = eng.synthesis(c, lon=lon, lat=lat, nmax=nmax, points=True)
    ms = timeit(lambda: eng.synthesis(c, lon=lon, lat=lat, nmax=nmax, points=True, out=f))
    plan = list(eng._plans.values())[-1]
    st = {k: round(v, 4) for k, v in plan.stage_times().items() if v > 0}
    out.append(dict(case="points", npoints=npts, nmax=nmax, batch=B, ms=ms, points_per_s=npts * B / (ms * 1e-3), stages_ms=st,
                    kernels=plan.last_kernels(), workspace_gb=plan.info.workspace_bytes / 1e9))
    print(json.dumps(out[-1]), flush=True)
    eng.clear()
# irregular longitudes: dense trig sums instead of the FFT, synthesis and analysis
for nmax, B in ((96, 24), (200, 8)):
    eng = sb.SHEngine(0, flags=capi.FLAG_TIMING | capi.FLAG_DENSE_LON)
    g = sb.lonlat_grid(nmax, gtype="regular")
    c = torch.from_numpy(S.random_coefficients(nmax, B, kaula=True)).cuda()
    f = eng.synthesis(c, lon=g["lon"], lat=g["lat"], nmax=nmax)
    a = eng.analysis(f, nmax, g["lon"], g["lat"])
    ts = timeit(lambda: eng.synthesis(c, lon=g["lon"], lat=g["lat"], nmax=nmax, out=f))
    ta = timeit(lambda: eng.analysis(f, nmax, g["lon"], g["lat"], out=a))
    st = {}
    for pl in eng._plans.values():
        st.update({k: round(v, 4) for k, v in pl.stage_times().items() if v > 0})
    out.append(dict(case="dense longitude stage", nmax=nmax, batch=B, grid=[len(g["lat"]), len(g["lon"])], synthesis_ms=ts, analysis_ms=ta, stages_ms=st))
    print(json.dumps(out[-1]), flush=True)
    eng.clear()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/points_r02.json", "w"), indent=1)
