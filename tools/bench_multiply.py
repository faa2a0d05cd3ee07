#!/usr/bin/env python
"""Device-resident spectral product (BASELINE cfg5 pattern, shtns.py:394-420): two syntheses on the Gauss grid of
nmax_a + nmax_b, pointwise product, Gauss-quadrature analysis -- one shb200_multiply call per batch."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import shxarray_b200 as sb
from shxarray_b200 import synthetic as S

eng = sb.SHEngine(0)
out = []
for na, nb, B in ((128, 128, 4), (128, 128, 1), (256, 256, 4), (360, 360, 16)):
    a = torch.from_numpy(S.random_coefficients(na, B, kaula=True)).cuda()
    b = torch.from_numpy(S.random_coefficients(nb, B, kaula=True, seed=5)).cuda()
    p = eng.multiply(a, na, b, nb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        eng.multiply(a, na, b, nb, out=p)          # same buffers every call: the sea-level loop's pattern (graph replay when small)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    g = sb.lonlat_grid(na + nb, gtype="gauss")
    res = dict(nmax_a=na, nmax_b=nb, batch=B, grid=[len(g["lat"]), len(g["lon"])], ms_per_product=ms, products_per_s=B / (ms * 1e-3))
    print(json.dumps(res), flush=True)
    out.append(res)
    eng.clear()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/multiply_r02.json", "w"), indent=1)
