"""Wall-clock and enqueue cost per engine call for the launch-latency-bound configs (cfg1, cfg5): host side vs GPU side."""
import os, sys, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import shxarray_b200 as sb
from shxarray_b200 import synthetic as S
eng = sb.SHEngine(0)
for nmax, B, gt, q in ((60, 1, "regular", "shlib"), (256, 4, "gauss", "gauss")):
    g = sb.lonlat_grid(nmax, gtype=gt)
    c = torch.from_numpy(S.random_coefficients(nmax, B, kaula=True)).cuda()
    f = eng.synthesis(c, lon=g["lon"], lat=g["lat"], nmax=nmax)
    a = eng.analysis(f, nmax, g["lon"], g["lat"], quadrature=q)
    torch.cuda.synchronize()
    for name, fn in (("syn", lambda: eng.synthesis(c, lon=g["lon"], lat=g["lat"], nmax=nmax, out=f)), ("ana", lambda: eng.analysis(f, nmax, g["lon"], g["lat"], quadrature=q, out=a))):
        t0 = time.perf_counter()
        for _ in range(2000): fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 2000
        # host-side cost alone: time to enqueue
        t0 = time.perf_counter()
        for _ in range(200): fn()
        t1 = time.perf_counter() - t0
        torch.cuda.synchronize()
        print(nmax, B, name, "wall per call %.1f us" % (dt * 1e6), "enqueue %.1f us" % (t1 / 200 * 1e6))
