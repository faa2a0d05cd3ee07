"""cfg3 with batch-first grids against shlib's batch-fastest [nlat, nlon, B] layout (device-resident, CUDA events)."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import shxarray_b200 as sb
from shxarray_b200 import synthetic as S
eng = sb.SHEngine(0)
nmax, B = 720, 64
g = sb.lonlat_grid(nmax, gtype="gauss")
c = torch.from_numpy(S.random_coefficients(nmax, B, kaula=True)).cuda()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for layout in ("batch_first", "shlib"):
    f = eng.synthesis(c, lon=g["lon"], lat=g["lat"], nmax=nmax, layout=layout)
    ts = t(lambda: eng.synthesis(c, lon=g["lon"], lat=g["lat"], nmax=nmax, layout=layout, out=f))
    a = eng.analysis(f, nmax, g["lon"], g["lat"], quadrature="gauss", layout=layout)
    ta = t(lambda: eng.analysis(f, nmax, g["lon"], g["lat"], quadrature="gauss", layout=layout, out=a))
    print(layout, tuple(f.shape), "synthesis %.3f ms analysis %.3f ms" % (ts, ta), "err", float(((a - c).abs().amax(dim=1) / c.abs().amax(dim=1)).max()))
