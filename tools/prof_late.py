#!/usr/bin/env python
"""One pass of the kernels added late in round 2, for one ncu capture (profiles/r02_ncu_late_kernels.txt): the single-field
nmax-2160 analysis (k_leg_ana_one, unsharded and as a 1-of-8 order shard in split mode), rows of 5760 longitudes (three-pass
16 x 18 x 20) and of 11520 longitudes (radix-2 split, k_lon_*_fft3s2) with 16 fields at nmax 720.
    ncu --set full --clock-control none -k regex:'k_leg_ana_one|fft3s2|Cfg<5760' -o gpurun_out/prof_late python tools/prof_late.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import shxarray_b200 as sb
from shxarray_b200 import synthetic as S

os.environ.setdefault("SHB200_GRAPHS", "0")
eng = sb.SHEngine(0)
nmax = 2160
d = 1.0 / 12
h = np.arange(d / 2, 90, d)
lat4 = np.concatenate([-h[::-1], h]); lon4 = np.arange(0, 360, d)
w = (d * np.pi / 180) ** 2 / (4 * np.pi)
kw = dict(quadrature="weights", lat_weights=np.cos(lat4 * (np.pi / 180)), weight=w)
f4 = torch.from_numpy(np.random.default_rng(1).standard_normal((1, len(lat4), len(lon4)))).cuda()
a4 = eng.analysis(f4, nmax, lon4, lat4, **kw)
a4s = eng.analysis(f4, nmax, lon4, lat4, order_shard=(0, 8), **kw)
eng.clear()
for K in (5760, 11520):
    nlat = K // 2
    step = 360.0 / K
    lat = -90 + step / 2 + np.arange(nlat) * step
    lon = -180 + step / 2 + np.arange(K) * step
    c = torch.from_numpy(S.random_coefficients(720, 16, kaula=True)).cuda()
    f = eng.synthesis(c, lon=lon, lat=lat, nmax=720)
    a = eng.analysis(f, 720, lon, lat, quadrature="shlib")
    print(K, [p.last_kernels() for p in eng._plans.values()][-1], float((a - c).abs().max()))
    del f, a
    eng.clear()
torch.cuda.synchronize()
print("ok", float(a4.abs().max()), float(a4s.abs().max()))
