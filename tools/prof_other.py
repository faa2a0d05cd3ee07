#!/usr/bin/env python
"""One pass of the kernels that are NOT on the cfg3 step, for one ncu capture (profiles/r02_ncu_other_kernels.txt):
cfg2 (nmax 96, 360x720, 240 fields: the HBM-bound config), the fused point kernel (1e5 points, nmax 200, 20 fields) and a
1-of-8 order shard of the single-field nmax-2160 analysis (split tiny-batch kernel + ordered reduction).
    ncu --set full --clock-control none -k regex:'k_lon_.*fft3|k_syn_points|k_leg_ana_small|k_ana_small_reduce|k_leg_syn_dfma|k_pack_tri' \
        -o gpurun_out/prof_other python tools/prof_other.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import shxarray_b200 as sb
from shxarray_b200 import synthetic as S

eng = sb.SHEngine(0)
# cfg2
g = sb.lonlat_grid(96, gtype="regular")
lon, lat = g["lon"], g["lat"]
step = 180.0 / len(lat)
lat = np.arange(-90 + 0.25, 90, 0.5); lon = np.arange(-180 + 0.25, 180, 0.5)
c = torch.from_numpy(S.random_coefficients(96, 240, kaula=True)).cuda()
f = eng.synthesis(c, lon=lon, lat=lat, nmax=96)
a = eng.analysis(f, 96, lon, lat, quadrature="shlib")
# points
rng = np.random.default_rng(3)
plon, plat = rng.uniform(-180, 180, 100000), rng.uniform(-90, 90, 100000)
cp = torch.from_numpy(S.random_coefficients(200, 20, kaula=True)).cuda()
fp = eng.synthesis(cp, lon=plon, lat=plat, nmax=200, points=True)
# cfg4 order shard
nmax = 2160
d = 1.0 / 12
h = np.arange(d / 2, 90, d)
lat4 = np.concatenate([-h[::-1], h]); lon4 = np.arange(0, 360, d)
w = (d * np.pi / 180) ** 2 / (4 * np.pi)
c4 = torch.from_numpy(S.random_coefficients(nmax, 1, kaula=True)).cuda()
f4 = eng.synthesis(c4, lon=lon4, lat=lat4, nmax=nmax)
a4 = eng.analysis(f4, nmax, lon4, lat4, quadrature="weights", lat_weights=np.cos(lat4 * (np.pi / 180)), weight=w, order_shard=(0, 8))
torch.cuda.synchronize()
print("ok", float(fp.abs().max()), float(a4.abs().max()))
