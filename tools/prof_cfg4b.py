#!/usr/bin/env python
"""One synthesis + analysis of 16 fields at nmax 2160 on the 5' grid (20 GB P table, 32-column tile shapes of the table kernels)
for an ncu capture:  ncu --set full --clock-control none -k regex:'k_leg_syn_tab|k_leg_ana_tab' -c 2 -o gpurun_out/prof_cfg4b python tools/prof_cfg4b.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import shxarray_b200 as sb
from shxarray_b200 import synthetic as S
eng = sb.SHEngine(0)
nmax, d = 2160, 1.0 / 12
h = np.arange(d / 2, 90, d)
lat = np.concatenate([-h[::-1], h]); lon = np.arange(0, 360, d)
c = torch.from_numpy(S.random_coefficients(nmax, 16, kaula=True)).cuda()
f = eng.synthesis(c, lon=lon, lat=lat, nmax=nmax)
a = eng.analysis(f, nmax, lon, lat, quadrature="weights", lat_weights=np.cos(lat * (np.pi / 180)), weight=(d * np.pi / 180) ** 2 / (4 * np.pi))
torch.cuda.synchronize()
print("ok", [p.last_kernels() for p in eng._plans.values()])
