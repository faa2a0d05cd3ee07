#!/usr/bin/env python
"""One synthesis + analysis of the single-field nmax=2160 / 5 arc-minute case (BASELINE cfg4), for ncu captures:
ncu --set full --import-source on --clock-control none -k regex:'k_leg_ana_small|k_leg_syn_dfma' -c 2 -o gpurun_out/cfg4 python tools/prof_cfg4.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import shxarray_b200 as sb
from shxarray_b200 import synthetic as S

nmax, B = 2160, int(sys.argv[1]) if len(sys.argv) > 1 else 1
d = 1.0 / 12
h = np.arange(d / 2, 90, d)
lat = np.concatenate([-h[::-1], h])
lon = np.arange(0, 360, d)
w = (d * np.pi / 180) ** 2 / (4 * np.pi)
eng = sb.SHEngine(0)
c = torch.from_numpy(S.random_coefficients(nmax, B, kaula=True)).cuda()
f = eng.synthesis(c, lon=lon, lat=lat, nmax=nmax)
a = eng.analysis(f, nmax, lon, lat, quadrature="weights", lat_weights=np.cos(lat * (np.pi / 180)), weight=w)
torch.cuda.synchronize()
print("roundtrip", float(((a - c).abs().amax(dim=1) / c.abs().amax(dim=1)).max()))
