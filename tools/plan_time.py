"""Plan creation time (host planner tables + device P table) for the BASELINE grids."""
import os, sys, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import shxarray_b200 as sb
from shxarray_b200 import capi
torch.cuda.init()
for nmax, B, gt in ((60, 1, "regular"), (96, 240, "regular"), (256, 4, "gauss"), (720, 64, "gauss"), (2160, 1, None), (2160, 16, None)):
    if gt:
        g = sb.lonlat_grid(nmax, gtype=gt); lon, lat = g["lon"], g["lat"]
    else:
        d = 1.0 / 12; h = np.arange(d / 2, 90, d); lat = np.concatenate([-h[::-1], h]); lon = np.arange(0, 360, d)
    t0 = time.perf_counter()
    p = capi.Plan(nmax, lat, lon, max_batch=B, quadrature=capi.QUAD_LATWEIGHTS, weight=1.0, lat_weights=np.cos(np.deg2rad(lat)))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(nmax, B, "plan creation %.1f ms" % (dt * 1e3), "workspace %.2f GB" % (p.info.workspace_bytes / 1e9), "dmma", p.info.dmma)
    p.close()
