"""Plugin-path end-to-end timing (SHEngine.synthesis + .analysis, pageable numpy in) for the current SHB200_HOST_THREADS /
SHB200_BOUNCE_MIB, plus the cost of cudaHostRegister on the same array (the alternative to the bounce ring).
    SHB200_HOST_THREADS=8 python tools/e2e_sweep.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import shxarray_b200 as sb  # noqa: E402
from shxarray_b200 import synthetic as S  # noqa: E402

NMAX, B = 720, 64
g = sb.lonlat_grid(NMAX, gtype="gauss")
lon, lat = g["lon"], g["lat"]
c = S.random_coefficients(NMAX, B, kaula=True)
eng = sb.SHEngine(0)


def t(fn, reps=6):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


res = {"threads": os.environ.get("SHB200_HOST_THREADS", "default"), "bounce_mib": os.environ.get("SHB200_BOUNCE_MIB", "default")}
res["synthesis_ms"] = t(lambda: eng.synthesis(c, lon=lon, lat=lat, nmax=NMAX))
gpin = eng.synthesis(c, lon=lon, lat=lat, nmax=NMAX)
res["analysis_pinned_in_ms"] = t(lambda: eng.analysis(gpin, NMAX, lon, lat, quadrature="gauss"))
gpage = np.array(gpin)
res["analysis_pageable_in_ms"] = t(lambda: eng.analysis(gpage, NMAX, lon, lat, quadrature="gauss"))
rt = torch.cuda.cudart()
t0 = time.perf_counter()
rc = rt.cudaHostRegister(c.ctypes.data, c.nbytes, 0)
t1 = time.perf_counter()
rt.cudaHostUnregister(c.ctypes.data)
t2 = time.perf_counter()
res["cudaHostRegister_266MB_ms"] = (t1 - t0) * 1e3
res["cudaHostUnregister_ms"] = (t2 - t1) * 1e3
res["register_rc"] = str(rc)
# plain memcpy rate of one thread (numpy copy) for reference
dst = np.empty_like(c)
np.copyto(dst, c)
t0 = time.perf_counter()
for _ in range(3):
    np.copyto(dst, c)
res["numpy_copy_gbs_1thread"] = 3 * c.nbytes / (time.perf_counter() - t0) * 1e-9
print(json.dumps(res))
