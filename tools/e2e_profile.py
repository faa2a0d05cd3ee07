"""Where the plugin-path round trip (SHEngine.synthesis + .analysis with pageable numpy coefficients, as bench.py's e2e) spends
its host time: per-call wall times and a cProfile of ten steps.   python tools/e2e_profile.py"""
import cProfile, io, json, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import shxarray_b200 as sb
from shxarray_b200 import synthetic as S

NMAX, B = 720, 64
g = sb.lonlat_grid(NMAX, gtype="gauss")
lon, lat = g["lon"], g["lat"]
c = S.random_coefficients(NMAX, B, kaula=True)
eng = sb.SHEngine(0)
ts, ta = [], []


def step():
    t0 = time.perf_counter()
    g_ = eng.synthesis(c, lon=lon, lat=lat, nmax=NMAX)
    t1 = time.perf_counter()
    a = eng.analysis(g_, NMAX, lon, lat, quadrature="gauss")
    t2 = time.perf_counter()
    ts.append((t1 - t0) * 1e3); ta.append((t2 - t1) * 1e3)
    return a


for _ in range(3):
    res = step()
ts.clear(); ta.clear()
t0 = time.perf_counter()
for _ in range(10):
    res = step()
tot = (time.perf_counter() - t0) / 10 * 1e3
print(json.dumps({"step_ms": tot, "synthesis_ms": [round(x, 2) for x in ts], "analysis_ms": [round(x, 2) for x in ta]}))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    res = step()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14)
print(s.getvalue()[:3500])
