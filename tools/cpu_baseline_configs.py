#!/usr/bin/env python
"""shlib arithmetic (oracle/_ref = reference headers + scipy OpenBLAS, else the C port) timed on the host cores for the
BASELINE.json configs: 1, 2, 5 in full, 3 and 4 on a latitude subset extrapolated linearly in grid points (SURVEY.md §8d:
every point costs the same in synthesis.pyx:179-183 / analysis.pyx:128-139).  Writes gpurun_out/cpu_configs_r01.json."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from shxarray_b200 import synthetic as S

kind = O.best_kind()
nthr = len(os.sched_getaffinity(0))
out = []


def run(tag, nmax, B, lon, lat, rows=None, ncols_s=None, ncols_a=None):
    n, m = O.nm_index(nmax)
    c = S.random_coefficients(nmax, B, kaula=True)
    L, K = len(lat), len(lon)
    ri = np.arange(L) if rows is None else rows
    cs = np.arange(K) if ncols_s is None else np.linspace(0, K - 1, ncols_s).round().astype(int)
    ca = np.arange(K) if ncols_a is None else np.linspace(0, K - 1, ncols_a).round().astype(int)
    t0 = time.perf_counter()
    O.synthesis(c, n, m, lon[cs], lat[ri], grid=True, kind=kind, nthreads=nthr)
    t1 = time.perf_counter()
    f = np.random.default_rng(0).standard_normal((B, len(ri), len(ca)))
    O.analysis(f, nmax, lon[ca], lat[ri], kind=kind, nthreads=nthr, weight=1e-6)
    t2 = time.perf_counter()
    ts = (t1 - t0) * (L * K) / (len(ri) * len(cs))
    ta = (t2 - t1) * (L * K) / (len(ri) * len(ca))
    r = dict(config=tag, nmax=nmax, batch=B, grid=[L, K], cores=nthr, kind=kind, synthesis_s=ts, analysis_s=ta,
             fields_per_s_roundtrip=B / (ts + ta),
             sample="full" if rows is None else f"{len(ri)} of {L} rows x {len(cs)}/{len(ca)} of {K} longitudes, extrapolated")
    print(json.dumps(r), flush=True)
    out.append(r)


lon, lat = O.shlib_grid(60)
run("cfg1 nmax60 shlib 1deg B=1", 60, 1, lon, lat)
lon, lat = O.shlib_grid(96)
run("cfg2 nmax96 0.5deg B=240", 96, 240, lon, lat, rows=np.linspace(0, len(lat) - 1, max(nthr, 8)).round().astype(int), ncols_s=48, ncols_a=8)
lon, lat, _ = S.gauss_grid(260, 576)
run("cfg5 nmax256 gauss B=4", 256, 4, lon, lat, rows=np.linspace(0, 259, max(nthr, 8)).round().astype(int), ncols_s=96, ncols_a=48)
lon, lat, _ = S.gauss_grid(724, 1536)
run("cfg3 nmax720 gauss B=64", 720, 64, lon, lat, rows=np.linspace(0, 723, max(nthr, 8)).round().astype(int), ncols_s=16, ncols_a=6)
d = 1.0 / 12
h = np.arange(d / 2, 90, d)
lat = np.concatenate([-h[::-1], h]); lon = np.arange(0, 360, d)
run("cfg4 nmax2160 5min B=1", 2160, 1, lon, lat, rows=np.linspace(0, 2159, max(nthr, 8)).round().astype(int), ncols_s=8, ncols_a=6)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/cpu_configs_r01.json", "w"), indent=1)
