"""Determinism stress of the kernels added late in round 2 (compute-sanitizer is closed on this pool): every repeat must give
the bits of the first call -- a race in the named-barrier ring of k_leg_ana_one or in the split long-row kernels would show
up as a difference.  python tools/stress_new_kernels.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import shxarray_b200 as sb
from shxarray_b200 import synthetic as S

os.environ.setdefault("SHB200_GRAPHS", "0")
eng = sb.SHEngine(0)
bad = 0
for nmax, gt, q, reps in ((2160, None, "weights", 150), (300, "regular", "shlib", 300), (95, "gauss", "gauss", 300)):
    if gt is None:
        d = 1.0 / 12
        h = np.arange(d / 2, 90, d)
        lat = np.concatenate([-h[::-1], h]); lon = np.arange(0, 360, d)
        kw = dict(quadrature="weights", lat_weights=np.cos(lat * np.pi / 180), weight=(d * np.pi / 180) ** 2 / (4 * np.pi))
    else:
        g = sb.lonlat_grid(nmax, gtype=gt); lon, lat = g["lon"], g["lat"]; kw = dict(quadrature=q)
    f = torch.from_numpy(np.random.default_rng(nmax).standard_normal((1, len(lat), len(lon)))).cuda()
    first = eng.analysis(f, nmax, lon, lat, **kw).clone()
    kern = [p.last_kernels() for p in eng._plans.values()][-1]
    out = torch.empty_like(first)
    nb = 0
    for _ in range(reps):
        eng.analysis(f, nmax, lon, lat, out=out, **kw)
        nb += int(not torch.equal(out, first))
    print(f"nmax {nmax} {kern} repeats {reps} differing {nb}", flush=True)
    bad += nb
    eng.clear()
# rows of 11520 longitudes, synthesis and analysis
K = 11520
lon = -180.0 + 180.0 / K + np.arange(K) * (360.0 / K)
lat = np.concatenate([-(np.arange(40) + 0.53125)[::-1], np.arange(40) + 0.53125])
c = torch.from_numpy(S.random_coefficients(200, 16, kaula=False, seed=2)).cuda()
s0 = eng.synthesis(c, lon=lon, lat=lat, nmax=200).clone()
a0 = eng.analysis(s0, 200, lon, lat, quadrature="weights", lat_weights=np.cos(lat * np.pi / 180), weight=1e-3).clone()
kern = [p.last_kernels() for p in eng._plans.values()][-1]
s, a = torch.empty_like(s0), torch.empty_like(a0)
nb = 0
for _ in range(100):
    eng.synthesis(c, lon=lon, lat=lat, nmax=200, out=s)
    eng.analysis(s, 200, lon, lat, quadrature="weights", lat_weights=np.cos(lat * np.pi / 180), weight=1e-3, out=a)
    nb += int(not (torch.equal(s, s0) and torch.equal(a, a0)))
print(f"K 11520 {kern} repeats 100 differing {nb}", flush=True)
bad += nb
print("STRESS", "OK" if bad == 0 else f"FAILED ({bad})")
sys.exit(1 if bad else 0)
