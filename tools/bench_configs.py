#!/usr/bin/env python
"""Times the BASELINE.json configs other than the headline one (device-resident, CUDA events) and reports parity of
the synthesis->analysis round trip where the quadrature is exact.  Not the driver's bench (that is bench.py)."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import shxarray_b200 as sb
from shxarray_b200 import capi, synthetic as S


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(tag, nmax, B, gtype, quad, flags=0, lon=None, lat=None, lw=None, weight=None):
    eng = sb.SHEngine(0, flags=flags | capi.FLAG_TIMING)
    if lon is None:
        g = sb.lonlat_grid(nmax, gtype=gtype)
        lon, lat = g["lon"], g["lat"]
    c = torch.from_numpy(S.random_coefficients(nmax, B, kaula=True)).cuda()
    f = eng.synthesis(c, lon=lon, lat=lat, nmax=nmax)
    kw = dict(quadrature=quad)
    if quad == "weights":
        kw.update(lat_weights=lw, weight=weight)
    a = eng.analysis(f, nmax, lon, lat, **kw)
    t_s = timeit(lambda: eng.synthesis(c, lon=lon, lat=lat, nmax=nmax, out=f))
    t_a = timeit(lambda: eng.analysis(f, nmax, lon, lat, out=a, **kw))
    plans = list(eng._plans.values())
    info = {("ana" if p.info.launches_analysis and p is plans[-1] else "syn"): (p.info.dmma, p.info.lon_fft, p.info.nrows) for p in plans}
    err = float(((a - c).abs().amax(dim=1) / c.abs().amax(dim=1)).max())
    stages = {}
    for pl in plans:
        try:
            stages.update({k: round(v, 4) for k, v in pl.stage_times().items() if v > 0})
        except Exception:
            pass
    vis = {}
    for pl in plans:
        vis = dict(synthesis=pl.info.legendre_visited_synthesis, analysis=pl.info.legendre_visited_analysis)
    res = dict(config=tag, nmax=nmax, batch=B, grid=[len(lat), len(lon)], synthesis_ms=t_s, analysis_ms=t_a, legendre_visited=vis,
               fields_per_s_roundtrip=B / ((t_s + t_a) * 1e-3), roundtrip_rel_err=err, kernels=str(info), stages_ms=stages)
    print(json.dumps(res), flush=True)
    eng.clear()
    return res


if __name__ == "__main__":
    only = sys.argv[1:]                 # optional substrings: run only the configs whose tag contains one of them
    out = []

    def add(tag, *a, **k):
        if not only or any((o[1:] == tag) if o.startswith("=") else (o in tag) for o in only):     # "=tag": exact match
            out.append(run(tag, *a, **k))

    add("cfg1 nmax60 shlib 1deg B=1", 60, 1, "regular", "shlib")
    add("cfg2 nmax96 0.5deg B=240", 96, 240, "regular", "shlib")
    add("cfg5 nmax256 gauss B=4", 256, 4, "gauss", "gauss")
    add("cfg3 nmax720 gauss B=64", 720, 64, "gauss", "gauss")
    d = 1.0 / 12
    h = np.arange(d / 2, 90, d)
    lat = np.concatenate([-h[::-1], h])          # bit-exact mirror symmetry (np.arange(-90+d/2, 90, d) is off by ulps for d=1/12)
    lon = np.arange(0, 360, d)
    w = (d * np.pi / 180) ** 2 / (4 * np.pi)
    kw = dict(lon=lon, lat=lat, lw=np.cos(lat * (np.pi / 180)), weight=w)
    add("cfg4 nmax2160 5min B=1", 2160, 1, None, "weights", **kw)
    add("cfg4b nmax2160 5min B=16 (20 GB P table)", 2160, 16, None, "weights", **kw)
    add("cfg4c nmax2160 5min B=16 (fused recursion kernels, no P table)", 2160, 16, None, "weights", flags=capi.FLAG_NO_PTABLE, **kw)
    add("cfg4d nmax2160 5min B=64 (20 GB P table)", 2160, 64, None, "weights", **kw)
    # shlib's own default grid at nmax 2160 (shlib.pyx:87-92: 5760 x 11520): 57 GB P table, longitude length 11520
    add("cfg4e nmax2160 shlib grid 5760x11520 B=16 (57 GB P table)", 2160, 16, "regular", "shlib")
    # shlib's own default grid at nmax 1080 (2880 x 5760): longitude length 5760
    add("cfg6 nmax1080 shlib grid 2880x5760 B=16", 1080, 16, "regular", "shlib")
    json.dump(out, open(os.environ.get("CONFIGS_OUT", "gpurun_out/configs_r02.json"), "w"), indent=1)
