#!/usr/bin/env python
"""Where the three-pass longitude kernels spend their time: clock64 deltas of thread 0 of every CTA
(shb200_debug_fft_cycles), averaged per CTA, for the BASELINE batched configs.  usage: python tools/fft_phases.py"""
import ctypes as C
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import shxarray_b200 as sb
from shxarray_b200 import capi, synthetic as S

L = capi.lib()
for tag, nmax, B, gtype in (("cfg3", 720, 64, "gauss"), ("cfg2", 96, 240, "regular"), ("n2160 B=16", 2160, 16, "5min")):
    if gtype == "5min":
        d = 1.0 / 12
        h = np.arange(d / 2, 90, d)
        lat, lon = np.concatenate([-h[::-1], h]), np.arange(0, 360, d)
        kw = dict(quadrature=capi.QUAD_LATWEIGHTS, weight=(d * np.pi / 180) ** 2 / (4 * np.pi), lat_weights=np.cos(lat * np.pi / 180))
    else:
        g = sb.lonlat_grid(nmax, gtype=gtype)
        lon, lat = g["lon"], g["lat"]
        kw = dict(quadrature=capi.QUAD_LATWEIGHTS, weight=1.0 / (2.0 * len(lon)), lat_weights=g["lat_weights"]) if gtype == "gauss" else dict(quadrature=capi.QUAD_SHLIB, weight=sb.engine.analysis_weight_shlib(lon, lat))
    plan = capi.Plan(nmax, lat, lon, max_batch=B, flags=capi.FLAG_TIMING, **kw)
    nsh = (nmax + 1) ** 2
    c = torch.from_numpy(S.random_coefficients(nmax, B, kaula=True)).cuda()
    f = torch.empty((B, len(lat), len(lon)), dtype=torch.float64, device="cuda")
    a = torch.empty_like(c)
    st = torch.cuda.current_stream().cuda_stream

    def step():
        plan.synthesis_dev(B, c.data_ptr(), nsh, 1, f.data_ptr(), f.stride(0), f.stride(1), 1, st)
        plan.analysis_dev(B, f.data_ptr(), f.stride(0), f.stride(1), 1, a.data_ptr(), nsh, 1, st)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    L.shb200_debug_fft_cycles(1, None)
    reps = 5
    for _ in range(reps):
        step()
    out = (C.c_int64 * 16)()
    L.shb200_debug_fft_cycles(0, out)
    v = list(out)
    stt = plan.stage_times()
    res = {"config": tag, "K": len(lon), "kernels": plan.last_kernels(), "syn_lon_ms": stt["syn_lon"], "ana_lon_ms": stt["ana_lon"]}
    if v[0]:
        res["syn_items_per_launch"] = v[0] // reps
        res["syn_cycles_per_item"] = dict(zip(("assembly", "pass0", "pass1", "last_pass_and_stores", "wait_for_staged_inputs"), [round(x / v[0]) for x in v[1:6]]))
    if v[8]:
        res["ana_items_per_launch"] = v[8] // reps
        res["ana_cycles_per_item"] = dict(zip(("wait_or_table_fill", "first_pass", "pass1", "pass0", "pickup_and_G_stores"), [round(x / v[8]) for x in v[9:14]]))
    print(json.dumps(res), flush=True)
    plan.close()
    del c, f, a
    torch.cuda.empty_cache()
