"""Race detector: repeated transforms of several kernel shapes against the bit-faithful DFMA kernels."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import shxarray_b200 as sb
from shxarray_b200 import capi, synthetic as S
eng = sb.SHEngine(0)
for nmax,B,reps,fl in ((360,40,60,capi.FLAG_NO_PTABLE),(96,100,60,capi.FLAG_NO_PTABLE),(200,64,60,0),(96,240,60,0),
                       (600,12,40,0),(300,24,60,0),(720,16,30,0),(400,3,40,0),(2160,1,6,0)):   # narrow table tiles, tiny-batch kernels
    g = sb.lonlat_grid(nmax, gtype="regular")
    c = torch.from_numpy(S.random_coefficients(nmax, B, kaula=False)).cuda()
    ref = eng.synthesis(c, lon=g['lon'], lat=g['lat'], nmax=nmax, flags=capi.FLAG_NO_DMMA)
    aref = eng.analysis(ref, nmax, g['lon'], g['lat'], flags=capi.FLAG_NO_DMMA)
    den = ref.abs().amax(dim=(1,2)); aden = aref.abs().amax(dim=1)
    nb_s = nb_a = 0; ws = wa = 0.0
    for i in range(reps):
        a = eng.synthesis(c, lon=g['lon'], lat=g['lat'], nmax=nmax, flags=fl)
        e = float(((a-ref).abs().amax(dim=(1,2))/den).max()); ws = max(ws,e); nb_s += e > 1e-12
        an = eng.analysis(ref, nmax, g['lon'], g['lat'], flags=fl)
        e = float(((an-aref).abs().amax(dim=1)/aden).max()); wa = max(wa,e); nb_a += e > 1e-12
    print(nmax, B, 'flags', fl, 'reps', reps, 'syn bad', nb_s, ws, 'ana bad', nb_a, wa, flush=True)
    eng.clear()
