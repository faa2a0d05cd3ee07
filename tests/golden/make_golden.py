"""Generates the committed golden fixtures under tests/golden/.  Run HERE (container with
/root/reference); the GPU box only reads the outputs.

    python tests/golden/make_golden.py

Sources of truth:
  * the reference's own known answers: tests/fixtures.py:66,69-75 (21 lon/lat/value triples, unit
    coefficients, nmax 200 and 20, from the author's rlftlbx Fortran toolbox) and
    tests/testdata/shanalysis-test-paracap-n200.nc (+ the analytic parabolic-cap coefficients of
    kernels/axial.py:47-67 scaled by 1/(2n+1), isokernelbase.py:100-113);
  * oracle/_ref/libshlib_ref.so = the reference's unmodified Legendre_nm.hpp / Ynm.hpp compiled in
    place and driven by the loops of synthesis.pyx:175-189 / analysis.pyx:125-139 with scipy's
    OpenBLAS (oracle/ref_harness.cpp).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from shxarray_b200 import synthetic as S  # noqa: E402

REFROOT = "/root/reference"
assert O.have_ref(), "build oracle/_ref first (make -C oracle)"
KIND = "ref"


def known_answers():
    # tests/fixtures.py:66 and :69-75 (verbatim numbers = data, float32-precision)
    v200 = [(-180, 60, -2.48350572586), (-120, 60, -2.68359804153), (-60, 60, 22.2082977295), (0, 60, 3538.09448242),
            (60, 60, 95.0010375977), (120, 60, -6.011282444), (180, 60, -2.48350572586), (-180, 0, 2.33900880814),
            (-120, 0, -3.7789080143), (-60, 0, -7.44546556473), (0, 0, 559.570007324), (60, 0, 3.45617771149),
            (120, 0, -0.587444782257), (180, 0, 2.33900880814), (-180, -60, 12.755859375), (-120, -60, 6.24348688126),
            (-60, -60, -0.0657368898392), (0, -60, 0.941503584385), (60, -60, 1.90674889088), (120, -60, -21.6708869934),
            (180, -60, 12.755859375)]
    v20 = [(-180, 60, -2.52639126778), (-120, 60, -3.54197382927), (-60, 60, -26.5734081268), (0, 60, 121.698486328),
           (60, 60, 26.0581665039), (120, 60, -5.30710315704), (180, 60, -2.52639126778), (-180, 0, 1.5869371891),
           (-120, 0, -2.11619329453), (-60, 0, -4.59407281876), (0, 0, 33.3332214355), (60, 0, 1.56787621975),
           (120, 0, 0.476241528988), (180, 0, 1.5869371891), (-180, -60, 3.8463101387), (-120, -60, -4.64725017548),
           (-60, -60, -0.319970816374), (0, -60, 0.918827295303), (60, -60, 1.72289800644), (120, -60, -6.23128652573),
           (180, -60, 3.8463101387)]
    with open(os.path.join(HERE, "ref_known_answers.json"), "w") as f:
        json.dump({"source": "reference tests/fixtures.py:66,69-75 (unit coefficients, rlftlbx)", "nmax200": v200, "nmax20": v20}, f, indent=1)


def paracap_coefficients(nmax=200, psi_deg=10.0, lon=(135, -71), lat=(-31, 65)):
    """ParabolicCap(nmax,psi).position(lon,lat): kernels/axial.py:56-67, isokernelbase.py:100-113.
    Returned in nm (n-major) order, [npoints, nsh]."""
    psi = np.radians(psi_deg)
    scales = np.zeros(nmax + 1)
    scales[0] = (1 - np.cos(psi)) / 3
    for n in range(1, nmax + 1):
        scales[n] = ((np.cos((n - 1) * psi) - np.cos(n * psi)) / (n - 0.5) - (np.cos((n + 1) * psi) - np.cos((n + 2) * psi)) / (n + 1.5)) / ((1.0 - np.cos(psi)) * 4.0)
    n, m = O.nm_index(nmax)
    y = O.ynm(n, m, np.asarray(lon, float), np.asarray(lat, float), kind=KIND)
    return y / (2 * n + 1)[None, :] * scales[n][None, :]


def paracap():
    from scipy.io import netcdf_file
    f = netcdf_file(os.path.join(REFROOT, "tests/testdata/shanalysis-test-paracap-n200.nc"), "r", mmap=False)
    z = np.array(f.variables["z"].data, dtype=np.float64)
    x = np.array(f.variables["x"].data, dtype=np.float64)
    y = np.array(f.variables["y"].data, dtype=np.float64)
    coef = paracap_coefficients()
    np.savez_compressed(os.path.join(HERE, "paracap_n200.npz"), z=z, x=x, y=y, cap_coef=coef,
                        note="z,x,y: reference tests/testdata/shanalysis-test-paracap-n200.nc; cap_coef: analytic ParabolicCap(200,10) at (135,-31),(-71,65) in nm order")


def rows_subset(L, k):
    """k row indices spread pole->equator->pole, always including both ends and a symmetric pair."""
    idx = sorted(set(np.linspace(0, L - 1, k).round().astype(int).tolist() + [0, 1, L - 2, L - 1, L // 2]))
    return np.array(idx)


def synthesis_goldens():
    out = {}
    # cfg1-like: nmax 60, shlib regular 1 deg grid, B=2 (kaula, white)
    for tag, nmax, gtype, B, nrows, nlons in (("n60_regular", 60, "regular", 2, 24, None),
                                               ("n96_lon0", 96, "regular_lon0", 4, 16, None),
                                               ("n256_regular", 256, "regular", 2, 12, 64)):
        lon, lat = O.shlib_grid(nmax, gtype)
        n, m = O.nm_index(nmax)
        c = np.concatenate([S.random_coefficients(nmax, B - B // 2, kaula=True), S.random_coefficients(nmax, B // 2, kaula=False, seed=S.SEED + 1)])
        ri = rows_subset(len(lat), nrows)
        ci = np.arange(len(lon)) if nlons is None else np.linspace(0, len(lon) - 1, nlons).round().astype(int)
        g = O.synthesis(c, n, m, lon[ci], lat[ri], grid=True, kind=KIND)  # [rows, cols, B]
        out[tag] = dict(nmax=nmax, gtype=gtype, B=B, rows=ri, cols=ci, values=g)
        print(tag, g.shape)
    np.savez_compressed(os.path.join(HERE, "synthesis_shlib_grids.npz"), **{f"{t}__{k}": v for t, d in out.items() for k, v in d.items()})

    # cfg3-like: nmax 720 on the Gauss grid 724 x 1536 (rows/cols subset), B=2 (kaula, white)
    from shxarray_b200.synthetic import gauss_grid
    lon, lat, _ = gauss_grid(724, 1536)
    nmax = 720
    n, m = O.nm_index(nmax)
    c = np.concatenate([S.random_coefficients(nmax, 1, kaula=True), S.random_coefficients(nmax, 1, kaula=False, seed=S.SEED + 1)])
    ri = rows_subset(724, 8)
    ci = np.linspace(0, 1535, 24).round().astype(int)
    g = O.synthesis(c, n, m, lon[ci], lat[ri], grid=True, kind=KIND)
    np.savez_compressed(os.path.join(HERE, "synthesis_n720_gauss.npz"), nmax=nmax, nlat=724, nlon=1536, rows=ri, cols=ci, values=g, lat=lat, lon=lon)
    print("n720", g.shape)

    # cfg4-like: nmax 2160, 5' grid 2160 x 4320 lon0 (rows/cols subset), B=1 kaula and B=1 white
    nmax = 2160
    d = 1.0 / 12
    lat = np.arange(-90 + d / 2, 90, d)
    lon = np.arange(0, 360, d)
    assert len(lat) == 2160 and len(lon) == 4320
    n, m = O.nm_index(nmax)
    c = np.concatenate([S.random_coefficients(nmax, 1, kaula=True), S.random_coefficients(nmax, 1, kaula=False, seed=S.SEED + 1)])
    ri = np.array([0, 1, 300, 1079, 1080, 1500, 2100, 2158, 2159])
    ci = np.linspace(0, 4319, 12).round().astype(int)
    g = O.synthesis(c, n, m, lon[ci], lat[ri], grid=True, kind=KIND)
    np.savez_compressed(os.path.join(HERE, "synthesis_n2160_5min.npz"), nmax=nmax, rows=ri, cols=ci, values=g, lat=lat, lon=lon)
    print("n2160", g.shape)


def point_and_subset_goldens():
    # point mode, truncated (nmin=3) and shuffled nm list with a duplicate entry, B=3  (Ynm.hpp:84-101 accepts any list)
    nmax, nmin, B = 120, 3, 3
    n, m = O.nm_index(nmax, nmin)
    rng = np.random.default_rng(S.SEED + 2)
    perm = rng.permutation(len(n))
    n, m = n[perm], m[perm]
    n = np.concatenate([n, n[:1]]).astype(np.int32)
    m = np.concatenate([m, m[:1]]).astype(np.int32)
    c = rng.standard_normal((B, len(n)))
    lon = rng.uniform(-180, 360, 40)
    lat = rng.uniform(-90, 90, 40)
    lat[0], lat[1] = 90.0, -90.0
    g = O.synthesis(c, n, m, lon, lat, grid=False, kind=KIND)
    np.savez_compressed(os.path.join(HERE, "synthesis_points_n120.npz"), n=n, m=m, coef=c, lon=lon, lat=lat, values=g)
    print("points", g.shape)


def analysis_goldens():
    # small random grids through the reference analysis loop
    res = {}
    for tag, nmax, gtype, B in (("n60_regular", 60, "regular", 3), ("n48_lon0", 48, "regular_lon0", 2)):
        lon, lat = O.shlib_grid(nmax, gtype)
        rng = np.random.default_rng(S.SEED + 3)
        f = rng.standard_normal((B, len(lat), len(lon)))
        co = O.analysis(f, nmax, lon, lat, kind=KIND)
        res[tag] = dict(nmax=nmax, gtype=gtype, B=B, seed=S.SEED + 3, values=co)
        print(tag, co.shape)
    np.savez_compressed(os.path.join(HERE, "analysis_shlib_grids.npz"), **{f"{t}__{k}": v for t, d in res.items() for k, v in d.items()})


if __name__ == "__main__":
    known_answers()
    paracap()
    analysis_goldens()
    point_and_subset_goldens()
    synthesis_goldens()
