"""Round-2 golden fixtures: the oracle meets the HEADLINE kernels directly.

Run HERE (container with /root/reference and oracle/_ref built); the GPU box only reads the outputs.

    python tests/golden/make_golden_r2.py            # ~6 minutes on 8 threads

Every vector comes from oracle/_ref/libshlib_ref.so = the reference's unmodified Legendre_nm.hpp / Ynm.hpp driven by
the loops of synthesis.pyx:175-189 / analysis.pyx:125-139 with scipy's OpenBLAS (oracle/ref_harness.cpp).  The batch
sizes are the BASELINE ones (64 fields at nmax 720, 240 at nmax 96, 16 at nmax 2160) so that on the GPU the plans
select k_pack_n -> k_leg_syn_tab -> k_lon_syn_fft (8 fields per CTA) and k_lon_ana_fft -> k_leg_ana_tab -> k_unpack_n,
i.e. the kernels the bench line is measured on, not the tiny-batch family.

How a full-size transform is checked with seconds of oracle time:
  synthesis : shlib evaluates every grid point independently (synthesis.pyx:179-183), so the oracle is run on a
              subset of rows x columns of the full grid; the GPU transforms the full grid and is compared there.
  analysis  : shlib's analysis is a sum over grid points (analysis.pyx:128-139), so a grid that is zero except on a
              subset of rows x columns has exactly the coefficients of the sum over that subset: the oracle loops
              over the subset only, the GPU transforms the full (mostly zero) grid with the cfg plan shape.
              Gauss / explicit latitude weights are folded into the data handed to the oracle (it multiplies by
              weight*cos(lat), analysis.pyx:129) -- one rounding, 1e-16.
  outputs   : analysis vectors are stored at a subset of nm positions that covers EVERY order m (cosine and sine)
              at degrees m, nmax and random ones in between, for every field, plus max|ref| over the full vector
              of each field (the normalisation of the parity metric).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from shxarray_b200 import synthetic as S  # noqa: E402

assert O.have_ref(), "build oracle/_ref first (make -C oracle)"
KIND = "ref"
NTHR = os.cpu_count() or 1


def gauss_grid_np(nlat, nlon):
    """Same grid as shxarray_b200.lonlat_grid(gtype='gauss') / synthetic.gauss_grid, stored in the fixture (the tests use the
    stored vectors, so the two never have to agree bit for bit)."""
    from shxarray_b200.synthetic import gauss_grid
    return gauss_grid(nlat, nlon)


def spread_rows(L, k, pairs=True):
    """k row indices spread over the latitude range; with ``pairs`` every picked row comes with its mirror row L-1-i
    (both rows of a mirror pair carry data) except two deliberately single ones."""
    base = np.linspace(0, L // 2 - 1, k // 2).round().astype(int)
    idx = set(base.tolist())
    if pairs:
        idx |= {L - 1 - i for i in base[:-2]}
        idx |= {L // 2 + 3, L // 2 + 40}
    else:
        idx |= set((L - 1 - base).tolist())
    return np.array(sorted(i for i in idx if 0 <= i < L))


def pick_positions(nmax, rng, per_m=2):
    """nm positions n^2+n+m covering every |m| (cos and sin) at n = |m|, nmax and ``per_m`` random degrees."""
    pos = set()
    for am in range(nmax + 1):
        degs = {am, nmax} | set(rng.integers(am, nmax + 1, per_m).tolist())
        for n in degs:
            pos.add(n * n + n + am)
            if am:
                pos.add(n * n + n - am)
    return np.array(sorted(pos), dtype=np.int64)


def mixed_coefficients(nmax, B, seed0):
    """first half white N(0,1) (worst case for parity), second half Kaula-scaled; the test regenerates them from the seeds."""
    return np.concatenate([S.random_coefficients(nmax, B // 2, kaula=False, seed=seed0),
                           S.random_coefficients(nmax, B - B // 2, kaula=True, seed=seed0 + 1)])


def sparse_grid_values(B, nr, nc, seed):
    return np.random.default_rng(seed).standard_normal((B, nr, nc))


def save(name, **kw):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **kw)
    print(f"  wrote {name}: {os.path.getsize(path) / 1e6:.2f} MB")


def timed(label):
    class T:
        def __enter__(self):
            self.t = time.time()
            print(label, "...", flush=True)

        def __exit__(self, *a):
            print(f"  {label}: {time.time() - self.t:.1f} s", flush=True)
    return T()


# ----------------------------------------------------------------------------------------------------------------------
def syn_cfg3_b64():
    nmax, B, seed = 720, 64, 7001
    lon, lat, _ = gauss_grid_np(724, 1536)
    n, m = O.nm_index(nmax)
    c = mixed_coefficients(nmax, B, seed)
    rows = spread_rows(724, 16)
    cols = np.linspace(0, 1535, 24).round().astype(int)
    with timed("synthesis cfg3 B=64"):
        g = O.synthesis(c, n, m, lon[cols], lat[rows], grid=True, kind=KIND, nthreads=NTHR)      # [rows, cols, B]
    save("r2_syn_cfg3_b64.npz", nmax=nmax, B=B, seed=seed, lon=lon, lat=lat, rows=rows, cols=cols, values=g)


def syn_cfg2_b240():
    nmax, B, seed = 96, 240, 7101
    lon, lat = O.shlib_grid(nmax, "regular")
    assert len(lat) == 360 and len(lon) == 720
    n, m = O.nm_index(nmax)
    c = mixed_coefficients(nmax, B, seed)
    rows = spread_rows(360, 16)
    cols = np.linspace(0, 719, 32).round().astype(int)
    with timed("synthesis cfg2 B=240"):
        g = O.synthesis(c, n, m, lon[cols], lat[rows], grid=True, kind=KIND, nthreads=NTHR)
    save("r2_syn_cfg2_b240.npz", nmax=nmax, B=B, seed=seed, rows=rows, cols=cols, values=g)


def syn_n2160_b16():
    nmax, B, seed = 2160, 16, 7201
    d = 1.0 / 12
    lat = np.arange(-90 + d / 2, 90, d)
    lon = np.arange(0, 360, d)
    n, m = O.nm_index(nmax)
    c = mixed_coefficients(nmax, B, seed)
    rows = np.array([0, 1, 300, 1079, 1080, 1500, 1859, 2100, 2158, 2159])
    cols = np.linspace(0, 4319, 10).round().astype(int)
    with timed("synthesis nmax 2160 B=16"):
        g = O.synthesis(c, n, m, lon[cols], lat[rows], grid=True, kind=KIND, nthreads=NTHR)
    save("r2_syn_n2160_b16.npz", nmax=nmax, B=B, seed=seed, rows=rows, cols=cols, values=g)


def _analysis_case(name, nmax, B, lon, lat, rows, cols, seed, weight, lat_weights=None, store_grid=False, per_m=2):
    """Oracle analysis of the sparse grid (rows x cols of the full grid carry N(0,1) data, everything else is zero)."""
    vals = sparse_grid_values(B, len(rows), len(cols), seed)
    data = vals
    if lat_weights is not None:      # fold w_i / cos(lat_i) into the data: the oracle multiplies by weight*cos(lat_i)
        data = vals * (lat_weights[rows] / np.cos(lat[rows] * (np.pi / 180)))[None, :, None]
    with timed(f"analysis {name}"):
        co = O.analysis(np.ascontiguousarray(data), nmax, lon[cols], lat[rows], kind=KIND, weight=weight, nthreads=NTHR)
    pos = pick_positions(nmax, np.random.default_rng(seed + 1), per_m)
    kw = dict(nmax=nmax, B=B, seed=seed, rows=rows, cols=cols, weight=weight, positions=pos, values=co[:, pos],
              fieldmax=np.abs(co).max(axis=1))
    if store_grid:
        kw.update(lon=lon, lat=lat)
    if lat_weights is not None:
        kw["lat_weights"] = lat_weights
    save(name, **kw)


def ana_cfg3_b64():
    lon, lat, w = gauss_grid_np(724, 1536)
    rows = spread_rows(724, 32)
    cols = np.linspace(0, 1535, 24).round().astype(int)
    _analysis_case("r2_ana_cfg3_b64.npz", 720, 64, lon, lat, rows, cols, 7301, 1.0 / (2.0 * 1536), lat_weights=w, store_grid=True)


def ana_cfg2_b240():
    lon, lat = O.shlib_grid(96, "regular")
    rows = spread_rows(360, 32)
    cols = np.linspace(0, 719, 48).round().astype(int)
    _analysis_case("r2_ana_cfg2_b240.npz", 96, 240, lon, lat, rows, cols, 7401, O.analysis_weight(lon, lat), per_m=6)


def ana_n256_b4():
    lon, lat = O.shlib_grid(256, "regular")          # 0.25 deg: 720 x 1440 (cfg5 on shlib's own grid)
    assert len(lat) == 720 and len(lon) == 1440
    rows = spread_rows(720, 24)
    cols = np.linspace(0, 1439, 32).round().astype(int)
    _analysis_case("r2_ana_n256_b4.npz", 256, 4, lon, lat, rows, cols, 7501, O.analysis_weight(lon, lat), per_m=4)


def ana_n2160_b16():
    d = 1.0 / 12
    lat = np.arange(-90 + d / 2, 90, d)
    lon = np.arange(0, 360, d)
    rows = spread_rows(2160, 16)
    cols = np.linspace(0, 4319, 16).round().astype(int)
    # the 5' grid is not exactly equidistant in FP64 (shlib's own step test would reject it): explicit midpoint weights,
    # the same numbers analysis.pyx:57-60,129 would use
    weight = (d * np.pi / 180) ** 2 / (4 * np.pi)
    _analysis_case("r2_ana_n2160_b16.npz", 2160, 16, lon, lat, rows, cols, 7601, weight, per_m=1)


def syn_shlib2160_b16():
    """nmax 2160 on shlib's OWN default grid (shlib.pyx:87-92: 1/32 degree, 5760 x 11520, lon from -180 + 1/64): the grid
    whose P table (57 GB) exceeded the round-1 budget."""
    nmax, B, seed = 2160, 16, 7801
    lon, lat = O.shlib_grid(nmax, "regular")
    assert len(lat) == 5760 and len(lon) == 11520
    n, m = O.nm_index(nmax)
    c = mixed_coefficients(nmax, B, seed)
    rows = np.array([0, 1, 700, 2879, 2880, 4000, 5000, 5758, 5759])
    cols = np.linspace(0, 11519, 8).round().astype(int)
    with timed("synthesis nmax 2160 on the shlib grid B=16"):
        g = O.synthesis(c, n, m, lon[cols], lat[rows], grid=True, kind=KIND, nthreads=NTHR)
    save("r2_syn_shlib2160_b16.npz", nmax=nmax, B=B, seed=seed, rows=rows, cols=cols, values=g)


def ana_shlib2160_b16():
    lon, lat = O.shlib_grid(2160, "regular")
    rows = spread_rows(5760, 12)
    cols = np.linspace(0, 11519, 8).round().astype(int)
    _analysis_case("r2_ana_shlib2160_b16.npz", 2160, 16, lon, lat, rows, cols, 7901, O.analysis_weight(lon, lat), per_m=1)


def multiply_128():
    """engine.multiply pattern (shtns.py:394-420) entirely through the oracle: two syntheses on the Gauss grid of
    nmax 128+128, pointwise product, analysis with Gauss weights."""
    na, nb, B, seed = 128, 128, 4, 7701
    nmax = na + nb
    import shxarray_b200 as sb
    g = sb.lonlat_grid(nmax, gtype="gauss")
    lon, lat, w = g["lon"], g["lat"], g["lat_weights"]
    a = S.random_coefficients(na, B, kaula=False, seed=seed)
    b = S.random_coefficients(nb, B, kaula=False, seed=seed + 1)
    n1, m1 = O.nm_index(na)
    n2, m2 = O.nm_index(nb)
    with timed("multiply: syntheses"):
        fa = O.synthesis(a, n1, m1, lon, lat, kind=KIND, nthreads=NTHR)       # [lat, lon, B]
        fb = O.synthesis(b, n2, m2, lon, lat, kind=KIND, nthreads=NTHR)
    prod = np.moveaxis(fa * fb, 2, 0)                                         # [B, lat, lon]
    data = prod * (w / np.cos(lat * (np.pi / 180)))[None, :, None]
    with timed("multiply: analysis"):
        co = O.analysis(np.ascontiguousarray(data), nmax, lon, lat, kind=KIND, weight=1.0 / (2.0 * len(lon)), nthreads=NTHR)
    pos = pick_positions(nmax, np.random.default_rng(seed + 2), 6)
    save("r2_multiply_128.npz", na=na, nb=nb, B=B, seed=seed, lon=lon, lat=lat, positions=pos, values=co[:, pos],
         fieldmax=np.abs(co).max(axis=1))


if __name__ == "__main__":
    which = sys.argv[1:] or ["syn_cfg3_b64", "syn_cfg2_b240", "syn_n2160_b16", "ana_cfg2_b240", "ana_n256_b4", "ana_cfg3_b64",
                             "ana_n2160_b16", "multiply_128"]
    for w in which:
        globals()[w]()
