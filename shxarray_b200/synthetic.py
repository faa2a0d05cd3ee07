"""Deterministic synthetic inputs for parity tests and bench.py.

No reference data is needed: seeded N(0,1) Stokes-like coefficients, optionally with a Kaula-like
degree decay, in the reference's nm order (n-major, m=-n..n; sh_indexing.py:66).
"""
import numpy as np

SEED = 20240607


def nm_arrays(nmax, nmin=0):
    """(n, m) int32 arrays in shxarray's nm MultiIndex order (sh_indexing.py:66)."""
    n = np.concatenate([np.full(2 * k + 1, k, dtype=np.int32) for k in range(nmin, nmax + 1)])
    m = np.concatenate([np.arange(-k, k + 1, dtype=np.int32) for k in range(nmin, nmax + 1)])
    return n, m


def random_coefficients(nmax, batch, kaula=True, seed=SEED, nmin=0):
    """[batch, nsh] float64.  kaula=True: 1e-5/(n+1)^2 decay with C00=1; False: white N(0,1)."""
    n, m = nm_arrays(nmax, nmin)
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((batch, len(n)))
    if kaula:
        scale = 1e-5 / (n.astype(np.float64) + 1.0) ** 2
        c *= scale[None, :]
        if nmin == 0:
            c[:, 0] = 1.0
    return c


def gauss_grid(nlat, nlon, nodes="library"):
    """Gauss-Legendre latitudes (degrees, ascending, exactly antisymmetric) and lon = arange(nlon)*360/nlon.
    nodes="library": shb200_gauss_nodes (long double Newton); "numpy": numpy.polynomial.legendre.leggauss -- for processes
    that must not load libshb200.so (the CPU reference arm of bench.py)."""
    if nodes == "numpy":
        x, w = np.polynomial.legendre.leggauss(nlat)
    else:
        from .capi import gauss_nodes
        x, w = gauss_nodes(nlat)
    lat = np.rad2deg(np.arcsin(x))
    half = nlat // 2
    lat[nlat - half:] = -lat[:half][::-1]  # enforce bit-exact north/south symmetry
    if nlat % 2:
        lat[half] = 0.0
    lon = np.arange(nlon) * (360.0 / nlon)
    return lon, lat, w
