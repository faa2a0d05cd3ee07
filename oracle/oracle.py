"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle.

Two libraries, same entry points (prefix ``port_`` / ``ref_``):

* ``oracle/liboracle_port.so``  -- plain-C restatement (``shlib_port.c``) of the reference
  algorithm (Legendre_nm.hpp:62-132, Ynm.hpp:84-134, synthesis.pyx:175-189, analysis.pyx:125-139).
* ``oracle/_ref/libshlib_ref.so`` -- the reference's unmodified headers compiled in place plus the
  two loops, calling the same scipy-OpenBLAS ``dgemv``/``dger`` shlib calls (``ref_harness.cpp``).
  Present only when built in a container that has ``/root/reference`` (the built .so travels).

``kind='ref'`` selects the second when available, ``kind='port'`` the first.
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(_HERE, "liboracle_port.so")
_REF = os.path.join(_HERE, "_ref", "libshlib_ref.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

_libs = {}


def build(force=False):
    """(Re)build the oracle libraries with the committed Makefile."""
    if force or not os.path.exists(_PORT) or (os.path.isdir("/root/reference") and not os.path.exists(_REF)):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))


def _scipy_openblas():
    import scipy
    cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
    if not cands:
        raise RuntimeError("scipy's bundled OpenBLAS not found")
    return os.path.abspath(cands[0])


def have_ref():
    return os.path.exists(_REF)


def _lib(kind):
    if kind in _libs:
        return _libs[kind]
    if kind == "port":
        build()
        lib = C.CDLL(_PORT)
        pre = "port_"
    elif kind == "ref":
        if not have_ref():
            raise RuntimeError("oracle/_ref/libshlib_ref.so is not built (needs /root/reference)")
        lib = C.CDLL(_REF)
        pre = "ref_"
        lib.ref_load_blas.argtypes = [C.c_char_p]
        lib.ref_load_blas.restype = C.c_int
        rc = lib.ref_load_blas(_scipy_openblas().encode())
        if rc != 0:
            raise RuntimeError(f"ref_load_blas failed rc={rc}")
    else:
        raise ValueError(kind)
    f = getattr(lib, pre + "legendre")
    f.argtypes = [C.c_int, C.c_double, _dp]
    f.restype = None
    f = getattr(lib, pre + "ynm")
    f.argtypes = [C.c_int, _ip, _ip, C.c_int, _dp, _dp, _dp]
    f.restype = None
    f = getattr(lib, pre + "synthesis")
    f.argtypes = [C.c_int, _ip, _ip, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp, _dp, C.c_int]
    f.restype = C.c_int
    f = getattr(lib, pre + "analysis")
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_int, _dp, C.c_int, _dp, C.c_double, _dp, C.c_int]
    f.restype = C.c_int
    _libs[kind] = (lib, pre)
    return _libs[kind]


def best_kind():
    return "ref" if have_ref() else "port"


# ----------------------------------------------------------------------------------------------
# index helpers (sh_indexing.py:48-67)
def nm_index(nmax, nmin=0):
    """(n, m) int32 arrays in the reference's nm MultiIndex order: n-major, m = -n..n."""
    n = np.concatenate([np.full(2 * k + 1, k, dtype=np.int32) for k in range(nmin, nmax + 1)])
    m = np.concatenate([np.arange(-k, k + 1, dtype=np.int32) for k in range(nmin, nmax + 1)])
    return n, m


def tri_size(nmax):
    return (nmax + 1) * (nmax + 2) // 2


def tri_index(n, m, nmax):
    return m * (nmax + 1) - (m * (m + 1)) // 2 + n


def legendre(nmax, costheta, kind=None):
    lib, pre = _lib(kind or best_kind())
    out = np.zeros(tri_size(nmax))
    getattr(lib, pre + "legendre")(int(nmax), float(costheta), out)
    return out


def ynm(n, m, lon, lat, kind=None):
    lib, pre = _lib(kind or best_kind())
    n = np.ascontiguousarray(n, dtype=np.int32)
    m = np.ascontiguousarray(m, dtype=np.int32)
    lon = np.ascontiguousarray(np.atleast_1d(lon), dtype=np.float64)
    lat = np.ascontiguousarray(np.atleast_1d(lat), dtype=np.float64)
    out = np.zeros((len(lon), len(n)))
    getattr(lib, pre + "ynm")(len(n), n, m, len(lon), lon, lat, out)
    return out


def synthesis(coef, n, m, lon, lat, grid=True, nm_first=False, kind=None, nthreads=0):
    """shlib synthesis (synthesis.pyx:76-189).

    coef : [B, nsh] (nm last; default) or [nsh, B] (``nm_first``), float64, C-contiguous.
    Returns [nlat, nlon, B] (grid) or [npoints, B] -- the reference's output layout (aux fastest).
    """
    lib, pre = _lib(kind or best_kind())
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    if coef.ndim == 1:
        coef = coef[None, :] if not nm_first else coef[:, None]
    n = np.ascontiguousarray(n, dtype=np.int32)
    m = np.ascontiguousarray(m, dtype=np.int32)
    lon = np.ascontiguousarray(lon, dtype=np.float64)
    lat = np.ascontiguousarray(lat, dtype=np.float64)
    nsh = len(n)
    if nm_first:
        assert coef.shape[0] == nsh
        B, ld, nm_fast = coef.shape[1], coef.shape[1], 0
    else:
        assert coef.shape[1] == nsh
        B, ld, nm_fast = coef.shape[0], nsh, 1
    if grid:
        out = np.zeros((len(lat), len(lon), B))
    else:
        assert len(lat) == len(lon)
        out = np.zeros((len(lat), B))
    rc = getattr(lib, pre + "synthesis")(nsh, n, m, B, coef.ctypes.data, nm_fast, ld, 1 if grid else 0,
                                          len(lat), lat, len(lon), lon, out.reshape(-1), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle synthesis rc={rc}")
    return out


def analysis_weight(lon, lat):
    """|stepx*stepy|/(4 pi) exactly as analysis.pyx:57-60 (steps from cf.py:25-30 = first diff when uniform)."""
    stepx = np.diff(lon).min() * np.pi / 180
    stepy = np.diff(lat).min() * np.pi / 180
    return float(abs(stepx * stepy) / (4 * np.pi))


def analysis(grid, nmax, lon, lat, kind=None, nthreads=0, weight=None):
    """shlib analysis (analysis.pyx:31-139).  grid : [B, nlat, nlon] float64 (any strides).  Returns [B, (nmax+1)^2]."""
    lib, pre = _lib(kind or best_kind())
    grid = np.asarray(grid, dtype=np.float64)
    if grid.ndim == 2:
        grid = grid[None]
    lon = np.ascontiguousarray(lon, dtype=np.float64)
    lat = np.ascontiguousarray(lat, dtype=np.float64)
    B, L, K = grid.shape
    assert L == len(lat) and K == len(lon)
    if weight is None:
        weight = analysis_weight(lon, lat)
    s = [x // 8 for x in grid.strides]
    nsh = (nmax + 1) ** 2
    out = np.zeros((B, nsh))
    rc = getattr(lib, pre + "analysis")(int(nmax), B, grid.ctypes.data, s[1], s[2], s[0], L, lat, K, lon,
                                         float(weight), out.reshape(-1), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle analysis rc={rc}")
    return out


def set_exact_trig(on):
    """Diagnostic switch of the C port only (see shlib_port.c): trig arguments formed in long double instead of the
    reference's doubly rounded k*lonr.  NOT reference behaviour; always switch it off again."""
    lib, _ = _lib("port")
    lib.port_set_exact_trig.argtypes = [C.c_int]
    lib.port_set_exact_trig.restype = None
    lib.port_set_exact_trig(1 if on else 0)


def max_threads():
    if have_ref():
        lib, _ = _lib("ref")
        return int(lib.ref_max_threads())
    return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------
# grid helper (shlib.pyx:61-117)
def shlib_grid(nmax, gtype="regular"):
    idres = 1
    while idres > 360 / nmax / 4:
        idres = idres / 2
    if gtype == "regular":
        lon = np.arange(-180 + idres / 2, 180, idres)
    elif gtype == "regular_lon0":
        lon = np.arange(0, 360, idres)
    else:
        raise ValueError(gtype)
    lat = np.arange(-90 + idres / 2, 90, idres)
    return lon, lat
