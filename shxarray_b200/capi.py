"""ctypes binding of libshb200.so (include/shb200.h).  Thin: pointers, sizes, error codes.

The library is built in-tree (``shxarray_b200/libshb200.so``) by ``__graft_entry__.build()`` /
``make -C shxarray_b200/csrc``.  There is no CPU fallback: if the library is missing the import of
any compute function raises, and without a CUDA device every plan creation fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libshb200.so")

OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED = 0, -1, -2, -3, -4
MODE_GRID, MODE_POINTS = 0, 1
QUAD_NONE, QUAD_SHLIB, QUAD_LATWEIGHTS = 0, 1, 2
FLAG_NO_SYMMETRY, FLAG_DENSE_LON, FLAG_NO_DMMA, FLAG_TIMING, FLAG_NO_PTABLE = 1, 2, 4, 8, 16

# every symbol declared in include/shb200.h
EXPORTS = [
    "shb200_version", "shb200_last_error", "shb200_plan_create", "shb200_plan_destroy", "shb200_plan_get_info",
    "shb200_synthesis", "shb200_analysis", "shb200_synthesis_host", "shb200_analysis_host", "shb200_multiply",
    "shb200_gauss_nodes", "shb200_debug_legendre", "shb200_plan_stage_times", "shb200_plan_set_degree_scale", "shb200_plan_set_order_shard",
]


class PlanDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("nmax", C.c_int32), ("max_batch", C.c_int32),
        ("nsh", C.c_int64), ("n", C.c_void_p), ("m", C.c_void_p),
        ("mode", C.c_int32), ("nlat", C.c_int32), ("lat_deg", C.c_void_p),
        ("nlon", C.c_int32), ("quadrature", C.c_int32), ("lon_deg", C.c_void_p),
        ("weight", C.c_double), ("lat_weights", C.c_void_p), ("flags", C.c_int32), ("ptable_budget_mib", C.c_int32),
    ]


class PlanInfo(C.Structure):
    _fields_ = [
        ("nmax", C.c_int32), ("max_batch", C.c_int32), ("padded_batch", C.c_int32),
        ("nlat", C.c_int32), ("nlon", C.c_int32), ("nrows", C.c_int32),
        ("lon_fft", C.c_int32), ("dmma", C.c_int32),
        ("nsh", C.c_int64), ("nsh_full", C.c_int64), ("workspace_bytes", C.c_int64),
        ("launches_synthesis", C.c_int32), ("launches_analysis", C.c_int32),
    ]


class SHB200Error(RuntimeError):
    pass


_lib = None


def lib():
    """Load libshb200.so (once).  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # a clean checkout: compile the CUDA library in-tree once (nvcc, sm_100a); still no fallback if that fails
        import shutil
        import subprocess
        if shutil.which("make") and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
            subprocess.run(["make", "-s", "-j8", "-C", os.path.join(_HERE, "csrc")], check=False)
    if not os.path.exists(LIB_PATH):
        raise SHB200Error(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C shxarray_b200/csrc` (the engine has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    i32, i64, vp, dbl = C.c_int32, C.c_int64, C.c_void_p, C.c_double
    L.shb200_version.restype = C.c_int
    L.shb200_last_error.restype = C.c_char_p
    L.shb200_plan_create.argtypes = [C.POINTER(PlanDesc), C.POINTER(vp)]
    L.shb200_plan_destroy.argtypes = [vp]
    L.shb200_plan_destroy.restype = None
    L.shb200_plan_get_info.argtypes = [vp, C.POINTER(PlanInfo)]
    L.shb200_synthesis.argtypes = [vp, i32, vp, i64, i64, vp, i64, i64, i64, vp]
    L.shb200_analysis.argtypes = [vp, i32, vp, i64, i64, i64, vp, i64, i64, vp]
    L.shb200_synthesis_host.argtypes = [vp, i32, vp, i64, i64, vp, i64, i64, i64]
    L.shb200_analysis_host.argtypes = [vp, i32, vp, i64, i64, i64, vp, i64, i64]
    L.shb200_multiply.argtypes = [vp, vp, vp, i32, vp, i64, i64, vp, i64, i64, vp, i64, i64, vp]
    L.shb200_plan_stage_times.argtypes = [vp, vp, i32]
    L.shb200_plan_set_degree_scale.argtypes = [vp, i32, vp, i32]
    L.shb200_plan_set_order_shard.argtypes = [vp, i32, i32]
    L.shb200_gauss_nodes.argtypes = [i32, vp, vp]
    L.shb200_debug_legendre.argtypes = [i32, i32, dbl, vp]
    for name in EXPORTS:
        f = getattr(L, name)
        if name not in ("shb200_last_error", "shb200_plan_destroy"):
            f.restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = lib().shb200_last_error().decode(errors="replace")
        if rc == EINVAL:
            raise ValueError(f"shb200: {msg}")
        raise SHB200Error(f"shb200 error {rc}: {msg}")


def gauss_nodes(n):
    """Gauss-Legendre nodes x=sin(lat) (ascending) and weights (sum 2).  Host-only."""
    x = np.zeros(n)
    w = np.zeros(n)
    check(lib().shb200_gauss_nodes(int(n), x.ctypes.data, w.ctypes.data))
    return x, w


def debug_legendre(nmax, lat_deg, device=0):
    out = np.zeros((nmax + 1) * (nmax + 2) // 2)
    check(lib().shb200_debug_legendre(int(device), int(nmax), float(lat_deg), out.ctypes.data))
    return out


class Plan:
    """Owner of one ``shb200_plan``.  Coordinates/index arrays are host numpy arrays."""

    def __init__(self, nmax, lat, lon, max_batch=1, n=None, m=None, mode=MODE_GRID, quadrature=QUAD_NONE,
                 weight=0.0, lat_weights=None, flags=0, device=0, ptable_budget_mib=0):
        L = lib()
        self._h = C.c_void_p()
        self._keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            self._keep.append(a)
            return a

        lat = arr(lat, np.float64)
        lon = arr(lon, np.float64)
        d = PlanDesc()
        d.struct_size = C.sizeof(PlanDesc)
        d.device = int(device)
        d.nmax = int(nmax)
        d.max_batch = int(max_batch)
        if n is not None:
            n = arr(n, np.int32)
            m = arr(m, np.int32)
            if len(n) != len(m):
                raise ValueError("n and m must have the same length")
            d.nsh = len(n)
            d.n = n.ctypes.data
            d.m = m.ctypes.data
        else:
            d.nsh = 0
        d.mode = int(mode)
        d.nlat = len(lat)
        d.lat_deg = lat.ctypes.data
        d.nlon = len(lon)
        d.lon_deg = lon.ctypes.data
        d.quadrature = int(quadrature)
        d.weight = float(weight)
        if lat_weights is not None:
            lw = arr(lat_weights, np.float64)
            if len(lw) != len(lat):
                raise ValueError("lat_weights must have nlat entries")
            d.lat_weights = lw.ctypes.data
        d.flags = int(flags)
        d.ptable_budget_mib = int(ptable_budget_mib)
        check(L.shb200_plan_create(C.byref(d), C.byref(self._h)))
        info = PlanInfo()
        check(L.shb200_plan_get_info(self._h, C.byref(info)))
        self.info = info
        self.device = int(device)
        self.nmax = int(nmax)
        self.max_batch = int(max_batch)
        self.nlat, self.nlon = len(lat), len(lon)
        self.mode = int(mode)
        self.nsh = int(info.nsh)
        self.nsh_full = int(info.nsh_full)

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().shb200_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def multiply_dev(self, syn_b, ana, batch, a_ptr, as_b, as_nm, b_ptr, bs_b, bs_nm, out_ptr, os_b, os_nm, stream=0):
        """self = synthesis plan of operand a; syn_b = synthesis plan of operand b; ana = analysis plan (same grid)."""
        check(lib().shb200_multiply(self._h, syn_b._h, ana._h, batch, a_ptr, as_b, as_nm, b_ptr, bs_b, bs_nm, out_ptr, os_b, os_nm, stream))

    def set_degree_scale(self, which, scale):
        """which: 0 = synthesis input, 1 = analysis output; scale: nmax+1 per-degree factors or None to switch off."""
        if scale is None:
            check(lib().shb200_plan_set_degree_scale(self._h, int(which), None, 0))
        else:
            s = np.ascontiguousarray(scale, dtype=np.float64)
            check(lib().shb200_plan_set_degree_scale(self._h, int(which), s.ctypes.data, len(s)))

    def set_order_shard(self, rank=0, world=1):
        """analysis integrates only orders m = rank (mod world); other coefficients come out as exact zeros."""
        check(lib().shb200_plan_set_order_shard(self._h, int(rank), int(world)))

    def stage_times(self):
        """dict of per-stage device ms of the last synthesis/analysis (plan needs FLAG_TIMING)."""
        ms = np.zeros(6)
        check(lib().shb200_plan_stage_times(self._h, ms.ctypes.data, 6))
        return dict(syn_pack=ms[0], syn_legendre=ms[1], syn_lon=ms[2], ana_lon=ms[3], ana_legendre=ms[4], ana_unpack=ms[5])

    # raw calls: pointers are integers (device or host addresses), strides in elements
    def synthesis_dev(self, batch, coef_ptr, cs_b, cs_nm, grid_ptr, gs_b, gs_lat, gs_lon, stream=0):
        check(lib().shb200_synthesis(self._h, batch, coef_ptr, cs_b, cs_nm, grid_ptr, gs_b, gs_lat, gs_lon, stream))

    def analysis_dev(self, batch, grid_ptr, gs_b, gs_lat, gs_lon, coef_ptr, cs_b, cs_nm, stream=0):
        check(lib().shb200_analysis(self._h, batch, grid_ptr, gs_b, gs_lat, gs_lon, coef_ptr, cs_b, cs_nm, stream))

    def synthesis_host(self, batch, coef_ptr, cs_b, cs_nm, grid_ptr, gs_b, gs_lat, gs_lon):
        check(lib().shb200_synthesis_host(self._h, batch, coef_ptr, cs_b, cs_nm, grid_ptr, gs_b, gs_lat, gs_lon))

    def analysis_host(self, batch, grid_ptr, gs_b, gs_lat, gs_lon, coef_ptr, cs_b, cs_nm):
        check(lib().shb200_analysis_host(self._h, batch, grid_ptr, gs_b, gs_lat, gs_lon, coef_ptr, cs_b, cs_nm))
