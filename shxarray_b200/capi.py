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
FLAG_NO_SYMMETRY, FLAG_DENSE_LON, FLAG_NO_DMMA, FLAG_TIMING, FLAG_NO_PTABLE, FLAG_NO_POLAR_SKIP = 1, 2, 4, 8, 16, 32

# every symbol declared in include/shb200.h
EXPORTS = [
    "shb200_version", "shb200_last_error", "shb200_plan_create", "shb200_plan_destroy", "shb200_plan_get_info",
    "shb200_synthesis", "shb200_analysis", "shb200_synthesis_host", "shb200_analysis_host", "shb200_multiply",
    "shb200_gauss_nodes", "shb200_debug_legendre", "shb200_plan_stage_times", "shb200_plan_last_kernels", "shb200_roundtrip_host",
    "shb200_host_alloc", "shb200_host_free", "shb200_host_trim", "shb200_debug_host_copy", "shb200_debug_host_chunks",
    "shb200_shard_size", "shb200_analysis_packed", "shb200_unpack_shards", "shb200_scatter_rows", "shb200_apply_mask",
    "shb200_gfc_parse", "shb200_parse_floats", "shb200_scatter_nm", "shb200_polygon_mask", "shb200_debug_fft_cycles",
    "shb200_peer_alloc", "shb200_peer_open", "shb200_peer_close", "shb200_peer_free", "shb200_push_rows",
]
VERSION = 200


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
        ("legendre_visited_synthesis", C.c_double), ("legendre_visited_analysis", C.c_double),
    ]


class CallOpts(C.Structure):
    """shb200_call_opts: what ONE call may change (nothing is kept on the plan)."""
    _fields_ = [("struct_size", C.c_int32), ("shard_rank", C.c_int32), ("shard_world", C.c_int32), ("reserved", C.c_int32),
                ("degree_scale", C.c_void_p)]


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
    po = C.POINTER(CallOpts)
    L.shb200_synthesis.argtypes = [vp, i32, vp, i64, i64, vp, i64, i64, i64, vp, po]
    L.shb200_analysis.argtypes = [vp, i32, vp, i64, i64, i64, vp, i64, i64, vp, po]
    L.shb200_synthesis_host.argtypes = [vp, i32, vp, i64, i64, vp, i64, i64, i64, po]
    L.shb200_analysis_host.argtypes = [vp, i32, vp, i64, i64, i64, vp, i64, i64, po]
    L.shb200_roundtrip_host.argtypes = [vp, i32, vp, i64, i64, vp, i64, i64, i64, vp, i64, i64, po, po]
    L.shb200_plan_last_kernels.argtypes = [vp, C.c_char_p, i32]
    L.shb200_host_alloc.argtypes = [i64, C.POINTER(vp)]
    L.shb200_host_free.argtypes = [vp]
    L.shb200_host_trim.argtypes = []
    L.shb200_debug_fft_cycles.argtypes = [i32, vp]
    L.shb200_debug_host_copy.argtypes = [i32, vp, vp, vp, vp, i32, vp]
    L.shb200_debug_host_chunks.argtypes = [i32, i32, i32, vp]
    L.shb200_shard_size.argtypes = [i32, i32, i32]
    L.shb200_analysis_packed.argtypes = [vp, i32, vp, i64, i64, i64, vp, i64, vp, po]
    L.shb200_unpack_shards.argtypes = [i32, i32, i32, vp, i64, i64, vp, i64, i64, vp]
    L.shb200_scatter_rows.argtypes = [i32, i32, i32, i32, vp, i64, i64, vp, vp, i64, i64, vp]
    L.shb200_peer_alloc.argtypes = [i64, vp, vp]
    L.shb200_peer_open.argtypes = [vp, vp]
    L.shb200_peer_close.argtypes = [vp]
    L.shb200_peer_free.argtypes = [vp]
    L.shb200_push_rows.argtypes = [i32, i32, i32, i32, vp, i64, vp, vp, i64, i64, vp]
    L.shb200_gfc_parse.argtypes = [vp, i64, C.c_char_p, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp]
    L.shb200_parse_floats.argtypes = [vp, vp, vp, i64, vp, vp, vp]
    L.shb200_scatter_nm.argtypes = [vp, vp, vp, i64, i64, i64, i32, i32, vp, i64, i64, vp]
    L.shb200_polygon_mask.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, i32, i32, vp, i64, i64, i64, vp]
    L.shb200_apply_mask.argtypes = [vp, vp, i32, vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, po, po]
    L.shb200_multiply.argtypes = [vp, vp, vp, i32, vp, i64, i64, vp, i64, i64, vp, i64, i64, vp]
    L.shb200_plan_stage_times.argtypes = [vp, vp, i32]
    L.shb200_gauss_nodes.argtypes = [i32, vp, vp]
    L.shb200_debug_legendre.argtypes = [i32, i32, dbl, vp]
    for name in EXPORTS:
        f = getattr(L, name)
        if name not in ("shb200_last_error", "shb200_plan_destroy", "shb200_shard_size"):
            f.restype = C.c_int
    L.shb200_shard_size.restype = C.c_int64
    if L.shb200_version() != VERSION:
        raise SHB200Error(f"{LIB_PATH} is version {L.shb200_version()}, this binding needs {VERSION}: rebuild it (make -C shxarray_b200/csrc)")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = lib().shb200_last_error().decode(errors="replace")
        if rc == EINVAL:
            raise ValueError(f"shb200: {msg}")
        raise SHB200Error(f"shb200 error {rc}: {msg}")


def _opts(degree_scale=None, order_shard=None, nmax=None):
    """(CallOpts or None, keep-alive).  degree_scale: nmax+1 per-degree factors; order_shard: (rank, world)."""
    if degree_scale is None and order_shard is None:
        return None, None
    o = CallOpts()
    o.struct_size = C.sizeof(CallOpts)
    keep = None
    if degree_scale is not None:
        keep = np.ascontiguousarray(degree_scale, dtype=np.float64)
        if nmax is not None and len(keep) != nmax + 1:
            raise ValueError(f"degree_scale needs nmax+1 = {nmax + 1} entries, got {len(keep)}")
        o.degree_scale = keep.ctypes.data
    if order_shard is not None:
        o.shard_rank, o.shard_world = int(order_shard[0]), int(order_shard[1])
    return o, keep


class _PinnedBlock:
    """Owner of one block of the library's page-locked host pool; numpy arrays made from it keep it alive (``.base``)."""

    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        check(lib().shb200_host_alloc(int(nbytes), C.byref(self.ptr)))
        self.nbytes = int(nbytes)
        self.__array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr.value, False), "version": 3}

    def __del__(self):
        try:
            if self.ptr.value:
                lib().shb200_host_free(self.ptr)
                self.ptr = C.c_void_p()
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64):
    """numpy array in page-locked host memory from the library's pool (shb200_host_alloc): transforms DMA straight into /
    out of it.  The block returns to the pool when the last view of the array is garbage collected."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    blk = _PinnedBlock(max(n, 1))
    return np.asarray(blk)[:n].view(dtype).reshape(shape)


def shard_size(nmax, rank, world):
    """elements per field of rank's m-order shard (orders rank, rank+world, ...; cosine and sine interleaved)"""
    n = int(lib().shb200_shard_size(int(nmax), int(rank), int(world)))
    if n < 0:
        raise ValueError("need nmax >= 0 and 0 <= rank < world")
    return n


def unpack_shards_dev(nmax, world, batch, all_ptr, rank_stride, ps_b, coef_ptr, cs_b, cs_nm, stream=0):
    check(lib().shb200_unpack_shards(int(nmax), int(world), int(batch), all_ptr, rank_stride, ps_b, coef_ptr, cs_b, cs_nm, stream))


def scatter_rows_dev(batch, world, rmax, nlon, src_ptr, rank_stride, src_b, row_of_slot_ptr, dst_ptr, dst_b, dst_lat, stream=0):
    check(lib().shb200_scatter_rows(int(batch), int(world), int(rmax), int(nlon), src_ptr, rank_stride, src_b, row_of_slot_ptr, dst_ptr, dst_b, dst_lat, stream))


def peer_alloc(nbytes):
    """Device memory that other processes of this box can map (cudaMalloc + IPC handle).  Returns (pointer, 64-byte handle)."""
    ptr = C.c_void_p()
    handle = C.create_string_buffer(64)
    check(lib().shb200_peer_alloc(int(nbytes), C.byref(ptr), handle))
    return int(ptr.value), handle.raw


def peer_open(handle):
    ptr = C.c_void_p()
    buf = C.create_string_buffer(bytes(handle), 64)
    check(lib().shb200_peer_open(buf, C.byref(ptr)))
    return int(ptr.value)


def peer_close(ptr):
    check(lib().shb200_peer_close(C.c_void_p(int(ptr))))


def peer_free(ptr):
    check(lib().shb200_peer_free(C.c_void_p(int(ptr))))


def push_rows_dev(batch, world, nown, nlon, src_ptr, src_b, rows_ptr, dst_ptrs, dst_b, dst_lat, stream=0):
    """dst_ptrs: sequence of `world` device pointers (base of every rank's grid)."""
    arr = (C.c_void_p * int(world))(*[C.c_void_p(int(p)) for p in dst_ptrs])
    check(lib().shb200_push_rows(int(batch), int(world), int(nown), int(nlon), src_ptr, src_b, rows_ptr, arr, dst_b, dst_lat, stream))


def host_trim():
    """Release the cached (free) blocks of the page-locked pool."""
    check(lib().shb200_host_trim())


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

    def stage_times(self):
        """dict of per-stage device ms of the last synthesis/analysis (plan needs FLAG_TIMING)."""
        ms = np.zeros(6)
        check(lib().shb200_plan_stage_times(self._h, ms.ctypes.data, 6))
        return dict(syn_pack=ms[0], syn_legendre=ms[1], syn_lon=ms[2], ana_lon=ms[3], ana_legendre=ms[4], ana_unpack=ms[5])

    def last_kernels(self):
        """names of the kernels the most recent transform call on this plan launched, in launch order"""
        buf = C.create_string_buffer(512)
        lib().shb200_plan_last_kernels(self._h, buf, 512)
        return [k for k in buf.value.decode().split(",") if k]

    # raw calls: pointers are integers (device or host addresses), strides in elements.  degree_scale / order_shard are
    # per-call options (shb200_call_opts): nothing sticks to the plan.
    def synthesis_dev(self, batch, coef_ptr, cs_b, cs_nm, grid_ptr, gs_b, gs_lat, gs_lon, stream=0, degree_scale=None):
        o, keep = _opts(degree_scale, None, self.nmax)
        check(lib().shb200_synthesis(self._h, batch, coef_ptr, cs_b, cs_nm, grid_ptr, gs_b, gs_lat, gs_lon, stream, o))

    def analysis_dev(self, batch, grid_ptr, gs_b, gs_lat, gs_lon, coef_ptr, cs_b, cs_nm, stream=0, degree_scale=None, order_shard=None):
        o, keep = _opts(degree_scale, order_shard, self.nmax)
        check(lib().shb200_analysis(self._h, batch, grid_ptr, gs_b, gs_lat, gs_lon, coef_ptr, cs_b, cs_nm, stream, o))

    def synthesis_host(self, batch, coef_ptr, cs_b, cs_nm, grid_ptr, gs_b, gs_lat, gs_lon, degree_scale=None):
        o, keep = _opts(degree_scale, None, self.nmax)
        check(lib().shb200_synthesis_host(self._h, batch, coef_ptr, cs_b, cs_nm, grid_ptr, gs_b, gs_lat, gs_lon, o))

    def analysis_host(self, batch, grid_ptr, gs_b, gs_lat, gs_lon, coef_ptr, cs_b, cs_nm, degree_scale=None, order_shard=None):
        o, keep = _opts(degree_scale, order_shard, self.nmax)
        check(lib().shb200_analysis_host(self._h, batch, grid_ptr, gs_b, gs_lat, gs_lon, coef_ptr, cs_b, cs_nm, o))

    def analysis_packed_dev(self, batch, grid_ptr, gs_b, gs_lat, gs_lon, packed_ptr, ps_b, order_shard, stream=0, degree_scale=None):
        """analysis of this rank's orders only, output = the compact m-major shard (shb200_analysis_packed)"""
        o, keep = _opts(degree_scale, order_shard, self.nmax)
        check(lib().shb200_analysis_packed(self._h, batch, grid_ptr, gs_b, gs_lat, gs_lon, packed_ptr, ps_b, stream, o))

    def apply_mask_dev(self, ana, batch, coef_ptr, cs_b, cs_nm, mask_ptr, ms_lat, ms_lon, out_ptr, os_b, os_nm, stream=0,
                       syn_scale=None, ana_scale=None):
        """out = ana.analysis(mask x self.synthesis(coef)) with the grid kept on the device (shb200_apply_mask)"""
        o1, k1 = _opts(syn_scale, None, self.nmax)
        o2, k2 = _opts(ana_scale, None, ana.nmax)
        check(lib().shb200_apply_mask(self._h, ana._h, batch, coef_ptr, cs_b, cs_nm, mask_ptr, ms_lat, ms_lon, out_ptr, os_b, os_nm, stream, o1, o2))

    def roundtrip_host(self, batch, coef_ptr, cs_b, cs_nm, grid_ptr, gs_b, gs_lat, gs_lon, out_ptr, os_b, os_nm,
                       syn_scale=None, ana_scale=None):
        o1, k1 = _opts(syn_scale, None, self.nmax)
        o2, k2 = _opts(ana_scale, None, self.nmax)
        check(lib().shb200_roundtrip_host(self._h, batch, coef_ptr, cs_b, cs_nm, grid_ptr, gs_b, gs_lat, gs_lon, out_ptr, os_b, os_nm, o1, o2))
