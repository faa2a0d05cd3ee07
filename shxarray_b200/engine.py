"""numpy / torch level host API of the B200 engine (no xarray needed).

Mirrors what the reference's ``Synthesis`` / ``Analysis`` operators do with raw arrays
(src/builtin_backend/synthesis.pyx:76-189, analysis.pyx:31-139) but hands the WHOLE batch to the
device in one call (the per-field Python loop of shtns.py:255-259 is what this engine avoids).

* numpy inputs  -> ``shb200_*_host`` (host->device copy, kernels, device->host copy inside the call)
* torch CUDA tensors -> ``shb200_synthesis`` / ``shb200_analysis`` on torch's current stream; torch is
  used purely as the owner of device memory.
"""
import functools
import hashlib
import threading
from collections import OrderedDict

import numpy as np

from . import capi


@functools.lru_cache(maxsize=16)
def _nm_index_cached(nmax, nmin):
    k = np.arange(nmin, nmax + 1, dtype=np.int64)
    n = np.repeat(k, 2 * k + 1).astype(np.int32)
    start = np.repeat(np.cumsum(2 * k + 1) - (2 * k + 1), 2 * k + 1)          # position of (n, -n) in the list
    m = (np.arange(len(n), dtype=np.int64) - start - n).astype(np.int32)
    n.setflags(write=False)
    m.setflags(write=False)
    return n, m


def nm_index(nmax, nmin=0):
    """(n, m) int32 arrays in shxarray's nm order: n-major, m=-n..n (sh_indexing.py:66)."""
    n, m = _nm_index_cached(int(nmax), int(nmin))
    return n.copy(), m.copy()


def _nice_nlon(nmin):
    """smallest 2^a * {1,3,5,9,15} >= nmin (FFT-friendly longitude count)."""
    best = None
    for odd in (1, 3, 5, 9, 15):
        k = odd
        while k < nmin:
            k *= 2
        best = k if best is None or k < best else best
    return best


@functools.lru_cache(maxsize=16)
def _gauss_grid_cached(nmax):
    """Gauss-Legendre grid of nmax (the nodes cost O(nlat^2) long-double operations: 10 ms at nmax 720)."""
    nlat = (nmax + 1 + 3) // 4 * 4
    nlon = _nice_nlon(2 * nmax + 2)
    x, w = capi.gauss_nodes(nlat)
    latv = np.rad2deg(np.arcsin(x))
    half = nlat // 2
    latv[nlat - half:] = -latv[:half][::-1]   # bit-exact mirror symmetry
    lonv = np.arange(nlon) * (360.0 / nlon)
    for a in (lonv, latv, w):
        a.setflags(write=False)
    return lonv, latv, w


def lonlat_grid(nmax=None, lon=None, lat=None, gtype=None):
    """Grid generator.  'regular' / 'regular_lon0' reproduce shlib.pyx:87-96 exactly; 'gauss' is a
    Gauss-Legendre grid (nlat = nmax+1 rounded up to a multiple of 4, nlon FFT-friendly >= 2nmax+2).
    Returns dict(lon=, lat=, gtype=, lat_weights= (gauss only))."""
    if gtype is None:
        gtype = "gauss" if nmax is not None else "regular"
    if gtype not in ("regular", "regular_lon0", "gauss", "point"):
        raise ValueError("Only 'gauss', 'regular', 'regular_lon0' and 'point' grid types are supported")
    if nmax is not None:
        if lon is not None or lat is not None:
            raise ValueError("If nmax is specified, lon and lat should not be specified")
        if gtype in ("regular", "regular_lon0"):
            idres = 1
            while idres > 360 / nmax / 4:
                idres = idres / 2
            lon = np.arange(-180 + idres / 2, 180, idres) if gtype == "regular" else np.arange(0, 360, idres)
            lat = np.arange(-90 + idres / 2, 90, idres)
            return dict(lon=lon, lat=lat, gtype=gtype)
        if gtype == "gauss":
            lonv, latv, w = _gauss_grid_cached(int(nmax))
            return dict(lon=lonv.copy(), lat=latv.copy(), gtype=gtype, lat_weights=w.copy())
        raise ValueError("a 'point' grid needs explicit lon and lat")
    if lon is None or lat is None:
        raise ValueError("Either nmax or lon and lat should be specified")
    lon = np.asarray(lon, dtype=np.float64)
    lat = np.asarray(lat, dtype=np.float64)
    if gtype == "point" and len(lon) != len(lat):
        raise ValueError("For point type 'grid', lon and lat should have the same length")
    return dict(lon=lon, lat=lat, gtype=gtype)


def uniform_step(x):
    """cf.py:25-30: the step if np.diff has min == max exactly, else None."""
    d = np.diff(np.asarray(x, dtype=np.float64))
    if len(d) == 0:
        return None
    return float(d.min()) if d.min() == d.max() else None


def analysis_weight_shlib(lon, lat):
    """|stepx*stepy|/(4 pi) evaluated like analysis.pyx:57-60."""
    sx, sy = uniform_step(lon), uniform_step(lat)
    if sx is None or sy is None:
        raise RuntimeError("grid is not equidistant in lon and lat directions")
    stepx = sx * np.pi / 180
    stepy = sy * np.pi / 180
    return float(abs(stepx * stepy) / (4 * np.pi))


_gauss_cache = {}


def gauss_lat_weights(lat):
    """Gauss-Legendre weights for a latitude vector that consists of the nlat Gauss nodes (either order)."""
    lat = np.asarray(lat, dtype=np.float64)
    key = _digest(lat)
    if key in _gauss_cache:
        return _gauss_cache[key]
    w = _gauss_lat_weights(lat)
    if len(_gauss_cache) > 64:
        _gauss_cache.clear()
    _gauss_cache[key] = w
    return w


def _gauss_lat_weights(lat):
    x, w = capi.gauss_nodes(len(lat))
    s = np.sin(np.deg2rad(lat))
    if np.allclose(s, x, rtol=0, atol=1e-10):
        return w
    if np.allclose(s, x[::-1], rtol=0, atol=1e-10):
        return w[::-1].copy()
    raise RuntimeError("latitudes are not the Gauss-Legendre nodes of this size; cannot use gauss quadrature")


def _is_torch(x):
    return type(x).__module__.startswith("torch")


_w64 = {}


def _weights64(n):
    """three sequences of odd 64-bit position weights (golden-ratio, sqrt(2) and sqrt(3) multipliers)"""
    if n not in _w64:
        if len(_w64) > 16:
            _w64.clear()
        i = np.arange(1, n + 1, dtype=np.uint64)
        _w64[n] = tuple((i * np.uint64(c)) | np.uint64(1) for c in (0x9E3779B97F4A7C15, 0x6A09E667F3BCC909, 0xBB67AE8584CAA73B))
    return _w64[n]


_small_memo = OrderedDict()
_small_lock = threading.Lock()


def _small_digest(a):
    """blake2b of a small contiguous array, memoised per array OBJECT and verified against a private copy (np.array_equal: a
    memcmp, ~5 us for the 4320 longitudes of a 5' grid against ~35 us of hashing), so that repeated calls with the same lon / lat
    arrays -- every transform of a loop -- do not hash them again.  An array changed in place fails the comparison and is
    re-hashed; a recycled id() with other contents likewise."""
    key = id(a)
    with _small_lock:
        hit = _small_memo.get(key)
    if hit is not None and hit[0].shape == a.shape and hit[0].dtype == a.dtype and np.array_equal(hit[0], a):
        return hit[1]
    dg = hashlib.blake2b(a.tobytes(), digest_size=16).digest()
    with _small_lock:
        _small_memo[key] = (a.copy(), dg)
        _small_memo.move_to_end(key)
        while len(_small_memo) > 64:
            _small_memo.popitem(last=False)
    return dg


def _digest(*arrays):
    """Plan-cache key of coordinate / index arrays.  Small arrays: blake2b of the bytes.  Large ones (the nm list of a
    high-degree field is megabytes): length, the plain sum, three position-weighted 64-bit checksums with unrelated weight
    sequences and the first / last 64 elements -- ~1.5 ms instead of ~15 ms per call; an accidental collision needs four
    independent 64-bit sums to agree."""
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        if a is None:
            h.update(b"-")
            continue
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode() + str(a.nbytes).encode())
        if a.nbytes <= 65536:
            h.update(_small_digest(a))
        else:
            raw = a.reshape(-1).view(np.uint8)
            n8 = raw.size // 8 * 8
            h.update(raw[n8:].tobytes())
            v = raw[:n8].view(np.uint64)
            with np.errstate(over="ignore"):
                h.update(int(np.add.reduce(v)).to_bytes(8, "little"))
                for w in _weights64(len(v)):
                    h.update(int(np.add.reduce(v * w)).to_bytes(8, "little"))
            h.update(a.reshape(-1)[:64].tobytes() + a.reshape(-1)[-64:].tobytes())
    return h.hexdigest()


_std_nm = {}


def _std_nm_digest(nmax):
    if nmax not in _std_nm:
        if len(_std_nm) > 32:
            _std_nm.clear()
        n, m = _nm_index_cached(int(nmax), 0)
        _std_nm[nmax] = _digest(np.ascontiguousarray(n, dtype=np.int32), np.ascontiguousarray(m, dtype=np.int32))
    return _std_nm[nmax]


class SHEngine:
    """Plan cache + array front-end.  One instance per process/device is enough (plans are reused)."""

    #: device memory budget for one plan's workspace (bytes); larger batches are processed in chunks
    WORKSPACE_BUDGET = 48 << 30
    #: the plan cache is LRU: at most MAX_PLANS plans and CACHE_BUDGET bytes of device workspace stay alive
    MAX_PLANS = 12
    CACHE_BUDGET = 64 << 30

    def __init__(self, device=0, flags=0):
        self.device = int(device)
        self.flags = int(flags)
        self._plans = OrderedDict()
        self._lock = threading.Lock()
        self.launches = 0   # kernels launched by this engine (gpu_launches accounting in bench.py)

    # ------------------------------------------------------------------------------------------
    def _chunk(self, nmax, nrows_guess, batch):
        tri = (nmax + 1) * (nmax + 2) // 2
        per_field = 16 * tri + 32 * (nmax + 1) * nrows_guess
        return int(max(1, min(batch, self.WORKSPACE_BUDGET // max(per_field, 1))))

    def get_plan(self, nmax, lat, lon, batch, n=None, m=None, mode=capi.MODE_GRID, quadrature=capi.QUAD_NONE,
                 weight=0.0, lat_weights=None, flags=None):
        flags = self.flags if flags is None else flags
        maxb = self._chunk(nmax, len(lat), batch)
        if n is not None and len(n) == (nmax + 1) ** 2:
            # an explicit nm list that is the full triangle in standard order (what the xarray accessor passes) is the
            # default list: same plan as for n = None
            std = _std_nm_digest(nmax)
            if _digest(np.ascontiguousarray(n, dtype=np.int32), np.ascontiguousarray(m, dtype=np.int32)) == std:
                n = m = None
        key = (nmax, mode, quadrature, float(weight), flags, _digest(lat, lon, n, m, lat_weights))
        # same grid and nm list, whatever the quadrature: a plan made for analysis also runs synthesis (the weights only
        # enter the analysis longitude stage), so a round trip on one grid needs ONE plan -- one workspace, one P table
        # (20 GB at nmax 2160 on the 5' grid)
        gkey = (nmax, mode, flags, _digest(lat, lon, n, m))
        with self._lock:
            plan = self._plans.get(key)
            if plan is not None and plan.max_batch >= maxb:
                self._plans.move_to_end(key)
                return plan
            if plan is None and quadrature == capi.QUAD_NONE:
                for k2, p2 in self._plans.items():
                    if getattr(p2, "_gkey", None) == gkey and p2.max_batch >= maxb:
                        self._plans.move_to_end(k2)
                        return p2
            # Replaced, superseded and evicted plans are only DROPPED from the cache: another thread (or the caller, in
            # multiply) may still be inside a call on them.  capi.Plan.__del__ destroys a plan with its last reference.
            if plan is not None:
                del self._plans[key]
            if quadrature != capi.QUAD_NONE:
                # the new plan supersedes synthesis-only plans of the same grid that are not larger
                for k2 in [k2 for k2, p2 in self._plans.items()
                           if k2[2] == capi.QUAD_NONE and getattr(p2, "_gkey", None) == gkey and p2.max_batch <= maxb]:
                    del self._plans[k2]
            del plan
            plan = capi.Plan(nmax, lat, lon, max_batch=maxb, n=n, m=m, mode=mode, quadrature=quadrature,
                             weight=weight, lat_weights=lat_weights, flags=flags, device=self.device)
            plan._gkey = gkey
            self._plans[key] = plan
            # least-recently-used eviction (a plan owns its workspace and, for tensor-core plans, a P_nm table)
            total = sum(int(q.info.workspace_bytes) for q in self._plans.values())
            while len(self._plans) > 3 and (len(self._plans) > self.MAX_PLANS or total > self.CACHE_BUDGET):   # multiply() holds 3 plans
                _, old = self._plans.popitem(last=False)
                total -= int(old.info.workspace_bytes)
                del old
            return plan

    def clear(self):
        with self._lock:
            self._plans.clear()

    #: numpy results up to this size are handed out of the library's page-locked pool (capi.pinned_empty): the device
    #: writes them by DMA, no bounce copy, no first-touch page faults, and a later analysis of that array uploads by DMA too
    PINNED_OUT_MAX = 8 << 30

    def _host_empty(self, shape):
        n = int(np.prod(shape, dtype=np.int64)) * 8
        if 0 < n <= self.PINNED_OUT_MAX:
            try:
                return capi.pinned_empty(shape)
            except capi.SHB200Error:
                pass
        return np.empty(shape, dtype=np.float64)

    # ------------------------------------------------------------------------------------------
    def synthesis(self, coef, n=None, m=None, lon=None, lat=None, nmax=None, points=False, nm_axis=-1,
                  layout="batch_first", out=None, flags=None, degree_scale=None):
        """Spherical-harmonic synthesis of a batch of coefficient vectors.

        coef : [B, nsh] (``nm_axis=-1``) or [nsh, B] (``nm_axis=0``) or [nsh]; numpy (host) or torch CUDA tensor.
        n, m : degree/order of each coefficient (any subset/order; m<0 = sine).  None => full triangle
               in nm order for ``nmax``.
        lon, lat : degrees.  ``points=True``: paired points (synthesis.pyx:186-189) else tensor grid.
        layout : 'batch_first' -> [B, nlat, nlon] ; 'shlib' -> [nlat, nlon, B] (synthesis.pyx:68-71).
        degree_scale : optional nmax+1 per-degree factors fused into the transform (isotropic kernels: Gaussian filter,
               gravity functionals; isokernelbase.py:75-92) -- identical to scaling the coefficients first.
        """
        lon = np.ascontiguousarray(lon, dtype=np.float64)
        lat = np.ascontiguousarray(lat, dtype=np.float64)
        if n is None:
            if nmax is None:
                raise ValueError("either (n, m) or nmax must be given")
            nsh = (nmax + 1) ** 2
        else:
            n = np.ascontiguousarray(n, dtype=np.int32)
            m = np.ascontiguousarray(m, dtype=np.int32)
            nsh = len(n)
            nmax = int(n.max()) if nmax is None else int(nmax)
        tor = _is_torch(coef)
        if not tor:
            coef = np.asarray(coef)
            if coef.dtype != np.float64:
                coef = coef.astype(np.float64)
        elif str(coef.dtype) != "torch.float64":
            raise TypeError("coefficients must be float64")
        squeeze = coef.ndim == 1
        if squeeze:
            coef = coef.reshape(1, nsh) if nm_axis in (-1, 0) else coef
            nm_axis = -1
        if coef.ndim != 2:
            raise ValueError("coef must be 1-D or 2-D ([B, nsh] or [nsh, B])")
        if nm_axis in (-1, 1):
            B = coef.shape[0]
        elif nm_axis == 0:
            coef = coef.T
            B = coef.shape[0]
        else:
            raise ValueError("nm_axis must be 0 or -1")
        if coef.shape[1] != nsh:
            raise ValueError(f"coefficient axis has {coef.shape[1]} entries, nm list has {nsh}")
        mode = capi.MODE_POINTS if points else capi.MODE_GRID
        if points and len(lon) != len(lat):
            raise ValueError("point synthesis needs equally sized lon and lat")
        plan = self.get_plan(nmax, lat, lon, B, n=n, m=m, mode=mode, flags=flags)
        L, K = len(lat), len(lon)
        if points:
            shape = (B, L) if layout == "batch_first" else (L, B)
        else:
            shape = (B, L, K) if layout == "batch_first" else (L, K, B)
        if tor:
            import torch
            if out is None:
                out = torch.empty(shape, dtype=torch.float64, device=coef.device)
            es = lambda t: [s for s in t.stride()]
            ptr = lambda t: t.data_ptr()
            stream = torch.cuda.current_stream(coef.device).cuda_stream
        else:
            if out is None:
                out = self._host_empty(shape)
            elif min(out.strides) < 0:
                raise ValueError("out= views with negative strides are not supported")
            if min(coef.strides) < 0:
                coef = np.ascontiguousarray(coef)
            es = lambda a: [s // 8 for s in a.strides]
            ptr = lambda a: a.ctypes.data
            stream = None
        os_ = es(out)
        if points:
            gs_b, gs_lat, gs_lon = (os_[0], os_[1], 0) if layout == "batch_first" else (os_[1], os_[0], 0)
        else:
            gs_b, gs_lat, gs_lon = (os_[0], os_[1], os_[2]) if layout == "batch_first" else (os_[2], os_[0], os_[1])
        cs = es(coef)
        step = plan.max_batch
        for b0 in range(0, B, step):
            nb = min(step, B - b0)
            cptr = ptr(coef) + 8 * b0 * cs[0]
            gptr = ptr(out) + 8 * b0 * gs_b
            if tor:
                plan.synthesis_dev(nb, cptr, cs[0], cs[1], gptr, gs_b, gs_lat, gs_lon, stream, degree_scale=degree_scale)
            else:
                plan.synthesis_host(nb, cptr, cs[0], cs[1], gptr, gs_b, gs_lat, gs_lon, degree_scale=degree_scale)
            self.launches += plan.info.launches_synthesis
        if squeeze:
            out = out[0] if layout == "batch_first" else out[..., 0]
        return out

    # ------------------------------------------------------------------------------------------
    def _analysis_plan(self, nmax, lon, lat, B, quadrature, lat_weights, weight, flags):
        if quadrature == "shlib":
            q, wgt, lw = capi.QUAD_SHLIB, analysis_weight_shlib(lon, lat), None
        elif quadrature == "gauss":
            q, wgt, lw = capi.QUAD_LATWEIGHTS, 1.0 / (2.0 * len(lon)), gauss_lat_weights(lat)
        elif quadrature == "weights":
            q, wgt, lw = capi.QUAD_LATWEIGHTS, (1.0 if weight is None else float(weight)), np.asarray(lat_weights, dtype=np.float64)
        else:
            raise ValueError("quadrature must be 'shlib', 'gauss' or 'weights'")
        return self.get_plan(nmax, lat, lon, B, quadrature=q, weight=wgt, lat_weights=lw, flags=flags)

    def analysis(self, grid, nmax, lon, lat, quadrature="shlib", lat_weights=None, weight=None, layout="batch_first",
                 out=None, flags=None, degree_scale=None, order_shard=None):
        """Spherical-harmonic analysis of a batch of lon/lat grids (analysis.pyx:31-139).

        grid : [B, nlat, nlon] ('batch_first') or [nlat, nlon, B] ('shlib'); any strides; numpy or torch CUDA.
        quadrature : 'shlib'  -> midpoint rule weight*cos(lat) on a uniform grid (parity with shlib),
                     'gauss'  -> Gauss-Legendre weights (lat must be the Gauss nodes),
                     'weights'-> ``weight * lat_weights[i]``.
        Returns [B, (nmax+1)^2] in nm order (n-major, m=-n..n), nmin=0 (analysis.pyx:29).
        """
        lon = np.ascontiguousarray(lon, dtype=np.float64)
        lat = np.ascontiguousarray(lat, dtype=np.float64)
        tor = _is_torch(grid)
        if not tor:
            grid = np.asarray(grid)
            if grid.dtype != np.float64:
                grid = grid.astype(np.float64)
        elif str(grid.dtype) != "torch.float64":
            raise TypeError("grid must be float64")
        squeeze = grid.ndim == 2
        if squeeze:
            grid = grid[None] if layout == "batch_first" else grid[..., None]
        if grid.ndim != 3:
            raise ValueError("grid must be [B, nlat, nlon] or [nlat, nlon, B]")
        L, K = len(lat), len(lon)
        if layout == "batch_first":
            B = grid.shape[0]
            ok = grid.shape[1] == L and grid.shape[2] == K
        else:
            B = grid.shape[2]
            ok = grid.shape[0] == L and grid.shape[1] == K
        if not ok:
            raise ValueError("grid shape does not match lon/lat")
        plan = self._analysis_plan(nmax, lon, lat, B, quadrature, lat_weights, weight, flags)
        nsh = (nmax + 1) ** 2
        if tor:
            import torch
            if out is None:
                out = torch.empty((B, nsh), dtype=torch.float64, device=grid.device)
            es = lambda t: [s for s in t.stride()]
            ptr = lambda t: t.data_ptr()
            stream = torch.cuda.current_stream(grid.device).cuda_stream
        else:
            if out is None:
                out = self._host_empty((B, nsh))
            elif min(out.strides) < 0:
                raise ValueError("out= views with negative strides are not supported")
            es = lambda a: [s // 8 for s in a.strides]
            ptr = lambda a: a.ctypes.data
            stream = None
        gs = es(grid)
        gs_b, gs_lat, gs_lon = (gs[0], gs[1], gs[2]) if layout == "batch_first" else (gs[2], gs[0], gs[1])
        cs = es(out)
        if not tor and min(gs_b, gs_lat, gs_lon) < 0:
            grid = np.ascontiguousarray(grid)
            gs = es(grid)
            gs_b, gs_lat, gs_lon = (gs[0], gs[1], gs[2]) if layout == "batch_first" else (gs[2], gs[0], gs[1])
        step = plan.max_batch
        for b0 in range(0, B, step):
            nb = min(step, B - b0)
            gptr = ptr(grid) + 8 * b0 * gs_b
            cptr = ptr(out) + 8 * b0 * cs[0]
            if tor:
                plan.analysis_dev(nb, gptr, gs_b, gs_lat, gs_lon, cptr, cs[0], cs[1], stream, degree_scale=degree_scale, order_shard=order_shard)
            else:
                plan.analysis_host(nb, gptr, gs_b, gs_lat, gs_lon, cptr, cs[0], cs[1], degree_scale=degree_scale, order_shard=order_shard)
            self.launches += plan.info.launches_analysis
        if squeeze:
            out = out[0]
        return out


    # ---- multi-GPU exchange helpers (device tensors; distributed.ShardedSH drives them) --------------------------------
    def analysis_packed(self, grid, nmax, lon, lat, order_shard, out, quadrature="shlib", lat_weights=None, weight=None,
                        flags=None, degree_scale=None):
        """Analysis of the orders m = rank (mod world) only; ``out`` [B, >= shard_size] (torch CUDA, last axis contiguous)
        receives this rank's compact m-major shard (shb200_analysis_packed).  grid: torch CUDA [B, nlat, nlon]."""
        import torch
        lon = np.ascontiguousarray(lon, dtype=np.float64)
        lat = np.ascontiguousarray(lat, dtype=np.float64)
        B = grid.shape[0]
        if tuple(grid.shape[1:]) != (len(lat), len(lon)) or out.shape[0] != B or out.stride(1) != 1:
            raise ValueError("analysis_packed: grid must be [B, nlat, nlon] and out [B, S] with a contiguous last axis")
        plan = self._analysis_plan(nmax, lon, lat, B, quadrature, lat_weights, weight, flags)
        if B > plan.max_batch:
            raise ValueError("analysis_packed: batch exceeds one plan call")
        stream = torch.cuda.current_stream(grid.device).cuda_stream
        plan.analysis_packed_dev(B, grid.data_ptr(), grid.stride(0), grid.stride(1), grid.stride(2), out.data_ptr(), out.stride(0),
                                 order_shard, stream, degree_scale=degree_scale)
        self.launches += plan.info.launches_analysis
        return out

    def unpack_shards(self, gathered, nmax, out=None):
        """gathered: torch CUDA [world, B, S] of m-major shards (all ranks) -> [B, (nmax+1)^2] in nm order."""
        import torch
        W, B, _ = gathered.shape
        if out is None:
            out = torch.empty((B, (nmax + 1) ** 2), dtype=torch.float64, device=gathered.device)
        stream = torch.cuda.current_stream(gathered.device).cuda_stream
        capi.unpack_shards_dev(nmax, W, B, gathered.data_ptr(), gathered.stride(0), gathered.stride(1), out.data_ptr(), out.stride(0), out.stride(1), stream)
        self.launches += 1
        return out

    def scatter_rows(self, gathered, row_of_slot, nlat, out=None):
        """gathered: torch CUDA [world, B, rmax, nlon] row blocks of all ranks; row_of_slot: int32 CUDA [world*rmax]
        (latitude index or -1) -> [B, nlat, nlon]."""
        import torch
        W, B, rmax, K = gathered.shape
        if out is None:
            out = torch.empty((B, nlat, K), dtype=torch.float64, device=gathered.device)
        stream = torch.cuda.current_stream(gathered.device).cuda_stream
        capi.scatter_rows_dev(B, W, rmax, K, gathered.data_ptr(), gathered.stride(0), gathered.stride(1), row_of_slot.data_ptr(),
                              out.data_ptr(), out.stride(0), out.stride(1), stream)
        self.launches += 1
        return out

    # ------------------------------------------------------------------------------------------
    def roundtrip(self, coef, nmax, lon, lat, quadrature="gauss", lat_weights=None, weight=None, grid_out=None, out=None):
        """Synthesis to the grid AND analysis of that grid in one pipelined call on host arrays (shb200_roundtrip_host):
        returns (grid [B, nlat, nlon], coef_out [B, (nmax+1)^2]).  Same results and same bytes over PCIe as
        ``g = synthesis(coef); c = analysis(g)`` -- the analysis input is read back from the host grid -- but the grid of
        chunk k+1 is downloaded while the grid of chunk k is uploaded.  coef: numpy [B, (nmax+1)^2], full triangle."""
        lon = np.ascontiguousarray(lon, dtype=np.float64)
        lat = np.ascontiguousarray(lat, dtype=np.float64)
        coef = np.asarray(coef, dtype=np.float64)
        if coef.ndim != 2 or coef.shape[1] != (nmax + 1) ** 2:
            raise ValueError("coef must be [B, (nmax+1)^2]")
        if min(coef.strides) < 0:
            coef = np.ascontiguousarray(coef)
        B, L, K = coef.shape[0], len(lat), len(lon)
        plan = self._analysis_plan(nmax, lon, lat, B, quadrature, lat_weights, weight, None)
        if grid_out is None:
            grid_out = capi.pinned_empty((B, L, K))
        if out is None:
            out = self._host_empty((B, (nmax + 1) ** 2))
        gs = [s // 8 for s in grid_out.strides]
        cs = [s // 8 for s in coef.strides]
        os_ = [s // 8 for s in out.strides]
        step = plan.max_batch
        for b0 in range(0, B, step):
            nb = min(step, B - b0)
            plan.roundtrip_host(nb, coef.ctypes.data + 8 * b0 * cs[0], cs[0], cs[1], grid_out.ctypes.data + 8 * b0 * gs[0], gs[0], gs[1], gs[2],
                                out.ctypes.data + 8 * b0 * os_[0], os_[0], os_[1])
            self.launches += plan.info.launches_synthesis + plan.info.launches_analysis
        return grid_out, out

    # ------------------------------------------------------------------------------------------
    def multiply(self, coef_a, nmax_a, coef_b, nmax_b, out=None):
        """Product of two SH fields through the spatial domain, kept on the device (shtns.py:394-420):
        Gauss grid of nmax_a+nmax_b, two syntheses, pointwise product, Gauss-quadrature analysis.
        coef_a [B, (nmax_a+1)^2], coef_b [B, (nmax_b+1)^2] full triangles in nm order; torch CUDA tensors or numpy.
        Returns [B, (nmax_a+nmax_b+1)^2]; ``out`` (torch CUDA, that shape): written in place -- a loop that calls this with the
        same input and output tensors (the sea-level iteration) replays one captured CUDA graph per call instead of ten launches."""
        import torch
        nmax = int(nmax_a) + int(nmax_b)
        lon, lat, lw = _gauss_grid_cached(nmax)
        host = not _is_torch(coef_a)
        dev = torch.device("cuda", self.device)
        ta = torch.as_tensor(np.ascontiguousarray(coef_a), device=dev) if host else coef_a
        tb = torch.as_tensor(np.ascontiguousarray(coef_b), device=dev) if host else coef_b
        if ta.ndim == 1:
            ta, tb = ta[None], tb[None]
        B = ta.shape[0]
        na, ma = _nm_index_cached(int(nmax_a), 0)
        nb_, mb = _nm_index_cached(int(nmax_b), 0)
        # the analysis plan first: a synthesis operand of the same degree (nmax_a or nmax_b = 0) then shares it
        pc = self.get_plan(nmax, lat, lon, B, quadrature=capi.QUAD_LATWEIGHTS, weight=1.0 / (2.0 * len(lon)), lat_weights=lw)
        pa = self.get_plan(int(nmax_a), lat, lon, B, n=na, m=ma)
        pb = pa if nmax_a == nmax_b else self.get_plan(int(nmax_b), lat, lon, B, n=nb_, m=mb)
        if out is None or host:
            out = torch.empty((B, (nmax + 1) ** 2), dtype=torch.float64, device=dev)
        elif tuple(out.shape) != (B, (nmax + 1) ** 2) or str(out.dtype) != "torch.float64" or not out.is_cuda:
            raise ValueError("multiply: out must be a float64 CUDA tensor of shape [B, (nmax_a+nmax_b+1)^2]")
        stream = torch.cuda.current_stream(dev).cuda_stream
        step = min(pa.max_batch, pb.max_batch, pc.max_batch)
        for b0 in range(0, B, step):
            nb = min(step, B - b0)
            pa.multiply_dev(pb, pc, nb, ta[b0].data_ptr(), ta.stride(0), ta.stride(1), tb[b0].data_ptr(), tb.stride(0), tb.stride(1),
                            out[b0].data_ptr(), out.stride(0), out.stride(1), stream)
            self.launches += 2 * pa.info.launches_synthesis + pc.info.launches_analysis + 1
        return out.cpu().numpy() if host else out


    def apply_mask(self, coef, nmax, mask, lon, lat, nmax_out=None, quadrature="gauss", lat_weights=None, weight=None,
                   syn_scale=None, ana_scale=None):
        """Grid-space operator  analysis(mask x synthesis(coef))  with the grid kept on the device (shb200_apply_mask).

        The sea-level equation's ocean function applied to a load: ``SpectralSeaLevelSolver.oceanf(load)``
        (spectralsealevel.py:97-104) multiplies by an O(nsh^2) product-to-sum matrix that only exists for nmax <= 150
        (:23-24); here the load is synthesised on ``lon`` x ``lat``, multiplied pointwise by the ocean mask and analysed
        back -- any nmax, O(N^3), nothing but coefficients crosses PCIe.  ``syn_scale`` / ``ana_scale``: per-degree
        factors fused into the two transforms (geoid / uplift kernels of ``load_earth``, :106-109).
        coef [B, (nmax+1)^2] (numpy or torch CUDA); mask [nlat, nlon] (numpy or torch CUDA; uploaded once per call when
        numpy -- keep it as a CUDA tensor across SLE iterations).  Returns [B, (nmax_out+1)^2]."""
        import torch
        nmax = int(nmax)
        nmax_out = nmax if nmax_out is None else int(nmax_out)
        lon = np.ascontiguousarray(lon, dtype=np.float64)
        lat = np.ascontiguousarray(lat, dtype=np.float64)
        host = not _is_torch(coef)
        dev = torch.device("cuda", self.device)
        tc = torch.as_tensor(np.ascontiguousarray(coef, dtype=np.float64), device=dev) if host else coef
        tm = torch.as_tensor(np.ascontiguousarray(mask, dtype=np.float64), device=dev) if not _is_torch(mask) else mask
        if tc.ndim == 1:
            tc = tc[None]
        if tuple(tm.shape) != (len(lat), len(lon)):
            raise ValueError("apply_mask: mask must be [nlat, nlon]")
        if tc.shape[1] != (nmax + 1) ** 2:
            raise ValueError("apply_mask: coef must hold the full triangle of nmax")
        B = tc.shape[0]
        n, m = _nm_index_cached(nmax, 0)
        pc = self._analysis_plan(nmax_out, lon, lat, B, quadrature, lat_weights, weight, None)
        pa = self.get_plan(nmax, lat, lon, B, n=n, m=m)
        out = torch.empty((B, (nmax_out + 1) ** 2), dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        step = min(pa.max_batch, pc.max_batch)
        for b0 in range(0, B, step):
            nb = min(step, B - b0)
            pa.apply_mask_dev(pc, nb, tc[b0].data_ptr(), tc.stride(0), tc.stride(1), tm.data_ptr(), tm.stride(0), tm.stride(1),
                              out[b0].data_ptr(), out.stride(0), out.stride(1), stream, syn_scale=syn_scale, ana_scale=ana_scale)
            self.launches += pa.info.launches_synthesis + pc.info.launches_analysis + 1
        return out.cpu().numpy() if host else out


_default = {}


def default_engine(device=0):
    if device not in _default:
        _default[device] = SHEngine(device)
    return _default[device]


def synthesis(coef, n=None, m=None, lon=None, lat=None, **kw):
    return default_engine(kw.pop("device", 0)).synthesis(coef, n, m, lon, lat, **kw)


def analysis(grid, nmax, lon, lat, **kw):
    return default_engine(kw.pop("device", 0)).analysis(grid, nmax, lon, lat, **kw)
