"""Multi-GPU planner: one process per GPU (torch.distributed, NCCL over NVLink on the B200 box; gloo in CPU tests).

The reference has no distributed path at all (single process, OpenMP over latitude rows:
synthesis.pyx:179, analysis.pyx:128).  The path shards three ways:

* batch sharding  -- fields are independent (the dgemv/dger columns never mix: synthesis.pyx:183,
  analysis.pyx:136): every rank transforms its own contiguous slice of the batch, NO collective.
  This is what bench.py scales (weak scaling).
* latitude sharding of a single (few-field) synthesis -- rows are independent (prange(nlat)); mirror latitude pairs stay
  on one rank so the E/O symmetry survives.  Every rank synthesises its rows straight into its slot of a
  [world, B, rmax, nlon] buffer; ONE all-gather (ncclAllGather, in place); one kernel (shb200_scatter_rows) puts the rows
  where they belong in the [B, nlat, nlon] grid.  With ``gather=False`` the rows stay sharded and nothing is exchanged.
* m-order sharding of a single analysis (BASELINE config 4, what north_star names) -- every rank holds the grid, runs the
  cheap longitude stage for all rows and integrates only the orders m = rank (mod world) (interleaved because the work per
  order is proportional to nmax-m+1).  It emits ONLY those orders, packed m-major (4.6 MB per rank at nmax 2160);
  ONE all-gather of the shards (37 MB in total); one kernel (shb200_unpack_shards) writes the nm vector -- bit-identical to
  the unsharded transform.  Latitude-sharded analysis (+ an all_reduce of partial sums) remains as the alternative for
  grids that are already distributed by rows.

The engine is injectable so the planner logic is testable on CPU (world_size 2, gloo) with the oracle; the numpy
functions below define the exchange formats (include/shb200.h) and serve those tests.
"""
import numpy as np


def shard_bounds(total, world, rank):
    """Contiguous balanced partition of range(total): sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


_shard_cache = {}


def lat_shard_indices(lat, world, rank):
    """Latitude indices owned by ``rank``: mirror pairs (lat, -lat) are kept together, unpaired rows are dealt last."""
    lat = np.asarray(lat, dtype=np.float64)
    key = (lat.tobytes(), int(world))
    if key not in _shard_cache:
        if len(_shard_cache) > 8:
            _shard_cache.clear()
        _shard_cache[key] = [_lat_shard_indices(lat, world, r) for r in range(world)]
    return _shard_cache[key][rank]


def _lat_shard_indices(lat, world, rank):
    L = len(lat)
    used = np.zeros(L, dtype=bool)
    where = {}
    for i, v in enumerate(lat):
        where.setdefault(float(v), []).append(i)
    units = []
    for i in range(L):
        if used[i]:
            continue
        used[i] = True
        mate = None
        if lat[i] != 0.0:
            for j in where.get(float(-lat[i]), []):
                if not used[j]:
                    mate = j
                    break
        if mate is not None:
            used[mate] = True
            units.append((i, mate))
        else:
            units.append((i,))
    lo, hi = work_bounds([_row_work(lat[u[0]], world) for u in units], world, rank)
    idx = sorted(k for u in units[lo:hi] for k in u)
    return np.asarray(idx, dtype=np.int64)


def _row_work(lat_deg, world=1):
    """Relative cost of a latitude row for a rank of a latitude-sharded transform.  Legendre stage: polar skipping (table kernels
    and seeded recursion kernels alike) leaves out the degrees n below ~m / cos(lat) of every order m, so a row costs about
    cos(lat) of an equatorial one plus a floor for the 1e-24 margin -- one nmax-2160 field visits 74 % of the triples, i.e.
    0.28 + 0.72 cos(lat) averaged over the rows.  Longitude stage and row exchange cost the same for every row: measured on the
    single-field nmax-2160 synthesis, 9 % of an equatorial row's Legendre time for the FFT and 4.5 % per peer for the NVLink
    stores of `shb200_push_rows` (`profiles/r02_multigpu_breakdown_n8.json`)."""
    return 0.28 + 0.72 * abs(np.cos(float(lat_deg) * np.pi / 180.0)) + 0.09 + 0.045 * (int(world) - 1)


def work_bounds(weights, world, rank):
    """Contiguous partition of the units into ``world`` runs of (nearly) equal total weight; every rank gets at least one unit
    when there are enough.  Equal weights reproduce shard_bounds."""
    n, world = len(weights), int(world)
    if n <= world or len(set(weights)) == 1:
        return shard_bounds(n, world, rank)
    cum = np.concatenate([[0.0], np.cumsum(np.asarray(weights, dtype=np.float64))])
    cuts = [0]
    for r in range(1, world):
        k = int(np.searchsorted(cum, cum[-1] * r / world, side="left"))
        if k > 0 and abs(cum[k - 1] - cum[-1] * r / world) <= abs(cum[k] - cum[-1] * r / world):
            k -= 1                                    # the nearer of the two candidate cuts
        cuts.append(min(max(k, cuts[-1] + 1), n - (world - r)))
    cuts.append(n)
    return cuts[rank], cuts[rank + 1]


# ---- exchange formats (numpy definitions; the CUDA kernels of csrc/shard.cu implement the same) -----------------------
def shard_size(nmax, rank, world):
    """elements per field of rank's m-order shard: orders rank, rank+world, ..., cosine and sine interleaved"""
    return int(sum(2 * (nmax - m + 1) for m in range(rank, nmax + 1, world)))


def pack_orders_np(coef, nmax, rank, world):
    """[B, (nmax+1)^2] nm vector -> [B, shard_size] m-major shard of the orders m = rank (mod world):
    element off_j + 2 (n - m_j) + cs  with  m_j = rank + j world  (include/shb200.h: shb200_analysis_packed)."""
    coef = np.asarray(coef)
    out = np.zeros((coef.shape[0], shard_size(nmax, rank, world)))
    off = 0
    for m in range(rank, nmax + 1, world):
        n = np.arange(m, nmax + 1)
        out[:, off:off + 2 * len(n):2] = coef[:, n * n + n + m]
        if m > 0:
            out[:, off + 1:off + 2 * len(n):2] = coef[:, n * n + n - m]
        off += 2 * len(n)
    return out


def unpack_shards_np(gathered, nmax):
    """[world, B, >= shard_size] gathered shards -> [B, (nmax+1)^2] in nm order (shb200_unpack_shards)."""
    gathered = np.asarray(gathered)
    W, B = gathered.shape[0], gathered.shape[1]
    out = np.zeros((B, (nmax + 1) ** 2))
    for r in range(W):
        off = 0
        for m in range(r, nmax + 1, W):
            n = np.arange(m, nmax + 1)
            out[:, n * n + n + m] = gathered[r, :, off:off + 2 * len(n):2]
            if m > 0:
                out[:, n * n + n - m] = gathered[r, :, off + 1:off + 2 * len(n):2]
            off += 2 * len(n)
    return out


class _Dist:
    """Thin view of torch.distributed that degrades to a single rank when no process group exists."""

    def __init__(self, group=None):
        try:
            import torch.distributed as dist
            self.dist = dist if dist.is_available() and dist.is_initialized() else None
        except Exception:
            self.dist = None
        self.group = group

    @property
    def world(self):
        return self.dist.get_world_size(self.group) if self.dist else 1

    @property
    def rank(self):
        return self.dist.get_rank(self.group) if self.dist else 0

    def all_gather_inplace(self, buf):
        """buf [world, ...] with this rank's slot filled -> every slot filled.  ONE collective (ncclAllGather in place)."""
        if self.world == 1:
            return buf
        mine = buf[self.rank]
        try:
            self.dist.all_gather_into_tensor(buf, mine, group=self.group)
        except (RuntimeError, NotImplementedError):
            # backends without the flat all-gather (older gloo): list form, then copy into place
            parts = [buf[r].clone() for r in range(self.world)]
            self.dist.all_gather(parts, mine.clone(), group=self.group)
            for r in range(self.world):
                buf[r].copy_(parts[r])
        return buf


class _CudaView:
    """Device memory owned by the C library, exposed through __cuda_array_interface__ so that torch can wrap it without a copy."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class _PeerGrids:
    """The grids of all ranks of one box, each mapped into every process (CUDA IPC): the latitude-sharded synthesis stores its
    rows straight into every rank's grid over NVLink (shb200_push_rows) instead of all-gathering a staging buffer and scattering
    it.  Two grids per rank, used alternately: a rank may be one exchange ahead of the slowest one, never two (every exchange
    ends with a barrier), so the grid a slow rank is still reading is never the one being written."""

    def __init__(self, d, capi, torch, device, shape):
        self.capi, self.shape, self.d = capi, tuple(shape), d
        nbytes = int(np.prod(shape)) * 8
        self.own, self.peers, self.opened = [], [], []
        ok = 1
        handles = [None, None]
        try:
            for k in range(2):
                ptr, h = capi.peer_alloc(nbytes)
                self.own.append(ptr)
                handles[k] = h
        except Exception:
            ok = 0
        gathered = [None] * d.world
        d.dist.all_gather_object(gathered, (ok, handles), group=d.group)
        ok = min(g[0] for g in gathered)
        if ok:
            try:
                for k in range(2):
                    row = []
                    for r in range(d.world):
                        if r == d.rank:
                            row.append(self.own[k])
                        else:
                            q = capi.peer_open(gathered[r][1][k])
                            self.opened.append(q)
                            row.append(q)
                    self.peers.append(row)
            except Exception:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        d.dist.all_reduce(flag, op=d.dist.ReduceOp.MIN, group=d.group)
        self.ok = bool(int(flag.item()))
        self.turn = 0
        self.token = torch.zeros(1, dtype=torch.int32, device=device)
        self.grids = [torch.as_tensor(_CudaView(p, shape), device=device) for p in self.own] if self.ok else []
        if not self.ok:
            self.close()

    def close(self):
        for q in self.opened:
            try:
                self.capi.peer_close(q)
            except Exception:
                pass
        for p in self.own:
            try:
                self.capi.peer_free(p)
            except Exception:
                pass
        self.opened, self.own, self.grids = [], [], []


class ShardedSH:
    """Batch-, latitude- and order-sharded transforms on top of an engine with the ``SHEngine`` array interface.

    peer=True (CUDA tensors, one box): the latitude-sharded synthesis exchanges its rows through peer memory (NVLink stores into
    every rank's grid, `_PeerGrids`) instead of ncclAllGather + row scatter.  The grid it returns is then one of two exchange
    buffers of this object and stays valid until the second next call; ``.clone()`` it to keep it longer."""

    def __init__(self, engine, group=None, peer=False):
        self.engine = engine
        self.d = _Dist(group)
        self._slots = {}
        self._peer = {} if peer else None
        self.last_exchange_bytes = 0      # bytes this rank contributed to + received from the last collective
        self.last_exchange = "none"

    def _peer_grids(self, shape, device):
        """Collective on first use per shape (IPC handle exchange); None if peer memory is unavailable on any rank."""
        if self._peer is None or self.d.world == 1 or self.d.world > 16:
            return None
        key = (tuple(shape), str(device))
        if key not in self._peer:
            import torch
            from . import capi
            for old in self._peer.values():
                if old is not None:
                    old.close()
            self._peer.clear()
            pg = _PeerGrids(self.d, capi, torch, device, shape)
            self._peer[key] = pg if pg.ok else None
        return self._peer[key]

    # ---- batch sharding: no collective -------------------------------------------------------------
    def local_batch(self, batch):
        return shard_bounds(batch, self.d.world, self.d.rank)

    def synthesis_batch(self, coef_local, **kw):
        """Every rank passes ITS slice [B_local, nsh]; returns its slice of the grids.  No communication."""
        return self.engine.synthesis(coef_local, **kw)

    def analysis_batch(self, grid_local, nmax, lon, lat, **kw):
        return self.engine.analysis(grid_local, nmax, lon, lat, **kw)

    # ---- latitude sharding of one synthesis ------------------------------------------------------------
    def _row_slots(self, lat, device):
        """(rows of every rank, rmax, row_of_slot int32 [world*rmax] on ``device`` or numpy)"""
        W = self.d.world
        key = (lat.tobytes(), W, str(device))
        if key not in self._slots:
            if len(self._slots) > 8:
                self._slots.clear()
            allrows = [lat_shard_indices(lat, W, r) for r in range(W)]
            rmax = max(len(r) for r in allrows)
            ros = np.full(W * rmax, -1, dtype=np.int32)
            for r, rows in enumerate(allrows):
                ros[r * rmax:r * rmax + len(rows)] = rows
            if device is not None:
                import torch
                ros = torch.from_numpy(ros).to(device)
            self._slots[key] = (allrows, rmax, ros)
        return self._slots[key]

    def synthesis_latsharded(self, coef, lon, lat, gather=True, **kw):
        """coef replicated on every rank ([B, nsh]); each rank synthesises its latitude rows.  Returns the full grid
        [B, nlat, nlon] on every rank when ``gather`` (one all-gather + one row scatter), else (rows, local grid)."""
        import torch
        lat = np.asarray(lat, dtype=np.float64)
        W, rank = self.d.world, self.d.rank
        is_t = hasattr(coef, "is_cuda")
        allrows, rmax, ros = self._row_slots(lat, coef.device if is_t else None)
        rows = allrows[rank]
        if not gather:
            return rows, self.engine.synthesis(coef, lon=lon, lat=lat[rows], **kw)
        B = coef.shape[0] if coef.ndim == 2 else 1
        K, L = len(lon), len(lat)
        pg = self._peer_grids((B, L, K), coef.device) if is_t and hasattr(self.engine, "scatter_rows") else None
        if pg is not None:
            from . import capi
            local = self.engine.synthesis(coef, lon=lon, lat=lat[rows], **kw)             # [B, own rows, K]
            rows_dev = ros[rank * rmax:rank * rmax + len(rows)]
            k = pg.turn
            pg.turn ^= 1
            stream = torch.cuda.current_stream(coef.device).cuda_stream
            capi.push_rows_dev(B, W, len(rows), K, local.data_ptr(), local.stride(0), rows_dev.data_ptr(), pg.peers[k], L * K, K, stream)
            self.engine.launches += 1
            # barrier: every rank's push kernel has completed (stream order) before any rank reads its grid
            self.d.dist.all_reduce(pg.token, group=self.d.group)
            self.last_exchange_bytes = B * len(rows) * K * 8 * W
            self.last_exchange = "peer stores (shb200_push_rows) + 4-byte all-reduce as barrier"
            return pg.grids[k]
        if is_t and hasattr(self.engine, "scatter_rows"):
            self.last_exchange = "ncclAllGather + shb200_scatter_rows"
            buf = torch.empty((W, B, rmax, K), dtype=torch.float64, device=coef.device)
            self.engine.synthesis(coef, lon=lon, lat=lat[rows], out=buf[rank][:, :len(rows), :], **kw)
            self.d.all_gather_inplace(buf)
            self.last_exchange_bytes = buf.numel() * 8
            return self.engine.scatter_rows(buf, ros, L)
        local = np.asarray(self.engine.synthesis(coef, lon=lon, lat=lat[rows], **kw))
        buf = torch.zeros((W, B, rmax, K), dtype=torch.float64)
        buf[rank][:, :len(rows), :] = torch.from_numpy(np.ascontiguousarray(local).reshape(B, len(rows), K))
        self.d.all_gather_inplace(buf)
        self.last_exchange_bytes = buf.numel() * 8
        full = np.empty((B, L, K))
        g = buf.numpy()
        for r, rr in enumerate(allrows):
            full[:, rr, :] = g[r][:, :len(rr), :]
        return full

    def analysis_latsharded(self, grid, nmax, lon, lat, lat_weights, weight, **kw):
        """grid [B, nlat, nlon] replicated (or at least the rank's rows valid); every rank integrates its latitude rows
        with the given per-latitude quadrature weights; ONE all_reduce(SUM) of [B, (nmax+1)^2] combines the partial sums."""
        import torch
        lat = np.asarray(lat, dtype=np.float64)
        lw = np.asarray(lat_weights, dtype=np.float64)
        rows = lat_shard_indices(lat, self.d.world, self.d.rank)
        if hasattr(grid, "is_cuda"):
            sub = grid[:, torch.as_tensor(rows, device=grid.device), :]
        else:
            sub = np.ascontiguousarray(np.asarray(grid)[:, rows, :])
        part = self.engine.analysis(sub, nmax, lon, lat[rows], quadrature="weights", lat_weights=lw[rows], weight=weight, **kw)
        if self.d.world == 1:
            return part
        is_t = hasattr(part, "is_cuda")
        pt = part if is_t else torch.from_numpy(np.ascontiguousarray(part))
        self.d.dist.all_reduce(pt, op=self.d.dist.ReduceOp.SUM, group=self.d.group)
        self.last_exchange_bytes = pt.numel() * 8
        return pt if is_t else pt.numpy()

    # ---- m-order sharding of one analysis ----------------------------------------------------------------
    def analysis_msharded(self, grid, nmax, lon, lat, **kw):
        """grid [B, nlat, nlon] replicated on every rank; rank r integrates the orders m = r (mod world) and emits only
        those (compact m-major shard); ONE all-gather over NVLink; one unpack kernel writes the nm vector.  Bit-identical
        to the unsharded analysis."""
        import torch
        W, rank = self.d.world, self.d.rank
        smax = shard_size(nmax, 0, W)                 # rank 0 owns m = 0: the largest shard
        is_t = hasattr(grid, "is_cuda")
        B = grid.shape[0]
        if is_t and hasattr(self.engine, "unpack_shards"):
            buf = torch.empty((W, B, smax), dtype=torch.float64, device=grid.device)
            self.engine.analysis_packed(grid, nmax, lon, lat, (rank, W), buf[rank], **kw)
            self.d.all_gather_inplace(buf)
            self.last_exchange_bytes = buf.numel() * 8
            return self.engine.unpack_shards(buf, nmax)
        # numpy engines (CPU tests with the oracle): same exchange format through the numpy definitions above
        part = np.asarray(self.engine.analysis(grid, nmax, lon, lat, order_shard=(rank, W), **kw))
        buf = torch.zeros((W, B, smax), dtype=torch.float64)
        mine = pack_orders_np(part, nmax, rank, W)
        buf[rank][:, :mine.shape[1]] = torch.from_numpy(mine)
        self.d.all_gather_inplace(buf)
        self.last_exchange_bytes = buf.numel() * 8
        return unpack_shards_np(buf.numpy(), nmax)
