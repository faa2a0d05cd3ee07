"""Multi-GPU planner: one process per GPU (torch.distributed, NCCL over NVLink on the B200 box; gloo in CPU tests).

The reference has no distributed path at all (single process, OpenMP over latitude rows:
synthesis.pyx:179, analysis.pyx:128).  The path shards three ways:

* batch sharding  -- fields are independent (the dgemv/dger columns never mix: synthesis.pyx:183,
  analysis.pyx:136): every rank transforms its own contiguous slice of the batch, NO collective.
  This is what bench.py scales (weak scaling).
* latitude sharding of a single (few-field) transform -- rows are independent in synthesis
  (prange(nlat)); mirror latitude pairs stay on one rank so the E/O symmetry survives.  Synthesis needs no
  reduction (optional all_gather of the grid rows); analysis produces per-rank partial quadrature sums that
  one all_reduce(SUM) of the coefficient vector combines -- the only real exchange step of the path.

* m-order sharding of a single analysis (what BASELINE's config 4 names) -- every rank holds the grid, runs the cheap
  longitude stage for all rows and integrates only the orders m = rank (mod world) (interleaved because the work per
  order is proportional to nmax-m+1); all other coefficients come out as exact zeros, so one all_reduce(SUM) -- a gather
  in effect, 37 MB at nmax 2160 -- assembles the full vector bit-identically to the unsharded transform.

The engine is injectable so the planner logic is testable on CPU (world_size 2, gloo) with the oracle.
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
    lo, hi = shard_bounds(len(units), world, rank)
    idx = sorted(k for u in units[lo:hi] for k in u)
    return np.asarray(idx, dtype=np.int64)


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


class ShardedSH:
    """Batch- and latitude-sharded transforms on top of an engine with the ``SHEngine`` array interface."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.d = _Dist(group)

    # ---- batch sharding: no collective -------------------------------------------------------------
    def local_batch(self, batch):
        return shard_bounds(batch, self.d.world, self.d.rank)

    def synthesis_batch(self, coef_local, **kw):
        """Every rank passes ITS slice [B_local, nsh]; returns its slice of the grids.  No communication."""
        return self.engine.synthesis(coef_local, **kw)

    def analysis_batch(self, grid_local, nmax, lon, lat, **kw):
        return self.engine.analysis(grid_local, nmax, lon, lat, **kw)

    # ---- latitude sharding of one transform ----------------------------------------------------------
    def synthesis_latsharded(self, coef, lon, lat, gather=True, **kw):
        """coef replicated on every rank ([B, nsh]); each rank synthesises its latitude rows.  Returns the full grid
        [B, nlat, nlon] on every rank when ``gather`` (one all_gather of row blocks), else (rows, local grid)."""
        import torch
        lat = np.asarray(lat, dtype=np.float64)
        rows = lat_shard_indices(lat, self.d.world, self.d.rank)
        local = self.engine.synthesis(coef, lon=lon, lat=lat[rows], **kw)           # [B, nrows_local, nlon]
        if not gather or self.d.world == 1:
            return (rows, local) if not gather else _scatter_rows(local, [rows], len(lat))
        is_t = hasattr(local, "is_cuda")
        loc_t = local if is_t else torch.from_numpy(np.ascontiguousarray(local))
        counts = [len(lat_shard_indices(lat, self.d.world, r)) for r in range(self.d.world)]
        B, K = loc_t.shape[0], loc_t.shape[2]
        bufs = [torch.empty((B, c, K), dtype=loc_t.dtype, device=loc_t.device) for c in counts]
        self.d.dist.all_gather(bufs, loc_t.contiguous(), group=self.d.group)
        allrows = [lat_shard_indices(lat, self.d.world, r) for r in range(self.d.world)]
        full = _scatter_rows(bufs, allrows, len(lat))
        return full if is_t else full.numpy()

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
        return pt if is_t else pt.numpy()


    def analysis_msharded(self, grid, nmax, lon, lat, **kw):
        """grid [B, nlat, nlon] replicated on every rank; rank r integrates the orders m = r (mod world)
        (shb200_plan_set_order_shard) and one all_reduce(SUM) over NVLink gathers the disjoint shards."""
        import torch
        part = self.engine.analysis(grid, nmax, lon, lat, order_shard=(self.d.rank, self.d.world), **kw)
        if self.d.world == 1:
            return part
        is_t = hasattr(part, "is_cuda")
        pt = part if is_t else torch.from_numpy(np.ascontiguousarray(part))
        self.d.dist.all_reduce(pt, op=self.d.dist.ReduceOp.SUM, group=self.d.group)
        return pt if is_t else pt.numpy()


def _scatter_rows(blocks, rows_per_rank, nlat):
    import torch
    if not isinstance(blocks, (list, tuple)):
        blocks = [blocks]
    b0 = blocks[0]
    if hasattr(b0, "is_cuda"):
        full = torch.empty((b0.shape[0], nlat, b0.shape[2]), dtype=b0.dtype, device=b0.device)
        for blk, rows in zip(blocks, rows_per_rank):
            full[:, torch.as_tensor(rows, device=b0.device), :] = blk
        return full
    full = np.empty((b0.shape[0], nlat, b0.shape[2]), dtype=b0.dtype)
    for blk, rows in zip(blocks, rows_per_rank):
        full[:, rows, :] = blk
    return full
