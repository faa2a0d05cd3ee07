"""CPU tests of the multi-GPU planner: world_size-2 gloo process group, the transform itself injected from the oracle
(the CUDA engine cannot run here and has no CPU fallback)."""
import os
import socket

import numpy as np
import pytest

from shxarray_b200.distributed import ShardedSH, lat_shard_indices, shard_bounds


def test_shard_bounds_cover_without_overlap():
    for total in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_bounds(total, world, r)
                assert 0 <= lo <= hi <= total
                seen += list(range(lo, hi))
            assert seen == list(range(total))
            sizes = [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_lat_shards_keep_mirror_pairs_together():
    lat = np.arange(-89.5, 90, 1.0)
    lat = np.concatenate([lat, [0.0, 33.3]])      # an equator row and an unpaired row
    for world in (1, 2, 4, 8):
        allidx = []
        for r in range(world):
            idx = lat_shard_indices(lat, world, r)
            vals = set(lat[idx])
            for v in lat[idx]:
                if v != 0.0 and v != 33.3:
                    assert -v in vals
            allidx += idx.tolist()
        assert sorted(allidx) == list(range(len(lat)))


class OracleEngine:
    """SHEngine-shaped adapter around the CPU oracle (test infrastructure only)."""

    def synthesis(self, coef, n=None, m=None, lon=None, lat=None, nmax=None, **kw):
        from oracle import oracle as O
        if n is None:
            n, m = O.nm_index(nmax)
        out = O.synthesis(np.asarray(coef), n, m, lon, lat, kind="port", nthreads=2)
        return np.ascontiguousarray(np.moveaxis(out, 2, 0))

    def analysis(self, grid, nmax, lon, lat, quadrature="weights", lat_weights=None, weight=None, order_shard=None, **kw):
        from oracle import oracle as O
        lat = np.asarray(lat)
        if quadrature == "shlib":
            out = O.analysis(np.asarray(grid), nmax, lon, lat, kind="port", nthreads=2)
        else:
            # the oracle applies weight*cos(lat); fold the requested per-latitude weight into the data instead
            scale = np.asarray(lat_weights) / np.cos(lat * (np.pi / 180))
            out = O.analysis(np.asarray(grid) * scale[None, :, None], nmax, lon, lat, kind="port", weight=weight, nthreads=2)
        if order_shard is not None:          # emulate shb200_plan_set_order_shard: other ranks' orders are exact zeros
            rank, world = order_shard
            _, m = O.nm_index(nmax)
            out = out * ((np.abs(m) % world) == rank)[None, :]
        return out


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from shxarray_b200 import synthetic as S
        nmax, B = 16, 5
        lon, lat = O.shlib_grid(nmax)
        n, m = O.nm_index(nmax)
        coef = S.random_coefficients(nmax, B, kaula=False)
        sh = ShardedSH(OracleEngine())
        # batch sharding: slices are disjoint and cover the batch; no collective involved
        lo, hi = sh.local_batch(B)
        mine = sh.synthesis_batch(coef[lo:hi], n=n, m=m, lon=lon, lat=lat)
        # latitude sharding + all_gather
        full = sh.synthesis_latsharded(coef, lon, lat, n=n, m=m)
        ref = np.moveaxis(O.synthesis(coef, n, m, lon, lat, kind="port", nthreads=2), 2, 0)
        ok_syn = np.abs(full - ref).max() <= 1e-13 * np.abs(ref).max()
        ok_batch = np.array_equal(mine, ref[lo:hi])
        # latitude-sharded analysis + all_reduce
        w = O.analysis_weight(lon, lat)
        lw = np.cos(lat * (np.pi / 180))
        part = sh.analysis_latsharded(ref, nmax, lon, lat, lat_weights=lw, weight=w)
        aref = O.analysis(ref, nmax, lon, lat, kind="port", nthreads=2)
        ok_ana = np.abs(part - aref).max() <= 1e-12 * np.abs(aref).max()
        # m-order-sharded analysis + all_reduce (disjoint shards: the sum is a gather, exact)
        msh = sh.analysis_msharded(ref, nmax, lon, lat, quadrature="shlib")
        ok_ana = ok_ana and np.array_equal(msh, aref)
        q.put((rank, lo, hi, bool(ok_batch), bool(ok_syn), bool(ok_ana)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharded_transforms():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 5)]
    for r in res:
        assert r[3] and r[4] and r[5], r


def test_work_bounds_balance_contiguous_runs():
    """Latitude bands of a sharded transform are contiguous runs of (nearly) equal total weight (distributed.work_bounds): they
    cover every unit once, no rank is empty, equal weights give shard_bounds, and on the 5' grid the polar rank of 8 owns more
    rows than the equatorial one while the modelled work per rank stays within 2 % of the mean."""
    from shxarray_b200.distributed import work_bounds, _row_work
    rng = np.random.default_rng(4)
    for n, world in ((10, 3), (37, 8), (1080, 8), (5, 5), (3, 8)):
        w = list(rng.uniform(0.2, 1.0, n))
        cuts = [work_bounds(w, world, r) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
        if n >= world:
            assert all(hi > lo for lo, hi in cuts)
        assert [work_bounds([1.0] * n, world, r) for r in range(world)] == [shard_bounds(n, world, r) for r in range(world)]
    d = 1.0 / 12
    h = np.arange(d / 2, 90, d)
    lat = np.concatenate([-h[::-1], h])
    sizes = [len(lat_shard_indices(lat, 8, r)) for r in range(8)]
    work = [sum(_row_work(lat[i], 8) for i in lat_shard_indices(lat, 8, r)) for r in range(8)]
    assert sum(sizes) == len(lat) and sizes[0] > sizes[-1] * 1.5
    assert max(work) / (sum(work) / 8) < 1.02 and min(work) / (sum(work) / 8) > 0.98
