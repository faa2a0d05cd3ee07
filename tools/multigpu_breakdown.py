#!/usr/bin/env python
"""Where the time of a sharded cfg4 transform goes on each rank: CUDA events between the pieces of
ShardedSH.synthesis_latsharded / analysis_msharded (local transform, all-gather, scatter/unpack) plus host-side call time.
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multigpu_breakdown.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import shxarray_b200 as sb
from shxarray_b200 import capi, synthetic as S
from shxarray_b200.distributed import ShardedSH, shard_size

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nmax = int(os.environ.get("NMAX", "2160"))
d = 180.0 / nmax
h = np.arange(d / 2, 90, d)
lat = np.concatenate([-h[::-1], h])
lon = np.arange(0, 360, d)
w = (d * np.pi / 180) ** 2 / (4 * np.pi)
lw = np.cos(lat * (np.pi / 180))
eng = sb.SHEngine(device=local)
sh = ShardedSH(eng)
c = torch.from_numpy(S.random_coefficients(nmax, 1, kaula=True)).cuda()
L, K = len(lat), len(lon)
allrows, rmax, ros = sh._row_slots(lat, c.device)
rows = allrows[rank]
akw = dict(quadrature="weights", lat_weights=lw, weight=w)


def ev():
    return torch.cuda.Event(enable_timing=True)


def run(pieces, reps=5):
    """pieces: list of (name, fn); returns mean ms per piece (device) and host ms for the whole sequence"""
    for _, fn in pieces:
        fn()
    torch.cuda.synchronize()
    acc = {n: 0.0 for n, _ in pieces}
    host = 0.0
    for _ in range(reps):
        dist.barrier()
        torch.cuda.synchronize()
        marks = [ev()]
        marks[0].record()
        t0 = time.perf_counter()
        for n, fn in pieces:
            fn()
            e = ev()
            e.record()
            marks.append(e)
        host += (time.perf_counter() - t0) * 1e3 / reps
        torch.cuda.synchronize()
        for i, (n, _) in enumerate(pieces):
            acc[n] += marks[i].elapsed_time(marks[i + 1]) / reps
    acc["host_issue_ms"] = host
    return acc


buf_s = torch.empty((world, 1, rmax, K), dtype=torch.float64, device="cuda")
full = torch.empty((1, L, K), dtype=torch.float64, device="cuda")
syn = run([("local_synthesis", lambda: eng.synthesis(c, lon=lon, lat=lat[rows], nmax=nmax, out=buf_s[rank][:, :len(rows), :])),
           ("all_gather", lambda: sh.d.all_gather_inplace(buf_s)),
           ("scatter_rows", lambda: eng.scatter_rows(buf_s, ros, L, out=full))])
smax = shard_size(nmax, 0, world)
buf_a = torch.empty((world, 1, smax), dtype=torch.float64, device="cuda")
outc = torch.empty((1, (nmax + 1) ** 2), dtype=torch.float64, device="cuda")
ana = run([("local_analysis_packed", lambda: eng.analysis_packed(full, nmax, lon, lat, (rank, world), buf_a[rank], **akw)),
           ("all_gather", lambda: sh.d.all_gather_inplace(buf_a)),
           ("unpack_shards", lambda: eng.unpack_shards(buf_a, nmax, out=outc))])
# stage times inside the local transforms (library events)
ps = eng.get_plan(nmax, lat[rows], lon, 1, flags=capi.FLAG_TIMING)
tmp = torch.empty((1, len(rows), K), dtype=torch.float64, device="cuda")
for _ in range(3):
    ps.synthesis_dev(1, c.data_ptr(), c.stride(0), 1, tmp.data_ptr(), tmp.stride(0), K, 1, torch.cuda.current_stream().cuda_stream)
st_s = ps.stage_times()
res = {"rank": rank, "world": world, "synthesis": syn, "analysis": ana, "local_synthesis_stages": st_s, "rows": int(len(rows)),
       "kernels_syn": ps.last_kernels()}
allres = [None] * world
dist.all_gather_object(allres, res)
if rank == 0:
    print(json.dumps({"nmax": nmax, "ranks": allres}))
dist.barrier()
dist.destroy_process_group()
