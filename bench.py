#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: SH synthesis+analysis fields/s (FP64) at nmax=720, batch 64, on N B200s.

One step = one pass of the hot path over one batch: 64 random Stokes fields synthesised to the
724 x 1536 Gauss grid AND analysed back (a "field" = one round trip).  Multi-GPU = batch sharding
(SURVEY.md §8e.1): every rank owns its own 64 fields, no data-path collective, weak scaling.

  value    : device-resident round trips/s (inputs already in HBM), whole job, max-over-ranks time
  e2e      : the same through the calls the plugin makes -- SHEngine.synthesis / .analysis with PLAIN (pageable) numpy
             coefficients in, numpy out, H2D + D2H inside the timed region; beside it the all-pinned C-ABI variant, the
             fused pipelined round trip (shb200_roundtrip_host) and the PCIe copy rates measured in the same run
  roofline : dominant kernel (Legendre stage, FP64 tensor pipe), algorithmic flops / CUDA-event time
  cpu_baseline : the shlib arithmetic (oracle/_ref, else the C port) on the host cores, bounded sample

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NMAX, NLAT, NLON, BATCH = 720, 724, 1536, 64
METRIC = "SH synthesis+analysis fields/sec (FP64) at nmax=720 batch 64"
UNIT = "fields/s"


def workload_config(ngpu):
    return {"workload": f"cfg3: nmax={NMAX} Gauss grid {NLAT}x{NLON}, batch {BATCH} per GPU, synthesis+analysis round trip",
            "nmax": NMAX, "grid": [NLAT, NLON], "batch_per_gpu": BATCH, "parallelism": f"batch-shard x{ngpu} (no collective)",
            "l2": "inputs larger than L2 (coefficients 266 MB + grid 570 MB per step)"}


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smmax, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                try:
                    sm.append(float(f[1]))
                    smmax = float(f[2])
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smmax, "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed ncu --set full capture."""
    try:
        for name in ("r02_traffic.json", "r01_traffic.json"):
            path = os.path.join(ROOT, "profiles", name)
            if os.path.exists(path):
                d = json.load(open(path))["kernels"][kernel]
                return d["dram_bytes_read"] + d["dram_bytes_write"]
        return None
    except Exception:
        return None


def fp64_peak_tflops():
    """Measured FP64 tensor-pipe (DMMA m8n8k4) peak on this pool's B200 (tools/fp64_peaks.cu -> profiles/)."""
    path = os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")
    try:
        d = json.load(open(path))
        return float(d["dmma884_tflops_occ8"]), "measured FP64 DMMA peak, profiles/fp64_peaks_r01.json (MEASURED_PEAKS.json has no FP64 entry)"
    except Exception:
        return 37.2, "fallback: nominal 148 SM x 64 FMA/clk x 1.965 GHz"


# --------------------------------------------------------------------------------------------------
def cpu_sample(steps, warmup, per_step_scale=1.0):
    """Time the shlib arithmetic on a bounded sample of cfg3 and extrapolate linearly in the number of grid
    points (every point costs the same in synthesis.pyx:179-183 / analysis.pyx:128-139)."""
    from oracle import oracle as O
    from shxarray_b200 import synthetic as S
    kind = O.best_kind()
    try:
        nthr = len(os.sched_getaffinity(0))      # torchrun exports OMP_NUM_THREADS=1: ask for every usable host thread explicitly
    except AttributeError:
        nthr = os.cpu_count() or 1
    lon, lat, _ = S.gauss_grid(NLAT, NLON, nodes="numpy")     # the CPU arm never loads libshb200.so
    n, m = O.nm_index(NMAX)
    c = S.random_coefficients(NMAX, BATCH, kaula=True)
    rows = np.linspace(0, NLAT - 1, max(nthr, 2)).round().astype(int)       # >= 1 latitude row per thread (prange over rows)
    nl_syn = max(2, int(round(16 * per_step_scale)))
    nl_ana = max(1, int(round(96 * per_step_scale / len(rows))) or 1)
    cols_s = np.linspace(0, NLON - 1, nl_syn).round().astype(int)
    cols_a = np.linspace(0, NLON - 1, nl_ana).round().astype(int)
    rng = np.random.default_rng(1)
    f = rng.standard_normal((BATCH, len(rows), len(cols_a)))
    w = 1.0 / (2.0 * NLON)
    vals, t_syn, t_ana = [], [], []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        O.synthesis(c, n, m, lon[cols_s], lat[rows], grid=True, kind=kind, nthreads=nthr)
        t1 = time.perf_counter()
        O.analysis(f, NMAX, lon[cols_a], lat[rows], kind=kind, weight=w, nthreads=nthr)
        t2 = time.perf_counter()
        if it >= warmup:
            full_s = (t1 - t0) * (NLAT * NLON) / (len(rows) * len(cols_s))
            full_a = (t2 - t1) * (NLAT * NLON) / (len(rows) * len(cols_a))
            t_syn.append(full_s)
            t_ana.append(full_a)
            vals.append(BATCH / (full_s + full_a))
    sample = (f"shlib loops ({'reference headers + scipy OpenBLAS' if kind == 'ref' else 'C port'}) on {len(rows)} of {NLAT} latitude rows x "
              f"{len(cols_s)} (synthesis) / {len(cols_a)} (analysis) of {NLON} longitudes, batch {BATCH}, extrapolated linearly in grid points")
    return dict(value=float(np.mean(vals)), unit=UNIT, cores=int(nthr), kind="reference" if kind == "ref" else "port", sample=sample,
                synthesis_s_per_batch=float(np.mean(t_syn)), analysis_s_per_batch=float(np.mean(t_ana))), float(np.mean(t_syn) + np.mean(t_ana))


def run_reference(args, rank):
    if rank != 0:
        return
    base, tstep = cpu_sample(args.steps, args.warmup, per_step_scale=0.5)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tstep * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
def cfg4_sharded(torch, dist, sb, S, rank, world, local, nmax=2160, reps=5):
    """BASELINE config 4 on the N GPUs of the run: ONE nmax-2160 field on the 5' grid (2160 x 4320), synthesis sharded by
    latitude rows, analysis sharded by orders m (distributed.ShardedSH: one ncclAllGather each), against the same
    transform on one GPU.  Device time (CUDA events), max over ranks.  Extra keys of the N>1 bench line only.
    Synthesis exchanges its rows through peer memory (one kernel of NVLink stores into every rank's grid + a 4-byte all-reduce
    as barrier) where CUDA IPC works, else one ncclAllGather + a row scatter; analysis: one ncclAllGather of packed shards."""
    from shxarray_b200.distributed import ShardedSH
    d = 180.0 / nmax
    h = np.arange(d / 2, 90, d)
    lat = np.concatenate([-h[::-1], h])
    lon = np.arange(0, 360, d)
    w = (d * np.pi / 180) ** 2 / (4 * np.pi)
    lw = np.cos(lat * (np.pi / 180))
    eng = sb.SHEngine(device=local)
    sh = ShardedSH(eng, peer=not os.environ.get("SHB200_NO_PEER"))     # rows exchanged through peer memory (NVLink stores) when the box allows it
    c = torch.from_numpy(S.random_coefficients(nmax, 1, kaula=True)).cuda()

    def timed(fn):
        out = fn()
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t.item())

    akw = dict(quadrature="weights", lat_weights=lw, weight=w)
    f_sh, t_syn = timed(lambda: sh.synthesis_latsharded(c, lon, lat, nmax=nmax))
    bytes_syn, how_syn = sh.last_exchange_bytes, sh.last_exchange
    f_sh = f_sh.clone()         # the peer-memory path returns one of its two exchange buffers
    a_sh, t_ana = timed(lambda: sh.analysis_msharded(f_sh, nmax, lon, lat, **akw))
    bytes_ana = sh.last_exchange_bytes
    f_1, t_syn1 = timed(lambda: eng.synthesis(c, lon=lon, lat=lat, nmax=nmax))
    a_1, t_ana1 = timed(lambda: eng.analysis(f_1, nmax, lon, lat, **akw))
    a_on_sharded_grid = eng.analysis(f_sh, nmax, lon, lat, **akw)
    ok = torch.tensor([float(torch.equal(a_sh, a_on_sharded_grid)), -float(((f_sh - f_1).abs().max() / f_1.abs().max()).item())],
                      device="cuda", dtype=torch.float64)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return {"workload": f"cfg4: ONE field nmax={nmax}, grid {len(lat)}x{len(lon)} (5'), sharded over {world} GPUs",
            "synthesis_latsharded_ms": t_syn, "analysis_msharded_ms": t_ana, "synthesis_1gpu_ms": t_syn1, "analysis_1gpu_ms": t_ana1,
            "speedup_synthesis": t_syn1 / t_syn, "speedup_analysis": t_ana1 / t_ana,
            "exchange_bytes_synthesis": int(bytes_syn), "exchange_synthesis": how_syn, "allgather_bytes_analysis": int(bytes_ana),
            "analysis_bit_identical_to_1gpu": bool(ok[0].item() == 1.0), "synthesis_max_rel_diff_vs_1gpu": float(-ok[1].item()),
            "collectives_per_transform": 1}


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flags", type=int, default=0, help="extra SHB200_FLAG_* bits (e.g. 4 = force DFMA kernels)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from shxarray_b200 import capi, synthetic as S
    import shxarray_b200 as sb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    # host locality before the library creates its copy threads / page-locked buffers: stay on the GPU's NUMA node (restored
    # before the CPU baseline runs, which uses every core it is allowed)
    from shxarray_b200 import affinity
    cpus_before = os.sched_getaffinity(0)
    numa = affinity.bind_to_device_numa(local) if not os.environ.get("SHB200_NO_NUMA_BIND") else {"bound": False, "reason": "SHB200_NO_NUMA_BIND"}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- plans and inputs (device resident) ------------------------------------------------------
    grid = sb.lonlat_grid(NMAX, gtype="gauss")
    lon, lat, latw = grid["lon"], grid["lat"], grid["lat_weights"]
    assert len(lat) == NLAT and len(lon) == NLON
    flags = capi.FLAG_TIMING | args.flags
    plan = capi.Plan(NMAX, lat, lon, max_batch=BATCH, quadrature=capi.QUAD_LATWEIGHTS, weight=1.0 / (2.0 * NLON), lat_weights=latw, flags=flags, device=local)
    nsh = plan.nsh_full
    coef_h = torch.from_numpy(S.random_coefficients(NMAX, BATCH, kaula=True, seed=S.SEED + rank)).pin_memory()
    coef = coef_h.cuda()
    gridbuf = torch.empty((BATCH, NLAT, NLON), dtype=torch.float64, device="cuda")
    coef_out = torch.empty((BATCH, nsh), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        plan.synthesis_dev(BATCH, coef.data_ptr(), nsh, 1, gridbuf.data_ptr(), NLAT * NLON, NLON, 1, stream)
        plan.analysis_dev(BATCH, gridbuf.data_ptr(), NLAT * NLON, NLON, 1, coef_out.data_ptr(), nsh, 1, stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    rt_err = float(((coef_out - coef).abs().amax(dim=1) / coef.abs().amax(dim=1)).max().item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    value = world * BATCH * args.steps / (ms_total * 1e-3)

    # ---- per-kernel durations (CUDA events recorded by the library on the launching stream) -----
    acc = {}
    for _ in range(args.steps):
        step()
        st = plan.stage_times()
        for k, v in st.items():
            acc[k] = acc.get(k, 0.0) + v / args.steps
    kernels = plan.last_kernels()          # of the last device-resident analysis call
    L_rows = plan.info.nrows
    paired = L_rows < NLAT
    flops_full = (1.0 if paired else 2.0) * nsh * NLAT * BATCH         # F_sym / F_alg Legendre term x batch
    # executed work only (never credit skipped work): the table kernels do not visit (n, m, row) triples with
    # |P_nm| < 1e-24 (polar skipping); the library reports the fraction each direction visits
    visited = {"syn_legendre": plan.info.legendre_visited_synthesis, "ana_legendre": plan.info.legendre_visited_analysis}
    dom = "syn_legendre" if acc["syn_legendre"] >= acc["ana_legendre"] else "ana_legendre"
    flops_leg = flops_full * visited[dom]
    peak, peak_src = fp64_peak_tflops()
    achieved = flops_leg / (acc[dom] * 1e-3) * 1e-12
    kname = {"syn_legendre": "k_leg_syn_tab", "ana_legendre": "k_leg_ana_tab"}[dom] if plan.info.dmma == 2 else None
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": ncu_traffic(kname) if kname else None, "traffic_unit": "bytes per launch (dram read+write, ncu)", "flops_per_launch": flops_leg, "ms_per_launch": acc[dom], "peak_source": peak_src,
                "symmetry": "mirror-latitude pairing used: F_sym = nsh*L per field" if paired else "no pairing: F_alg = 2*nsh*L per field",
                "executed_fraction": visited, "flops_per_launch_unskipped": flops_full,
                "executed_note": "flops_per_launch counts EXECUTED work: F_sym x the fraction of (n,m,row) triples the kernel visits (|P_nm| < 1e-24 towards the poles is skipped)",
                "stages_ms": acc}

    # ---- FP64 peak re-measured on this box: cuBLAS DGEMM 8192^3 through torch (a library GEMM as yardstick only) --------
    def dgemm_tflops():
        n = 8192
        a = torch.randn((n, n), dtype=torch.float64, device="cuda")
        b = torch.randn((n, n), dtype=torch.float64, device="cuda")
        torch.matmul(a, b)
        t0_, t1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0_.record()
        for _ in range(3):
            torch.matmul(a, b)
        t1_.record()
        torch.cuda.synchronize()
        return 3 * 2.0 * n ** 3 / (t0_.elapsed_time(t1_) * 1e-3) * 1e-12

    dgemm = dgemm_tflops() if not os.environ.get("BENCH_DEBUG_SKIP_DGEMM") else 1.0
    flops_step = flops_full * (visited["syn_legendre"] + visited["ana_legendre"]) + 2.0 * 2.5 * NLAT * NLON * np.log2(NLON) * BATCH     # both directions: executed Legendre work + FFT terms
    roofline["fp64_peak_check"] = {"cublas_dgemm_8192_tflops": dgemm, "frac_vs_dgemm": achieved / dgemm,
                                   "note": "cuBLAS DGEMM measured in this run; `peak` is the DMMA issue-rate microbenchmark (tools/fp64_peaks.cu)"}
    roofline["whole_step"] = {"flops_per_step": flops_step, "achieved_tflops": flops_step / (ms_total / args.steps * 1e-3) * 1e-12,
                              "frac": flops_step / (ms_total / args.steps * 1e-3) * 1e-12 / peak}
    torch.cuda.empty_cache()

    # ---- PCIe copy rates of this box, measured in this run: pinned H2D, D2H and both at once (all ranks concurrently) ----
    def pcie_rates():
        nbytes = 512 << 20
        hp = [torch.empty(nbytes // 8, dtype=torch.float64).pin_memory() for _ in range(2)]
        dv = [torch.empty(nbytes // 8, dtype=torch.float64, device="cuda") for _ in range(2)]
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def run(h2d, d2h, reps=4):
            barrier()
            torch.cuda.synchronize()
            t0_ = time.perf_counter()
            for _ in range(reps):
                if h2d:
                    with torch.cuda.stream(s1):
                        dv[0].copy_(hp[0], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        hp[1].copy_(dv[1], non_blocking=True)
            torch.cuda.synchronize()
            return reps * nbytes / (time.perf_counter() - t0_) * 1e-9

        run(True, True, 1)
        out = {"h2d_gbs": run(True, False), "d2h_gbs": run(False, True), "duplex_each_gbs": run(True, True)}
        if world > 1:
            for k in list(out):
                v = torch.tensor([out[k]], dtype=torch.float64, device="cuda")
                dist.all_reduce(v, op=dist.ReduceOp.SUM)
                out[k + "_sum_all_ranks"] = float(v.item())
        return out

    pcie = pcie_rates() if not os.environ.get("BENCH_DEBUG_SKIP_PCIE") else {"h2d_gbs": 55.0, "d2h_gbs": 52.0, "duplex_each_gbs": 44.0}

    # ---- end to end -------------------------------------------------------------------------------------------------
    # (1) headline: the calls the plugin makes (xr_backend.py: SHEngine.synthesis / .analysis) with a plain pageable numpy
    #     coefficient array in and numpy arrays out; the engine hands its results out of the library's page-locked pool.
    # (2) all-pinned buffers straight through the C ABI (shb200_*_host), two calls.
    # (3) the fused, pipelined round trip (shb200_roundtrip_host): same bytes over PCIe, both directions busy.
    eng = sb.SHEngine(device=local)
    c_page = S.random_coefficients(NMAX, BATCH, kaula=True, seed=S.SEED + rank)         # ordinary numpy memory
    nb_c, nb_g = BATCH * nsh * 8, BATCH * NLAT * NLON * 8
    ke = max(3, min(args.steps, 10))

    def timed_e2e(fn):
        res = fn()
        res = fn()
        res = fn()    # holding the previous result like the timed loop does: the page-locked pool reaches its steady state (a held
        #               result + a fresh one; the first step that needs the second block pays ~100 ms of cudaHostAlloc) before the timed region
        barrier()
        torch.cuda.synchronize()
        t0_ = time.perf_counter()
        for _ in range(ke):
            res = fn()
        torch.cuda.synchronize()
        dt_ = torch.tensor([time.perf_counter() - t0_], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt_, op=dist.ReduceOp.MAX)
        return world * BATCH * ke / float(dt_.item()), float(dt_.item()) / ke, res

    dbg = []

    def step_plugin():
        t0_ = time.perf_counter()
        g_ = eng.synthesis(c_page, lon=lon, lat=lat, nmax=NMAX)
        t1_ = time.perf_counter()
        a_ = eng.analysis(g_, NMAX, lon, lat, quadrature="gauss")
        dbg.append((round((t1_ - t0_) * 1e3, 2), round((time.perf_counter() - t1_) * 1e3, 2)))
        return a_

    v_plugin, t_plugin, a_plugin = timed_e2e(step_plugin)
    if os.environ.get("BENCH_DEBUG_E2E"):
        print("plugin (synthesis ms, analysis ms) per step:", dbg, file=sys.stderr)
    e2e_err = float((np.abs(a_plugin - c_page).max(axis=1) / np.abs(c_page).max(axis=1)).max())
    del a_plugin

    grid_h = capi.pinned_empty((BATCH, NLAT, NLON))
    out_h = capi.pinned_empty((BATCH, nsh))
    c_pin = capi.pinned_empty((BATCH, nsh))
    c_pin[...] = c_page

    def step_pinned():
        plan.synthesis_host(BATCH, c_pin.ctypes.data, nsh, 1, grid_h.ctypes.data, NLAT * NLON, NLON, 1)
        plan.analysis_host(BATCH, grid_h.ctypes.data, NLAT * NLON, NLON, 1, out_h.ctypes.data, nsh, 1)

    v_pinned, t_pinned, _ = timed_e2e(step_pinned)

    def step_fused():
        plan.roundtrip_host(BATCH, c_pin.ctypes.data, nsh, 1, grid_h.ctypes.data, NLAT * NLON, NLON, 1, out_h.ctypes.data, nsh, 1)

    v_fused, t_fused, _ = timed_e2e(step_fused)
    fused_err = float((np.abs(out_h - c_page).max(axis=1) / np.abs(c_page).max(axis=1)).max())
    # floor of one step on this link: each direction carries coefficients + grid once; both directions can run at once
    t_floor = (nb_c + nb_g) / (pcie["duplex_each_gbs"] * 1e9)
    e2e = {"value": v_plugin, "unit": UNIT, "h2d_bytes_per_step": nb_c + nb_g, "d2h_bytes_per_step": nb_c + nb_g, "steps": ke,
           "path": "SHEngine.synthesis + SHEngine.analysis, pageable numpy coefficients in, numpy out (page-locked pool)",
           "ms_per_step": t_plugin * 1e3, "roundtrip_max_rel_err": e2e_err,
           "pinned_c_abi": {"value": v_pinned, "ms_per_step": t_pinned * 1e3, "path": "shb200_synthesis_host + shb200_analysis_host, all buffers page-locked"},
           "fused_roundtrip": {"value": v_fused, "ms_per_step": t_fused * 1e3, "path": "shb200_roundtrip_host (grid to the host and back, pipelined)", "roundtrip_max_rel_err": fused_err},
           "pcie": pcie, "pcie_floor_ms_per_step": t_floor * 1e3, "numa_bind": numa,
           "frac_of_pcie_floor": {"plugin": t_floor / t_plugin, "pinned_c_abi": t_floor / t_pinned, "fused_roundtrip": t_floor / t_fused}}
    del grid_h, out_h, c_pin
    os.sched_setaffinity(0, cpus_before)

    launches = args.steps * (plan.info.launches_synthesis + plan.info.launches_analysis)
    cfg4 = None
    info_dmma, info_fft = bool(plan.info.dmma), bool(plan.info.lon_fft)
    if world > 1:
        plan.close()
        del coef, gridbuf, coef_out, eng
        torch.cuda.empty_cache()
        cfg4 = cfg4_sharded(torch, dist, sb, S, rank, world, local)
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": workload_config(world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "roundtrip_max_rel_err": rt_err,
                "dmma": info_dmma, "lon_fft": info_fft, "kernels_last_call": kernels}
        if cfg4 is not None:
            line["cfg4_sharded"] = cfg4
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_sample(1, 0)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
