#!/bin/bash
# A/B of the longitude kernels: generic in-place FFT (SHB200_FFT3=0) vs the three-pass kernels (default), stage times of the
# BASELINE configs + parity tests with the three-pass kernels.  Run through gpurun from the repo root.
mkdir -p gpurun_out
for v in 0 1 2; do
  echo "== variant $v (0: in-place generic, 1: three-pass two CTAs per SM, 2: three-pass persistent staged)"
  SHB200_FFT3=$((v>0)) SHB200_FFT3P=$((v>1)) python tools/bench_configs.py cfg2 cfg3 cfg5 "cfg4b" > gpurun_out/ab_fft3_$v.log 2>&1
  python - <<P
import json
for l in open("gpurun_out/ab_fft3_$v.log"):
    if l.startswith("{"):
        d = json.loads(l); s = d["stages_ms"]
        print(d["config"][:28], "syn_lon", s.get("syn_lon"), "ana_lon", s.get("ana_lon"), "step", round(d["synthesis_ms"] + d["analysis_ms"], 4), "rt_err", d["roundtrip_rel_err"])
    elif "Error" in l or "error" in l: print(l.strip()[:300])
P
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
