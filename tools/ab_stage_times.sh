#!/bin/bash
# A/B of the device-resident cfg3 step (bench.py stage times, CUDA events) under environment switches of the library:
#   tools/ab_stage_times.sh "SHB200_PACK_M=16" "SHB200_PACK_M=32" ...      -> gpurun_out/ab_<n>.json (one bench line each)
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs BENCH_DEBUG_SKIP_DGEMM=1 BENCH_DEBUG_SKIP_PCIE=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
  python - "$envs" gpurun_out/ab_$i.json <<'PY'
import json, sys
b = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], "ms_per_step %.4f" % b["ms_per_step"], {k: round(v, 4) for k, v in b["roofline"]["stages_ms"].items()}, "err %.2e" % b["roundtrip_max_rel_err"])
PY
done
