#!/bin/bash
# A/B of the fold of k_leg_ana_one: transposing shuffle butterfly (0) against a transpose through shared memory (1).
run() { echo "== $*"; env "$@" CONFIGS_OUT=gpurun_out/configs_ana_one_tmp.json timeout 200 python tools/bench_configs.py "=cfg4 nmax2160 5min B=1" 2>&1 | grep -o '"analysis_ms": [0-9.]*\|"ana_legendre": [0-9.]*\|"roundtrip_rel_err": [0-9.e-]*' | tr '\n' ' '; echo; }
run SHB200_ANA_ONE_FOLD=0
run SHB200_ANA_ONE_FOLD=1
run SHB200_ANA_ONE_FOLD=1 SHB200_ANA_ONE_NT=64
run SHB200_ANA_ONE_FOLD=1 SHB200_ANA_ONE_NT=128
