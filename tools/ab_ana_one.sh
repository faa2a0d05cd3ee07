#!/bin/bash
# A/B of the single-field analysis kernel (k_leg_ana_one) on one nmax-2160 field: off (k_leg_ana_small), rows per thread / block shapes.
run() { echo "== $*"; env "$@" CONFIGS_OUT=gpurun_out/configs_ana_one_tmp.json timeout 200 python tools/bench_configs.py "=cfg4 nmax2160 5min B=1" 2>&1 | grep -o '"analysis_ms": [0-9.]*\|"ana_legendre": [0-9.]*\|"roundtrip_rel_err": [0-9.e-]*' | tr '\n' ' '; echo; }
run SHB200_ANA_ONE=0
run SHB200_ANA_ONE_RPT=4
run SHB200_ANA_ONE_RPT=4 SHB200_ANA_ONE_NT=64
run SHB200_ANA_ONE_RPT=4 SHB200_ANA_ONE_NT=128
run SHB200_ANA_ONE_RPT=2 SHB200_ANA_ONE_NT=128
run SHB200_ANA_ONE_RPT=2 SHB200_ANA_ONE_NT=192
run SHB200_ANA_ONE_RPT=2 SHB200_ANA_ONE_NT=256
