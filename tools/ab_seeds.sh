#!/bin/bash
# A/B of the polar-skipping seeds of the tiny-batch recursion kernels (SHB200_SEEDS=0 / 1) on one nmax-2160 field.
run() { echo "== $*"; env "$@" CONFIGS_OUT=gpurun_out/configs_seeds_tmp.json timeout 200 python tools/bench_configs.py "=cfg4 nmax2160 5min B=1" "=cfg5 nmax256 gauss B=4" 2>&1 | grep -o '"config": "[a-z0-9]*\|"synthesis_ms": [0-9.]*\|"analysis_ms": [0-9.]*\|"syn_legendre": [0-9.]*\|"ana_legendre": [0-9.]*\|"legendre_visited": {[^}]*}' | tr '\n' ' '; echo; }
run SHB200_SEEDS=0
run SHB200_SEEDS=1
