#!/bin/bash
# A/B of the rows per block of k_leg_syn_dfma on one nmax-2160 field (finer blocks skip more of the polar triples).
for R in 64 32 128; do
  echo "== SHB200_SYN_ROWS=$R"
  SHB200_SYN_ROWS=$R CONFIGS_OUT=gpurun_out/configs_tmp.json timeout 100 python tools/bench_configs.py "=cfg4 nmax2160 5min B=1" 2>&1 | grep -o '"synthesis_ms": [0-9.]*\|"syn_legendre": [0-9.]*\|"legendre_visited": {[^}]*}' | tr '\n' ' '; echo
done
