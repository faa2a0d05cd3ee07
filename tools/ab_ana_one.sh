#!/bin/bash
# A/B of the single-field analysis kernel (k_leg_ana_one, SHB200_ANA_ONE=0 / 1) on the one-field configs.
for V in 0 1; do
  echo "== SHB200_ANA_ONE=$V"
  SHB200_ANA_ONE=$V CONFIGS_OUT=gpurun_out/configs_ana_one_$V.json python tools/bench_configs.py "=cfg1 nmax60 shlib 1deg B=1" "=cfg4 nmax2160 5min B=1" 2>&1 | cut -c1-900
done
