#!/bin/bash
# A/B of the CUDA-graph replay of launch-latency-bound transforms (SHB200_GRAPHS=0 / 1): per-call wall time of cfg1 / cfg5 through
# SHEngine (tools/latency_small.py) and of the fused spectral product (tools/bench_multiply.py).
for G in 0 1; do
  echo "== SHB200_GRAPHS=$G"
  SHB200_GRAPHS=$G python tools/latency_small.py 2>&1 | grep -v "^$"
  SHB200_GRAPHS=$G python tools/bench_multiply.py 2>&1 | grep -v "^$" | tail -4 | cut -c1-300
done
