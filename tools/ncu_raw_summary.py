#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` into one compact table per kernel (profiles/*.txt).
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_raw_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("lts__t_bytes.sum", "l2_bytes"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor(DMMA)_%"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_inst_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"), ("smsp__inst_executed.sum", "warp_insts")]
for d in data:
    print("-" * 100)
    for key, name in want:
        if key in idx:
            v = d[idx[key]]
            if key == "Kernel Name":
                v = v.split("(")[0]
            print(f"{name:22s} {v:>24s} {units[idx[key]]}")
