#!/usr/bin/env python
"""Per-source-line stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K`:
shows how a kernel's time splits over its source lines (samples are proportional to warp-resident time).
usage: python tools/ncu_line_summary.py dump.csv [min_share_percent]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
cur, hdr, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) >= 8 and r[0].isdigit() and r[2] == "-":          # a CUDA source line (SASS rows carry an address)
        i_s, i_n = hdr.index("# Samples"), hdr.index("Instructions Executed")
        out.append((cur, int(r[0]), int(r[i_s] or 0), int(r[i_n] or 0), r[1].strip()))
tot = sum(o[2] for o in out) or 1
toti = sum(o[3] for o in out) or 1
print(f"total samples {tot}, warp instructions {toti}")
for f, ln, s, n, src in out:
    if 100.0 * s / tot >= thr:
        print(f"{f}:{ln:4d} {100.0*s/tot:5.1f}% samples {100.0*n/toti:5.1f}% insts  {src[:110]}")
