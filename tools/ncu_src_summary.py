#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: top instructions by stall samples with their dominant stall reasons.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv ; python tools/ncu_src_summary.py src.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[0] != hdr[0]]
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
print("total samples", tot)
agg = {}
for r in data:
    for c in stall_cols:
        agg[c] = agg.get(c, 0) + int(r[idx[c]] or 0)
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[idx[c]] or 0), c) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {int(r[idx['# Samples']]):6d} {int(r[idx['Instructions Executed']]):9d}  {r[idx['Source']].strip():70s} {[(c[6:], v) for v, c in st if v]}")
