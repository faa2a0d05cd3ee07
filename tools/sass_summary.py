#!/usr/bin/env python
"""Per-kernel SASS evidence of the shipped library: counts of the instructions that prove the hardware path
(DMMA = FP64 tensor pipe, UBLKCP = bulk TMA copies, SYNCS = mbarrier, LDGSTS = cp.async, USETMAXREG = register reallocation,
DFMA/DADD/DMUL = FP64 CUDA-core pipe, BAR = CTA / named barriers).  Runs here (no GPU): cuobjdump -sass on libshb200.so.
usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "shxarray_b200", "libshb200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
keys = ["DMMA", "UBLKCP", "SYNCS", "LDGSTS", "USETMAXREG", "DFMA", "DADD", "DMUL", "LDS", "STS", "LDG", "STG", "BAR"]
cur, rows = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        rows[cur] = dict.fromkeys(keys, 0)
        rows[cur]["total"] = 0
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    rows[cur]["total"] += 1
    for k in keys:
        if op == k or op.startswith(k + "."):
            rows[cur][k] += 1
arch = re.search(r"arch = (sm_\w+)", txt)
print(f"# cuobjdump -sass shxarray_b200/libshb200.so ({arch.group(1) if arch else '?'}), instruction counts per kernel (static, whole function)")
print(f"# {'kernel':78s} " + " ".join(f"{k:>7s}" for k in ["total"] + keys))
tot = dict.fromkeys(["total"] + keys, 0)
for name in sorted(rows, key=lambda n: demangle(n)):
    d = demangle(name)
    d = re.sub(r"\(.*", "", d).replace("shb::", "").replace("void ", "")
    d = re.sub(r"\(int\)", "", d)
    r = rows[name]
    print(f"{d[:80]:80s} " + " ".join(f"{r[k]:7d}" for k in ["total"] + keys))
    for k in tot:
        tot[k] += r[k]
print(f"{'ALL KERNELS':80s} " + " ".join(f"{tot[k]:7d}" for k in ["total"] + keys))
lib = subprocess.run(["bash", "-c", f"ldd {so} | grep -i -E 'cublas|cufft|cusolver|nccl|triton' | wc -l"], capture_output=True, text=True).stdout.strip()
print(f"# linked vendor math libraries (cublas/cufft/cusolver/nccl): {lib}")
