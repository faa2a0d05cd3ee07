#!/usr/bin/env python
"""Rebuild the tracked profile artefacts from one gpurun session (run here, after the GPU call):
  gpurun_out/bench_{TAG}.json, {TAG}_launches.csv, prof_{TAG}.ncu-rep  ->  profiles/
usage: python tools/refresh_profiles.py [tag]      (tag = r02 ...: the one given to tools/gpu_session.sh)"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"

def sh(cmd):
    return subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout

# 1. bench line
b = json.loads(open(f"{G}/bench_{TAG}.json").read().strip().splitlines()[-1])
open(f"{P}/bench_{TAG}.json", "w").write(json.dumps(b) + "\n")
if os.path.exists(f"{G}/bench_{TAG}_reference.json"):
    open(f"{P}/bench_{TAG}_reference.json", "w").write(open(f"{G}/bench_{TAG}_reference.json").read().strip().splitlines()[-1] + "\n")
# 2. launch list + shares
src = open(f"{G}/{TAG}_launches.csv").read()
open(f"{P}/{TAG}_launches.csv", "w").write(src)
rows = [r for r in csv.reader(src.splitlines()) if len(r) > 5]
hdr, seq = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = d["Kernel Name"].split("(")[0].replace("void ", "")
    v, unit = float(d["Metric Value"].replace(",", "")), d["Metric Unit"]
    seq.append((name, v / 1000 if unit.startswith("n") else (v if unit.startswith("u") else v * 1000)))
cut = next((i for i, (n, _) in enumerate(seq) if n.endswith("tab<1>")), len(seq))
while cut > 0 and not (seq[cut - 1][0].endswith("FPT_") or "k_lon_ana_fft" in seq[cut - 1][0] or "k_unpack" in seq[cut - 1][0]):
    cut -= 1

def agg(s):
    a = {}
    for n, t in s:
        if n.startswith("at::"):
            n = "torch elementwise/reduce (round-trip error check)"
        x = a.setdefault(n, [0, 0.0]); x[0] += 1; x[1] += t
    return a

st = b["roofline"]["stages_ms"]; ssum = sum(st.values())
with open(f"{P}/{TAG}_launch_shares.txt", "w") as f:
    f.write(f"# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (gpu__time_duration.sum, --clock-control none), build of {TAG}\n"
            "# per-launch times are cold-cache and serialised: compare SHARES with bench.py's stages_ms, not absolutes\n")
    for title, s in (("device-resident steps (warm-up + timed + stage-timing pass: 64-field launches, <4> tile shapes)", seq[:cut]),
                     ("host-path e2e loop (shb200_*_host: four 16-field chunks per call, <1> tile shapes)", seq[cut:])):
        a = agg(s); tot = sum(x[1] for x in a.values()) or 1.0
        f.write(f"\n## {title}\n")
        for k, (n, t) in sorted(a.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:70]:70s} launches {n:4d}  total {t/1000:9.3f} ms  mean {t/n:9.1f} us  share {100*t/tot:5.1f}%\n")
    f.write("\n# bench.py CUDA-event stage times of the same build (device-resident step; ms per launch, share of the step):\n")
    for k, v in st.items():
        f.write(f"{k:16s} {v:7.4f} ms  {100*v/ssum:5.1f}%\n")
# 3. full capture: summary, traffic, stall profiles
raw = sh(f"ncu -i {G}/prof_{TAG}.ncu-rep --page raw --csv 2>/dev/null")
open("/tmp/final_raw.csv", "w").write(raw)
open(f"{P}/{TAG}_ncu_full_summary.txt", "w").write(sh(f"python {ROOT}/tools/ncu_raw_summary.py /tmp/final_raw.csv"))
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]; idx = {x: i for i, x in enumerate(h)}
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tr = {}
for r in rr[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].replace("fft::", "").replace("shb::", "")
    tr[name] = {k2: float(r[idx[k1]].replace(",", "")) * mul[units[idx[k1]]] for k1, k2 in (("dram__bytes_read.sum", "dram_bytes_read"), ("dram__bytes_write.sum", "dram_bytes_write"))}
json.dump({"source": f"ncu --set full --clock-control none, profiles/{TAG}_ncu_full_summary.txt (one launch each, cfg3, build of {TAG})", "kernels": tr},
          open(f"{P}/{TAG}_traffic.json", "w"), indent=1)
for k in ("k_leg_ana_tab", "k_leg_syn_tab", "k_lon_ana_fft", "k_lon_syn_fft"):
    sh(f"ncu -i {G}/prof_{TAG}.ncu-rep --page source --csv --kernel-name regex:{k} > /tmp/src_{k}.csv 2>/dev/null")
    body = sh(f"python {ROOT}/tools/ncu_src_summary.py /tmp/src_{k}.csv 24 | cut -c1-200")
    open(f"{P}/{TAG}_stalls_{k}.txt", "w").write(f"# ncu --set full --import-source on, source page of {k} (profiles/{TAG}_ncu_full_summary.txt is the same capture, build of {TAG}): top SASS lines by stall samples\n" + body)
print(open(f"{P}/{TAG}_launch_shares.txt").read())
print(list(tr))
