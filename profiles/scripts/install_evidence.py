#!/usr/bin/env python
"""Copy the files evidence.sh left in gpurun_out/ into profiles/ (run in the build container after the gpurun call)."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, "bench_final.json"), os.path.join(P, "r01_bench_1gpu.json"))
shutil.copy(os.path.join(G, "bench_ref.json"), os.path.join(P, "r01_bench_reference_arm.json"))
with open(os.path.join(G, "launches.csv")) as f, open(os.path.join(P, "r01_launch_list.csv"), "w") as o:
    o.writelines(l for l in f if not l.startswith("=="))
rows = [r for r in csv.reader(open(os.path.join(P, "r01_launch_list.csv"))) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = {}
for r in rows[1:]:
    agg.setdefault(r[ki].split("(")[0], []).append(int(r[vi]))
for k, v in agg.items():
    print(f"{k:60s} {len(v)} launches, mean {sum(v) / len(v) / 1e3:.1f} us")
d = json.loads(open(os.path.join(G, "bench_final.json")).read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "rx_ms", "tx_ms", "bit_errors")}, d["roofline"]["frac"], d["roofline"]["tx_frac"], d["e2e"]["value"])
