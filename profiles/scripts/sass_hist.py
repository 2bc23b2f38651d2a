#!/usr/bin/env python
"""SASS opcode histogram of every kernel in libcofdm_b200.so (cuobjdump -sass): the static proof that the hot kernels use
the Blackwell paths they claim -- UBLKCP (cp.async.bulk = TMA bulk copies, .S2G = stores), SYNCS (mbarrier), UCGABAR
(cluster barriers), packed FADD2 / FMUL2 / FFMA2 arithmetic, REDUX -- and how large each kernel is.
    python profiles/scripts/sass_hist.py > profiles/r02_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "c-ofdm_b200", "libcofdm_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
cur, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op, suffix = m.group(2), m.group(3) or ""
        hist[cur][op] += 1
        if op == "UBLKCP" and ".S2G" in suffix or (op == "UBLKCP" and "S." in line and "G.S" in line):
            hist[cur]["UBLKCP(store)"] += 1
KEY = ["UBLKCP", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "FADD2", "FMUL2", "FFMA2", "REDUX", "SHFL", "LDS", "STS", "LDG", "STG", "BAR", "MUFU", "DADD", "DMUL", "DFMA"]
print(f"# {os.path.basename(lib)}: static SASS opcode counts per kernel (sm_100a).  columns: total, then {', '.join(KEY)}")
for name, h in hist.items():
    tot = sum(v for k, v in h.items() if "(" not in k)
    short = demangle(name)
    short = re.sub(r"\(.*", "", short)[:110]
    print(f"{short}\n    total {tot:6d}  " + "  ".join(f"{k}:{h.get(k, 0)}" for k in KEY if h.get(k, 0)))
