#!/usr/bin/env python
"""Ad-hoc stress of the rx parity check: thousands of impaired frames per modulation against the oracle, every tap
compared (tests/parity_checks.py::check_rx_against_oracle), wider CFO and noise than the unit tests.  One JSON line."""
import json, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cofdm_b200 as cb  # noqa: E402
from cofdm_b200 import synth  # noqa: E402
import parity_checks as pc  # noqa: E402
from oracle import oracle as O  # noqa: E402

O.build("port")
d = tempfile.mkdtemp()
res = []
for mt, n, cfo, noise in ((4, 1500, 0.012, 3.0), (2, 1000, 0.02, 6.0), (6, 1000, 0.006, 1.0), (8, 600, 0.003, 0.4), (1, 600, 0.02, 10.0)):
    cfg = synth.write_config(os.path.join(d, f"c{mt}.txt"), modType=mt)
    o = O.Oracle("port", cfg)
    m = cb.Modem(cfg, device=0)
    for seed in (101, 202):
        pay, rec = pc.impaired_records(o, n, seed=seed + mt, cfo_max=cfo, noise=noise)
        for fmt in ("i16", "cf32"):
            st = pc.check_rx_against_oracle(m, o, rec, fmt)
            res.append({"mod": mt, "frames": n, "fmt": fmt, "cfo_max": cfo, "noise_lsb": noise,
                        **{k: (float(v) if isinstance(v, (float, np.floating)) else int(v)) for k, v in st.items()}})
    m.close()
print(json.dumps(res))
