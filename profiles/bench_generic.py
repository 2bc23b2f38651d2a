#!/usr/bin/env python
"""Any-size path (generic.cuh) on BASELINE.json configs[4]: FFT 4096, CP 1024, 1920 data + 128 pilot sub-carriers,
8 symbols, 64-QAM.  Device-resident cf32, batch-size sweep, CUDA-event timing.  One JSON line.
    python profiles/bench_generic.py > profiles/r01_generic_path.json
Algorithmic bytes: rx 9*5120*8 + 11520 = 380 160 B/frame; tx 11520 + frame_len*8.
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cofdm_b200 as cb  # noqa: E402
from cofdm_b200 import synth  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    d = tempfile.mkdtemp()
    cfg = synth.write_config(os.path.join(d, "config_big.txt"), fft_size=4096, cp_size=1024, num_data_subc=1920, num_pilot_subc=128,
                             num_symb=8, pr_sin_len=128, modType=6)
    m = cb.Modem(cfg, device=0)
    m.use_torch_stream()
    s = m.sizes
    peak_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peak_file))["hbm_gbs"] if os.path.exists(peak_file) else 6650.0
    rx_bytes = s.rx_len * 8 + s.usefull_size
    tx_bytes = s.usefull_size + s.output_size * 8
    rows = []
    g = torch.Generator(device="cuda").manual_seed(7)
    batches = [int(x) for x in os.environ.get("BIG_BATCHES", "64,256,1024,4096,16384").split(",")]
    for n in batches:
        pay = torch.randint(0, 256, (n, s.usefull_size), dtype=torch.uint8, device="cuda", generator=g)
        frames = torch.empty((n, s.output_size), dtype=torch.complex64, device="cuda")
        out = torch.empty((n, s.usefull_size), dtype=torch.uint8, device="cuda")
        tx_ms = timed(lambda: m.tx_batch(pay, cb.CF32, out=frames))
        frames *= float(s.mult)
        rx_ms = timed(lambda: m.rx_aligned_batch(frames, n_frames=n, frame_stride=s.output_size, offset=s.t2sin_size, out=out, count_ambiguous=False))
        bad = int((out != pay).any(dim=1).sum().item())
        m.enable_timing(True)
        m.rx_aligned_batch(frames, n_frames=n, frame_stride=s.output_size, offset=s.t2sin_size, out=out, count_ambiguous=False)
        torch.cuda.synchronize()
        st = m.last_stage_ms()
        m.enable_timing(False)
        rows.append({"frames": n, "tx_ms": tx_ms, "rx_ms": rx_ms, "frames_with_errors": bad, "rx_acquire_ms": st["acquire"], "rx_demod_ms": st["demod"],
                     "tx_gbs": tx_bytes * n / tx_ms / 1e6, "rx_gbs": rx_bytes * n / rx_ms / 1e6,
                     "tx_frac": tx_bytes * n / tx_ms / 1e6 / peak, "rx_frac": rx_bytes * n / rx_ms / 1e6 / peak,
                     "rx_msamples_s": n * s.output_size / rx_ms / 1e3})
    print(json.dumps({"config": "fft 4096 / cp 1024 / 1920+128 sub-carriers / 8 symbols / 64-QAM", "fused_path": int(s.fused_path),
                      "rx_bytes_per_frame": rx_bytes, "tx_bytes_per_frame": tx_bytes, "peak_gbs": peak, "by_batch": rows}))


if __name__ == "__main__":
    main()
