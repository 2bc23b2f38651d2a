#!/usr/bin/env python
"""Device-time measurements of the synchronisation kernels (SURVEY.md section 8d byte counts):
sync-tone block detector, preamble search, and the rx.cpp-style streaming receiver.  One JSON line.
    python profiles/bench_sync.py > profiles/r01_sync_kernels.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cofdm_b200 as cb  # noqa: E402
from cofdm_b200 import synth  # noqa: E402


def timed(fn, reps=10, warm=3):
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
    cfg = os.path.join(ROOT, "config", "config.txt")
    m = cb.Modem(cfg, device=0)
    s = m.sizes
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    out = {"peak_gbs": peak}
    n = 1 << 27                                                   # samples: 1 GiB as cf32, far larger than L2
    g = torch.Generator(device="cuda").manual_seed(1)
    x16 = torch.randint(-500, 500, (n, 2), dtype=torch.int16, device="cuda", generator=g)
    x32 = torch.view_as_complex(x16[: n // 2].to(torch.float32).contiguous())
    for name, x, bps in (("ci16", x16, 4), ("cf32", x32, 8)):
        ns = x.shape[0]
        ms = timed(lambda: m.t2sin_metric(x))
        out[f"t2sin_{name}"] = {"samples": ns, "ms": ms, "gsamples_s": ns / ms / 1e6, "gb_s": ns * (bps + 4 / 256) / ms / 1e6,
                                "frac_of_peak": ns * (bps + 4 / 256) / ms / 1e6 / peak}
    nw = 1 << 16
    starts = (torch.arange(nw, device="cuda", dtype=torch.int64) * 1531) % (n // 2 - 2048)
    ms = timed(lambda: m.preamble_search(x16, starts), reps=5)
    flop = nw * s.cor_size * s.pr_sin_len * 8
    out["preamble_search_ci16"] = {"windows": nw, "ms": ms, "windows_s": nw / ms * 1e3, "gb_s": nw * (s.cor_size + s.pr_sin_len) * 4 / ms / 1e6,
                                   "tflop_s": flop / ms / 1e9}
    # streaming receiver (cofdm_rx_stream / cofdm_rx_stream_sharded): synthetic capture, frames every ~1.2 frame lengths
    nfr, reps = 4000, 8
    pay = synth.payloads(nfr, s.usefull_size, seed=5)
    fr = m.tx_batch(pay, cb.CI16)
    rng = np.random.default_rng(3)
    cap, _ = synth.capture(fr[..., 0].astype(np.float64) + 1j * fr[..., 1], gaps=rng.integers(300, 2500, nfr), noise_sigma=3.0, seed=4, tail=s.output_size * 41)
    blk = s.output_size * s.rx_buf_size
    cap = cap[: cap.shape[0] // blk * blk]
    t0 = time.perf_counter()
    pos, by = m.rx_stream(cap)
    dt = time.perf_counter() - t0
    ok = int(sum(np.array_equal(b, p) for b, p in zip(by, pay[: len(by)])))
    out["rx_stream_host_capture_1shard"] = {"capture_samples": int(cap.shape[0]), "frames_found": int(len(pos)), "frames_sent": nfr, "payload_ok": ok,
                                            "seconds": dt, "frames_s": len(pos) / dt, "msamples_s": cap.shape[0] / dt / 1e6,
                                            "note": "device-side state machine (one CTA), capture uploaded from pageable host memory inside the timed region"}
    big = torch.from_numpy(cap).cuda().repeat(reps, 1)            # the same capture back to back: reps*nfr frames
    m.enable_timing(True)
    rows = []
    for shards in (1, 8, 74, 148, 296, 592, 1184):
        m.rx_stream(big, shards=shards)                           # warm-up (buffers)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pos, by, unmerged = m.rx_stream(big, shards=shards, return_unmerged=True)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        pos2, _ = m.rx_stream(big, shards=shards, want_bytes=False)
        dt_scan = time.perf_counter() - t0
        scan_ms = m.last_kernel_ms()
        ok = int((by == np.tile(pay[: len(by) // reps + 1], (reps, 1))[: len(by)]).all(axis=1).sum()) if len(by) == reps * nfr else -1
        rows.append({"shards": shards, "frames_found": int(len(pos)), "payload_ok": ok, "unmerged": int(unmerged), "seconds_scan_plus_demod": dt,
                     "frames_s": len(pos) / dt, "msamples_s": big.shape[0] / dt / 1e6, "seconds_scan_only_wall": dt_scan, "scan_kernel_ms": scan_ms,
                     "scan_kernel_msamples_s": big.shape[0] / scan_ms / 1e3 if scan_ms else None})
    out["rx_stream_device_capture"] = {"capture_samples": int(big.shape[0]), "frames_sent": reps * nfr, "by_shards": rows,
                                       "note": "int16 capture resident in HBM; wall clock of the whole call (scan kernel, list merge on the host, gather + rx kernels, D2H of the bytes)"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
