#!/usr/bin/env python
"""Device-resident int16 (SDR wire format) variants of the tx and rx passes next to the cf32 ones.  One JSON line."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cofdm_b200 as cb  # noqa: E402


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
    m = cb.Modem(os.path.join(ROOT, "config", "config.txt"), device=0)
    m.use_torch_stream()
    s = m.sizes
    F = 1 << 18
    g = torch.Generator(device="cuda").manual_seed(3)
    pay = torch.randint(0, 256, (F, s.usefull_size), dtype=torch.uint8, device="cuda", generator=g)
    f16 = torch.empty((F, s.output_size, 2), dtype=torch.int16, device="cuda")
    f32 = torch.empty((F, s.output_size), dtype=torch.complex64, device="cuda")
    out = torch.empty((F, s.usefull_size), dtype=torch.uint8, device="cuda")
    res = {"frames": F}
    res["tx_cf32_ms"] = timed(lambda: m.tx_batch(pay, cb.CF32, out=f32))
    res["tx_ci16_ms"] = timed(lambda: m.tx_batch(pay, cb.CI16, out=f16))
    f32 *= float(s.mult)
    res["rx_cf32_ms"] = timed(lambda: m.rx_aligned_batch(f32, n_frames=F, frame_stride=s.output_size, offset=s.t2sin_size, out=out, count_ambiguous=False))
    res["rx_cf32_bad"] = int((out != pay).any(dim=1).sum().item())
    res["rx_ci16_ms"] = timed(lambda: m.rx_aligned_batch(f16, n_frames=F, frame_stride=s.output_size, offset=s.t2sin_size, out=out, count_ambiguous=False))
    res["rx_ci16_bad"] = int((out != pay).any(dim=1).sum().item())
    for k in ("tx_cf32", "tx_ci16", "rx_cf32", "rx_ci16"):
        res[k + "_mframes_s"] = F / res[k + "_ms"] / 1e3
    print(json.dumps(res))


if __name__ == "__main__":
    main()
