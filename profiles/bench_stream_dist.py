#!/usr/bin/env python
"""BASELINE.json configs[3] (stream sync over N GPUs): ONE long int16 capture (default 2^30 samples = 4 GiB, frames at
random gaps over a noise floor) cut into contiguous ranges of whole SDR blocks, one per rank (+ one overlap block); every
rank scans and demodulates its range on its own GPU (itself split into CTA-sized shards), then ONE fixed-size NCCL
all_gather of the int64 position lists (cofdm_b200.stream.gather_frame_lists) and a local merge on every rank.  Payloads
stay on the rank that decoded them.  Strong scaling: the capture is the same for every N.
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/bench_stream_dist.py [--log2-samples 30]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cofdm_b200 as cb  # noqa: E402
from cofdm_b200 import stream as st, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-samples", type=int, default=30)
    ap.add_argument("--host-bytes", action="store_true", help="copy the payloads back to (pinned) host memory instead of leaving them on the GPU")
    ap.add_argument("--shards", type=int, default=888, help="capture ranges per GPU = scanner CTAs: 6 resident per SM x 148")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m = cb.Modem(os.path.join(ROOT, "config", "config.txt"), device=local)
    m.use_torch_stream()
    s = m.sizes
    nfr = 4000
    pay = synth.payloads(nfr, s.usefull_size, seed=5)
    fr = m.tx_batch(pay, cb.CI16)
    rng = np.random.default_rng(3)
    cap, starts = synth.capture(fr[..., 0].astype(np.float64) + 1j * fr[..., 1], gaps=rng.integers(300, 2500, nfr), noise_sigma=3.0, seed=4,
                                tail=s.output_size * 41)
    blk = st.block_samples(s)
    cap = cap[: cap.shape[0] // blk * blk]
    L = cap.shape[0]
    n_total = (1 << args.log2_samples) // blk * blk
    # every rank materialises only its own slice (+ overlap) of the tiled capture, on its GPU
    s0, s1, b0, b1 = st.shard_slice(n_total, s, rank, world)
    cap_d = torch.from_numpy(cap).to(dev)
    mine = torch.empty((s1 - s0, 2), dtype=torch.int16, device=dev)
    step = 1 << 25
    for a in range(s0, s1, step):
        e = min(s1, a + step)
        mine[a - s0:e - s0] = cap_d[torch.arange(a, e, device=dev) % L]
    del cap_d

    max_frames = (n_total // world + 2 * blk) // (s.ofdm_len * s.num_symb) + 2       # upper bound on a rank's frames, the same on every rank

    def one_pass():
        pos, by = m.rx_stream(mine, shards=args.shards, bytes_on_device=not args.host_bytes)
        pos_abs = np.asarray(pos, dtype=np.int64) + s0
        if world == 1:
            return pos_abs, by, (pos_abs, [(0, len(pos_abs))], 0)
        lists = st.gather_positions(pos_abs, b0, b1, max_frames, device=dev)      # the one collective of the job
        return pos_abs, by, st.merge_ranges(lists, s)

    for _ in range(2):
        one_pass()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    m.enable_timing(True)
    t0 = time.perf_counter()
    pos_abs, by, (mpos, ranges, unmerged) = one_pass()
    torch.cuda.synchronize()
    t_local = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    stages = m.last_stage_ms()
    m.enable_timing(False)
    # every payload this rank decoded, against what was sent: the frame at capture offset r carries payload j(r)
    r = pos_abs % L
    j = np.clip(np.searchsorted(starts + s.t2sin_size - 64, r) - 1, 0, nfr - 1)
    near = np.abs(r - (starts[j] + s.t2sin_size)) < 64
    if args.host_bytes:
        good = int((near & (by == pay[j]).all(axis=1)).sum())
    else:   # payloads are on the GPU: compare there
        pay_d = torch.from_numpy(pay).to(dev)
        good = int((torch.from_numpy(near).to(dev) & (by == pay_d[torch.from_numpy(j).to(dev)]).all(dim=1)).sum().item())
    tot = torch.tensor([good, len(pos_abs), int(t_local * 1e6), int(stages["scan"] * 1e3)], dtype=torch.int64, device=dev)
    mx = tot.clone()
    if world > 1:
        dist.all_reduce(tot)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    if rank == 0:
        scan_ms = mx[3].item() / 1e3
        print(json.dumps({"n_gpus": world, "capture_samples": int(n_total), "capture_bytes": int(n_total) * 4, "frames_after_merge": int(len(mpos)),
                          "frames_decoded_all_ranks_incl_overlap": int(tot[1]), "payload_ok_all_ranks_incl_overlap": int(tot[0]),
                          "strictly_increasing": bool((np.diff(mpos) > 0).all()), "unmerged_boundaries": int(unmerged), "seconds": dt,
                          "seconds_max_rank_local": mx[2].item() / 1e6, "frames_s": len(mpos) / dt, "msamples_s": n_total / dt / 1e6,
                          "scan_kernel_ms_max_rank": scan_ms, "scan_gbs_aggregate": n_total * 4 / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else None,
                          "stage_ms_rank0": stages, "shards_per_gpu": args.shards, "scaling": "strong",
                          "payloads": "host (pinned)" if args.host_bytes else "device",
                          "note": "wall clock incl. per-rank scan + gather + demod of device-resident int16, NCCL all_gather of the int64 position lists, merge on every rank; payloads stay on their rank"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
