#!/usr/bin/env python
"""BASELINE.json configs[3] (stream sync over N GPUs): a long int16 capture with frames at random gaps, cut into
contiguous ranges of whole SDR blocks (one per rank, + one overlap block), every rank scanning and demodulating its
range on its own GPU (itself split into CTA-sized shards), rank 0 merging the lists.  No data-path collective: the
only communication is the gather of the frame lists.
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/bench_stream_dist.py
"""
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
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = cb.Modem(os.path.join(ROOT, "config", "config.txt"), device=local)
    m.use_torch_stream()
    s = m.sizes
    nfr, reps = 4000, 8 * world                                  # weak scaling: 32 000 frames per GPU
    pay = synth.payloads(nfr, s.usefull_size, seed=5)
    fr = m.tx_batch(pay, cb.CI16)
    rng = np.random.default_rng(3)
    cap, starts = synth.capture(fr[..., 0].astype(np.float64) + 1j * fr[..., 1], gaps=rng.integers(300, 2500, nfr), noise_sigma=3.0, seed=4,
                           tail=s.output_size * 41)
    blk = st.block_samples(s)
    cap = cap[: cap.shape[0] // blk * blk]
    n_total = cap.shape[0] * reps
    # every rank materialises only its own slice (+ overlap) of the `reps`-fold capture on its GPU
    s0, s1, b0, b1 = st.shard_slice(n_total, s, rank, world)
    idx = np.arange(s0, s1) % cap.shape[0]
    mine = torch.from_numpy(cap[idx]).cuda()
    starts_pr = None

    def one_pass():
        """per rank: scan + demodulate its slice; then the lists (positions + (rank, index) tags, not the payloads) go to
        rank 0, which merges the chains.  The payloads stay on the rank that decoded them."""
        pos, by = m.rx_stream(mine, shards=296)
        tag = np.stack([np.full(len(pos), rank, np.int64), np.arange(len(pos), dtype=np.int64)], axis=1)
        item = (pos + s0, tag, b0, b1)
        if world == 1:
            return pos, by, st.merge_shards([item], s)
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(item, gathered, dst=0)
        return pos, by, (st.merge_shards(gathered, s) if rank == 0 else None)

    for _ in range(2):
        one_pass()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    pos, by, merged = one_pass()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    # every payload this rank decoded, against what was sent: the frame at capture offset r carries payload j(r)
    r = (pos + s0) % cap.shape[0]
    j = np.clip(np.searchsorted(starts + s.t2sin_size - 64, r) - 1, 0, nfr - 1)
    near = np.abs(r - (starts[j] + s.t2sin_size)) < 64
    good = int((near & (by == pay[j]).all(axis=1)).sum())
    tot = torch.tensor([good, len(pos)], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tot)
    if rank == 0:
        mpos, mtag, unmerged = merged
        print(json.dumps({"n_gpus": world, "capture_samples": int(n_total), "frames_sent": reps * nfr, "frames_after_merge": int(len(mpos)),
                          "frames_decoded_all_ranks_incl_overlap": int(tot[1]), "payload_ok_all_ranks_incl_overlap": int(tot[0]),
                          "strictly_increasing": bool((np.diff(mpos) > 0).all()), "unmerged_boundaries": int(unmerged), "seconds": dt,
                          "frames_s": len(mpos) / dt, "msamples_s": n_total / dt / 1e6, "shards_per_gpu": 296, "scaling": "weak",
                          "note": "wall clock: per-rank scan + demod of device-resident int16, gather of the frame lists (positions + tags), merge on rank 0; payloads stay on their rank"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
