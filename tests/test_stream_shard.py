"""Sharding of the streaming receiver across ranks: on CPU with the oracle's rx.cpp loop as the per-rank
engine, the merged list must equal one sequential pass over the whole capture (dense traffic, frames
straddling every block boundary)."""
import numpy as np
import pytest

from cofdm_b200 import stream, synth


@pytest.mark.parametrize("world", [2, 3, 4])
def test_sharded_stream_equals_sequential(oracle_lib, cfg_dir, world):
    o = oracle_lib.Oracle("port", cfg_dir["stream"])         # rx_buf_size = 10: one SDR block = 60 160 samples
    s = o.sizes
    n = 90
    pay = synth.payloads(n, s.usefull_size, seed=17)
    tx16 = np.stack([o.tx(p)[1] for p in pay]).reshape(n, -1, 2)
    rng = np.random.default_rng(23)
    fr = synth.channel(tx16, seed=4, cfo=rng.uniform(-0.003, 0.003, n), phase=rng.uniform(0, 1, n), noise_sigma=1.0)
    cap, _ = synth.capture(fr, gaps=rng.integers(260, 1500, n), noise_sigma=3.0, seed=5, tail=s.output_size * 12)
    blk = stream.block_samples(s)
    cap = cap[: (cap.shape[0] // blk) * blk]
    assert cap.shape[0] // blk >= 2 * world
    want_pos, want_by = o.rx_stream(cap)
    assert len(want_pos) > 60
    pos, by, unmerged = stream.rx_stream_sharded(lambda c: o.rx_stream(c), cap, s, world)
    assert unmerged == 0
    assert pos.tolist() == want_pos.tolist()
    assert np.array_equal(by, want_by)
    # merge_ranges (what the multi-GPU job uses: no per-frame rows, one contiguous slice of every rank's list) gives the same
    lists, rows = [], []
    for r in range(world):
        s0, s1, b0, b1 = stream.shard_slice(cap.shape[0], s, r, world)
        p, b = o.rx_stream(cap[s0:s1])
        lists.append((np.asarray(p, dtype=np.int64) + s0, b0, b1))
        rows.append(b)
    mpos, ranges, um = stream.merge_ranges(lists, s)
    assert um == 0 and mpos.tolist() == want_pos.tolist()
    assert np.array_equal(np.concatenate([rows[r][lo:hi] for r, (lo, hi) in enumerate(ranges)]), want_by)


def _gather_worker(rank, world, port, cfg, q):
    import os
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from cofdm_b200 import dist as cd
    from oracle import oracle as O
    cd.init_from_env(backend="gloo")
    o = O.Oracle("port", cfg)
    s = o.sizes
    n = 60
    pay = synth.payloads(n, s.usefull_size, seed=17)
    tx16 = np.stack([o.tx(p)[1] for p in pay]).reshape(n, -1, 2)
    rng = np.random.default_rng(23)
    fr = synth.channel(tx16, seed=4, cfo=rng.uniform(-0.003, 0.003, n), phase=rng.uniform(0, 1, n), noise_sigma=1.0)
    cap, _ = synth.capture(fr, gaps=rng.integers(260, 1500, n), noise_sigma=3.0, seed=5, tail=s.output_size * 12)
    blk = stream.block_samples(s)
    cap = cap[: (cap.shape[0] // blk) * blk]
    # every rank: the reference loop on its own slice, then the fixed-size all_gather of the position lists and a local merge
    s0, s1, b0, b1 = stream.shard_slice(cap.shape[0], s, rank, world)
    pos, by = o.rx_stream(cap[s0:s1])
    lists = stream.gather_frame_lists(np.asarray(pos, dtype=np.int64) + s0, b0, b1)
    mpos, mtag, unmerged = stream.merge_shards(lists, s)
    mine = mtag[mtag[:, 0] == rank][:, 1]                    # which of MY frames belong to the merged list
    # the one-collective form: fixed-width gather + range merge must agree
    l2 = stream.gather_positions(np.asarray(pos, dtype=np.int64) + s0, b0, b1, max_frames=200)
    mpos2, ranges, um2 = stream.merge_ranges(l2, s)
    assert um2 == unmerged and mpos2.tolist() == mpos.tolist()
    assert mine.tolist() == list(range(*ranges[rank]))
    want_pos, want_by = o.rx_stream(cap) if rank == 0 else (None, None)
    q.put((rank, mpos.tolist(), unmerged, by[mine].tolist(), mpos[mtag[:, 0] == rank].tolist(),
           None if want_pos is None else want_pos.tolist(), None if want_by is None else want_by.tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_gather_of_frame_lists(oracle_lib, cfg_dir):
    """world-size 2 over gloo: the all_gather of int64 position lists that replaces the pickled object gather; both ranks
    arrive at the sequential pass's list and know which of their own payloads belong to it"""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, cfg_dir["stream"], q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_pos, want_by = out[0][5], out[0][6]
    assert len(want_pos) > 40
    got_by = {}
    for rank, mpos, unmerged, my_by, my_pos, _, _ in out:
        assert mpos == want_pos and unmerged == 0
        got_by.update(dict(zip(my_pos, my_by)))
    assert [got_by[p] for p in want_pos] == want_by
