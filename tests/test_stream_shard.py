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
