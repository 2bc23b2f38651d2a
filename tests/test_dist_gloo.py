"""N>1 host logic on CPU: world_size-2 gloo run of the sharding + counter/time reduction that bench.py
uses (frames shard with no data-path collective; one SUM of counters and one MAX of device times)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from cofdm_b200 import dist as cd


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            parts = [cd.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    r, w, _ = cd.init_from_env(backend="gloo")
    b, e = cd.shard_range(1001, r, w)
    # each rank "decodes" its shard: counters = (bit errors, frames, ambiguous), time = its own device time
    counters, times = cd.reduce_results([3 * r + 1, e - b, 7], [10.0 + 5 * r, 1.0])
    q.put((r, counters, times))
    dist.destroy_process_group()


def test_two_rank_gloo_reduction():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r, counters, times in out:
        assert counters == [1 + 4, 1001, 14]        # SUM over ranks; the shards cover every frame once
        assert times == [15.0, 1.0]                 # MAX over ranks


def test_single_rank_is_identity():
    assert cd.reduce_results([1, 2], [3.5]) == ([1, 2], [3.5])
