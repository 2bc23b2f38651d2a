"""Multi-GPU plumbing: frames are independent, so a batch shards over ranks with NO data-path
collective; the only communication is one all-reduce of a handful of 64-bit counters (bit errors,
frames, ...) and a MAX of the elapsed device time -- tens of bytes over NCCL/NVLink (gloo on CPU in
the tests).  SURVEY.md section 8(e)."""
import os


def shard_range(n_total, rank, world):
    """contiguous [begin, end) of the frame index range owned by `rank`; sizes differ by at most 1"""
    base, rem = divmod(int(n_total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def init_from_env(backend=None, device=None):
    """init torch.distributed from the torchrun environment (RANK/WORLD_SIZE/MASTER_*); no-op for 1 rank.
    Returns (rank, world, local_rank)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {"device_id": device} if (device is not None and backend == "nccl") else {}
        dist.init_process_group(backend or "nccl", rank=rank, world_size=world, **kw)
    return rank, world, local


def reduce_results(counters, times_ms, device=None):
    """counters: list of ints (summed over ranks); times_ms: list of floats (max over ranks: device time
    of the slowest rank defines the job).  Returns (counters, times_ms) identical on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(c) for c in counters], [float(t) for t in times_ms]
    c = torch.tensor([int(x) for x in counters], dtype=torch.int64, device=device)
    t = torch.tensor([float(x) for x in times_ms], dtype=torch.float64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [int(x) for x in c.tolist()], [float(x) for x in t.tolist()]
