"""Bare pinned H2D || D2H copies on N GPUs at once (no kernels): torchrun --nproc-per-node N pcie_probe_multi.py
Every rank copies the e2e step's bytes (822 MB each way, int16 frames of 32768 frames) on its own GPU, all ranks start
together after a barrier; prints per-rank and aggregate GB/s per direction.  Answers: does the host side of the box scale?"""
import json, os, sys, time
import torch, torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 822_083_584
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.fill_(1); h_out.fill_(2)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps

run(True, True, 1)
res = {}
for name, a, b in (("h2d", 1, 0), ("d2h", 0, 1), ("both", 1, 1)):
    dt = run(a, b)
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]; dist.all_gather(allt, t); ts = [float(x.item()) for x in allt]
    else:
        ts = [dt]
    res[name] = {"per_rank_gbs": [round(n / x / 1e9, 1) for x in ts], "aggregate_gbs_per_direction": round(world * n / max(ts) / 1e9, 1), "ms_max": round(max(ts) * 1e3, 2)}
if rank == 0:
    try:
        aff = len(os.sched_getaffinity(0))
    except Exception:
        aff = None
    print(json.dumps({"probe": "pinned H2D||D2H, no kernels", "n_gpus": world, "bytes_each_way": n, "host_cores": aff, **res}))
if world > 1:
    dist.destroy_process_group()
