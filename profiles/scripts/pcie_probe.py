import torch, time, threading
n = 822_083_584
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
run(True, True, 1)
for name, a, b in (("h2d", 1, 0), ("d2h", 0, 1), ("both", 1, 1)):
    dt = run(a, b); print(name, f"{dt*1e3:.2f} ms  {n/dt/1e9:.1f} GB/s per direction")
