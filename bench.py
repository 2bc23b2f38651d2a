#!/usr/bin/env python
"""bench.py -- headline benchmark of the C-OFDM baseband hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

Metric (BASELINE.json): tx+rx Msamples/s (complex baseband samples, 6016 per frame) and OFDM symbols/s,
with the fused rx kernel's achieved HBM GB/s against the measured peak.

Workload (BASELINE.json configs[2], SURVEY.md section 8d "config 3"): synthetic batched rx of
--frames (default 1 Mi) frames per GPU at the shipped config (fft 512, cp 128, 8 symbols + preamble,
16-QAM), frames produced on the GPU by the tx kernel from seeded payloads and passed through a light
channel (per-frame CFO + phase + AWGN, torch ops, untimed), complex64 in HBM.
One "step" = one tx pass (payload bytes -> frames) + one rx pass (frames -> payload bytes) over the
whole batch.  value = samples through tx and rx per second with everything resident in HBM;
e2e = the same through the C ABI with HOST (pinned) buffers in the SDR's int16 wire format, H2D and D2H
inside the timed region.
The input (tens of GB) is far larger than the 126 MB L2, so no explicit flush is needed.

--impl reference times the reference's own CPU implementation of the same path (oracle/_ref = the
unmodified reference sources + stand-in FFT when built, else the C restatement) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
CONFIG = os.path.join(ROOT, "config", "config.txt")

RX_BYTES_PER_FRAME = 46080 + 1024      # (N+CP)(NS+NPR)*8 read + ND*NS*mod/8 written, 16-QAM (recomputed from the sizes at run time)
TX_BYTES_PER_FRAME = 1024 + 6016 * 8   # payload read + frame written
WORKLOAD = "default"                   # "big": BASELINE.json configs[4] (fft 4096, cp 1024, 1920 + 128 sub-carriers, 64-QAM)
MOD_NAME = {1: "BPSK", 2: "QPSK", 4: "16-QAM", 6: "64-QAM", 8: "256-QAM"}
RX_DRAM_BYTES_PER_FRAME_NCU_BIG = 381789  # big workload: acquire 40992 + 1204, demod 327763 + 11830 (profiles/r02_big_ncu_summary.txt)
RX_DRAM_BYTES_PER_FRAME_NCU = 47361     # measured DRAM read+write of the rx pass (acquire + demod kernels), see roofline.traffic_source


def n_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's implementation of tx + aligned rx on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_run(n_frames_total, threads, seed=0):
    """tx + aligned rx of n_frames_total frames split over `threads` workers; returns (seconds, kind,
    bit_errors).  Each worker owns its own FRAME_FORM pair (the class is not thread-safe)."""
    from oracle import oracle as O
    kind = "reference" if O.available("reference") else "port"
    if kind == "port":
        O.build("port")
    per = max(1, n_frames_total // threads)
    workers = [O.Oracle(kind, CONFIG) for _ in range(threads)]
    s = workers[0].sizes
    rng = np.random.default_rng(seed)
    pay = rng.integers(0, 256, (threads, per, s.usefull_size), dtype=np.uint8)
    errs = [0] * threads

    def work(t):
        errs[t] = workers[t].txrx_loop(pay[t])          # whole loop inside the C library (no GIL held)

    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    dt = time.perf_counter() - t0
    return dt, kind, per * threads, sum(errs), s


def workload_name(s, native=False):
    geo = f"fft{s.fft_size} cp{s.cp_size} {s.num_symb}sym+preamble {s.num_data_subc}+{s.num_pilot_subc} sub-carriers " + MOD_NAME.get(s.mod_type, "?")
    which = "default config.txt" if WORKLOAD == "default" else "BASELINE configs[4] large-FFT dense-pilot config"
    if native:
        return f"synthetic batched tx + fused aligned rx, {which} ({geo}), complex64 in HBM"
    return f"batched tx + aligned rx, {which} ({geo})"


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = n_cores()
    frames = args.cpu_frames or (4000 if WORKLOAD == "default" else 500) * cores
    times = []
    for _ in range(args.warmup):
        cpu_run(max(cores, frames // 8), cores)
    n_done = 0
    for _ in range(args.steps):
        dt, kind, n_done, errs, s = cpu_run(frames, cores)
        times.append(dt)
    dt = float(np.mean(times))
    samples = 2 * n_done * s.output_size
    v = samples / dt / 1e6
    line = {"impl": "reference", "metric": "tx+rx Msamples/s", "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(s), "frames_per_step": n_done, "frame_samples": s.output_size},
            "ofdm_symbols_s": 2 * n_done * (s.num_symb + s.num_pr_symb) / dt,
            "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": kind,
                             "sample": f"{n_done} frames tx+rx per step on {cores} threads; FFT = stand-in mixed-radix (FFTW3 not installed)"},
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def oracle_subset_check(rx_in, out_bytes, s, n, seed):
    """n seeded random frames of the timed rx batch through the CPU oracle; returns counts"""
    import torch
    from oracle import oracle as O
    kind = "reference" if O.available("reference") else "port"
    if kind == "port":
        O.build("port")
    o = O.Oracle(kind, CONFIG)
    idx = np.sort(np.random.default_rng(seed).choice(rx_in.shape[0], size=n, replace=False))
    ti = torch.from_numpy(idx).to(rx_in.device)
    recs = rx_in[ti][:, s.t2sin_size:s.t2sin_size + s.rx_len].cpu().numpy().astype(np.complex128)
    got = out_bytes[ti].cpu().numpy()
    mod = s.mod_type
    L = 1 << (mod // 2)
    equal = amb_diff = unexplained = 0
    for i in range(n):
        r = o.rx_aligned(recs[i])
        if np.array_equal(got[i], r["bytes"]):
            equal += 1
            continue
        bits_g, bits_w = np.unpackbits(got[i]), np.unpackbits(np.asarray(r["bytes"], dtype=np.uint8))
        k = len(bits_g) // mod
        w = 1 << np.arange(mod - 1, -1, -1)
        diff = (bits_g[:k * mod].reshape(k, mod) @ w) != (bits_w[:k * mod].reshape(k, mod) @ w)
        p = np.asarray(r["constell"]).ravel()[:k]
        if mod == 1:
            amb = np.abs(p.real + p.imag) < 2e-4
        else:
            amb = np.zeros(k, bool)
            for v in (p.real, p.imag):
                u = (np.clip(v, -1, 1) + 1) * (L - 1) / 2 + 0.5
                rr = np.rint(u)
                amb |= (np.abs(u - rr) < 2e-4) & (rr >= 1) & (rr <= L - 1)
        if np.any(diff & ~amb):
            unexplained += 1
        else:
            amb_diff += 1
    return {"oracle": kind, "frames": int(n), "seed": seed, "frames_bytes_equal": equal, "frames_differing_only_at_oracle_ambiguous_symbols": amb_diff,
            "frames_differing_elsewhere": unexplained}


# ---------------------------------------------------------------------------------------------------
def native_arm(args):
    import torch
    import torch.distributed as dist
    import cofdm_b200 as cb
    from cofdm_b200 import dist as cd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device: there is no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rank, world, local = cd.init_from_env(backend="nccl", device=dev)

    m = cb.Modem(CONFIG, device=local)
    m.use_torch_stream()
    s = m.sizes
    F = args.frames
    global RX_BYTES_PER_FRAME, TX_BYTES_PER_FRAME
    RX_BYTES_PER_FRAME = s.rx_len * 8 + s.usefull_size           # every rx sample read once + the payload written
    TX_BYTES_PER_FRAME = s.usefull_size + s.output_size * 8
    n_sym = s.num_symb + s.num_pr_symb
    # ---- workload: payloads -> tx kernel -> light channel (untimed) -----------------------------------
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    payload = torch.randint(0, 256, (F, s.usefull_size), dtype=torch.uint8, device=dev, generator=g)
    rx_in = torch.empty((F, s.output_size), dtype=torch.complex64, device=dev)
    tx_out = torch.empty((F, s.output_size), dtype=torch.complex64, device=dev)
    m.tx_batch(payload, cb.CF32, out=rx_in)
    n = torch.arange(s.output_size, device=dev, dtype=torch.float64)
    CH = max(64, 8192 * 6016 // s.output_size)
    # CFO range: well inside the coarse estimator's window (+-10 bins of the preamble-length grid); noise on the int16 grid
    cfo_max, noise_sigma = (0.003, 1.5) if WORKLOAD == "default" else (0.0005, 0.5)
    for f0 in range(0, F, CH):
        f1 = min(F, f0 + CH)
        cfo = (torch.rand((f1 - f0, 1), device=dev, generator=g, dtype=torch.float64) - 0.5) * (2 * cfo_max)
        ph = torch.rand((f1 - f0, 1), device=dev, generator=g, dtype=torch.float64)
        rot = torch.polar(torch.ones_like(cfo * n), 2 * torch.pi * (cfo * n + ph)).to(torch.complex64)
        blk = rx_in[f0:f1] * rot * float(s.mult)
        noise = torch.randn((f1 - f0, s.output_size, 2), device=dev, generator=g) * noise_sigma
        rx_in[f0:f1] = torch.view_as_complex(torch.round(torch.view_as_real(blk) + noise))   # int16-grid samples, like an ADC
    out_bytes = torch.empty((F, s.usefull_size), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def step():
        m.tx_batch(payload, cb.CF32, out=tx_out)
        m.rx_aligned_batch(rx_in, n_frames=F, frame_stride=s.output_size, offset=s.t2sin_size, out=out_bytes, count_ambiguous=False)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = m.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps + 1)]
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    ev[0].record()
    for k in range(args.steps):
        m.tx_batch(payload, cb.CF32, out=tx_out)
        ev[3 * k + 1].record()
        m.rx_aligned_batch(rx_in, n_frames=F, frame_stride=s.output_size, offset=s.t2sin_size, out=out_bytes, count_ambiguous=False)
        ev[3 * k + 2].record()
        ev[3 * k + 3].record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    launches = m.launch_count() - l0
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[3 * args.steps])
    tx_ms = float(np.mean([ev[3 * k].elapsed_time(ev[3 * k + 1]) for k in range(args.steps)]))
    rx_ms = float(np.mean([ev[3 * k + 1].elapsed_time(ev[3 * k + 2]) for k in range(args.steps)]))
    # correctness of what was timed: decoded bytes vs payload (bit errors), boundary-ambiguous symbols
    _, amb = m.rx_aligned_batch(rx_in, n_frames=min(F, 65536), frame_stride=s.output_size, offset=s.t2sin_size, out=out_bytes[:min(F, 65536)])
    # what the optional count of boundary-ambiguous decisions costs (it is off inside the timed region): one rx pass with it on
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    m.rx_aligned_batch(rx_in, n_frames=F, frame_stride=s.output_size, offset=s.t2sin_size, out=out_bytes, count_ambiguous=True)
    eb.record()
    torch.cuda.synchronize()
    rx_ms_amb = ea.elapsed_time(eb)
    m.rx_aligned_batch(rx_in, n_frames=F, frame_stride=s.output_size, offset=s.t2sin_size, out=out_bytes, count_ambiguous=False)
    diff = (out_bytes ^ payload)
    bit_err = int(torch.sum(torch.bitwise_count(diff).to(torch.int64)).item()) if hasattr(torch, "bitwise_count") else int((diff != 0).sum().item())
    frames_bad = int((diff != 0).any(dim=1).sum().item())

    # a seeded random subset of the timed batch against the oracle (rank 0, CPU, after the timed region):
    # bytes must be equal except at symbols the ORACLE itself places within 2e-4 level units of a decision boundary
    oracle_check = None
    if rank == 0 and args.oracle_frames > 0:
        oracle_check = oracle_subset_check(rx_in, out_bytes, s, min(args.oracle_frames, F), seed=4321)

    # the only collective of the job: SUM of 4 counters + MAX of the device times, over NCCL
    (bit_err, frames_bad, frames_all, amb), (total_ms, tx_ms, rx_ms) = cd.reduce_results(
        [bit_err, frames_bad, F, amb], [total_ms, tx_ms, rx_ms], device=dev)

    # ---- end to end through the C ABI with HOST (pinned) buffers; H2D + D2H inside the timed region ----
    # Host-side samples use the SDR wire format (int16 I,Q: FRAME_FORM::get_int16 / from_sdr_int16_buf), which
    # is what the reference apps and the CPU arm move between modem and radio.  The tx pass (payload -> frames)
    # and the rx pass (frames -> payload) of a step are independent, so they run concurrently on two handles
    # (two host threads, two streams each): PCIe is full duplex and tx is D2H-heavy, rx H2D-heavy.
    E = min(args.e2e_frames, F)
    m2 = cb.Modem(CONFIG, device=local)
    h_pay = torch.empty((E, s.usefull_size), dtype=torch.uint8).pin_memory()
    h_frames = torch.empty((E, s.output_size, 2), dtype=torch.int16).pin_memory()
    h_rx = torch.empty((E, s.output_size, 2), dtype=torch.int16).pin_memory()
    h_out = torch.empty((E, s.usefull_size), dtype=torch.uint8).pin_memory()
    h_pay.copy_(payload[:E])
    h_rx.copy_(torch.view_as_real(rx_in[:E]).to(torch.int16))       # the channel output already sits on the int16 grid
    np_pay, np_frames, np_rx, np_out = h_pay.numpy(), h_frames.numpy(), h_rx.numpy(), h_out.numpy()
    m.use_own_stream()

    def e2e_step():
        t = threading.Thread(target=lambda: m2.tx_batch(np_pay, cb.CI16, out=np_frames))
        t.start()
        m.rx_aligned_batch(np_rx, n_frames=E, frame_stride=s.output_size, offset=s.t2sin_size, out=np_out, count_ambiguous=False)
        t.join()

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    if world > 1:
        dist.barrier()
    e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_step()
    e_dt = (time.perf_counter() - t0) / e_steps
    e_bad = int((np_out != np_pay).any(axis=1).sum())
    # the frames the concurrent tx call wrote to host memory: equal to a device-resident tx of the same payload
    tx_dev = m.tx_batch(payload[:E], cb.CI16).cpu().numpy().reshape(np_frames.shape)
    e_tx_bad = int((tx_dev != np_frames).any(axis=(1, 2)).sum())
    e_ok = True if e_bad == 0 and e_tx_bad == 0 else f"{e_bad} rx frames, {e_tx_bad} tx frames differ"
    _, (e_dt,) = cd.reduce_results([], [e_dt], device=dev)
    e2e_value = world * 2 * E * s.output_size / e_dt / 1e6
    h2d = E * s.usefull_size + E * s.output_size * 4
    d2h = E * s.output_size * 4 + E * s.usefull_size

    if rank == 0:
        peak, peak_src = load_peaks()
        step_ms = total_ms / args.steps
        samples_per_step = world * 2 * F * s.output_size
        value = samples_per_step / (step_ms * 1e-3) / 1e6
        rx_gbs = RX_BYTES_PER_FRAME * F / (rx_ms * 1e-3) / 1e9
        tx_gbs = TX_BYTES_PER_FRAME * F / (tx_ms * 1e-3) / 1e9
        line = {
            "metric": "tx+rx Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(s, native=True),
                       "frames_per_gpu": F, "frame_samples": s.output_size, "l2": "inputs (GBs) far larger than the 126 MB L2, no flush needed",
                       "channel": f"per-frame CFO +-{cfo_max} cyc/sample, random phase, AWGN sigma {noise_sigma} LSB, int16 grid"},
            "ofdm_symbols_s": world * 2 * F * n_sym / (step_ms * 1e-3),
            "rx_msamples_s": world * F * s.output_size / (rx_ms * 1e-3) / 1e6,
            "tx_msamples_s": world * F * s.output_size / (tx_ms * 1e-3) / 1e6,
            "rx_frames_s": world * F / (rx_ms * 1e-3), "rx_ms": rx_ms, "tx_ms": tx_ms,
            "rx_msamples_s_on_rx_len": world * F * s.rx_len / (rx_ms * 1e-3) / 1e6,
            "bit_errors": bit_err, "frames_with_errors": frames_bad, "frames_checked": frames_all, "oracle_check": oracle_check,
            "boundary_ambiguous_symbols_in_first_64k_frames_per_gpu": amb, "rx_ms_with_ambiguity_count_on": rx_ms_amb,
            "roofline": {"bound": "hbm",
                         "kernel": ("rx pass = rx_acquire512w_kernel + rx_demod512_kernel" if WORKLOAD == "default" else "rx pass = big_acquire_kernel + big_demod_kernel (cluster of 8 CTAs per frame)")
                                   + " (together they read every sample exactly once)",
                         "achieved": rx_gbs, "peak": peak, "unit": "GB/s", "frac": rx_gbs / peak,
                         "traffic": (RX_DRAM_BYTES_PER_FRAME_NCU * F if (WORKLOAD == "default" and s.mod_type == 4) else
                                     (RX_DRAM_BYTES_PER_FRAME_NCU_BIG * F if WORKLOAD == "big" else None)),
                         "peak_source": peak_src,
                         "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of the two rx kernels "
                                           + ("over 32768 frames (profiles/r02_final_ncu_summary.txt) = 47361 B/frame" if WORKLOAD == "default"
                                              else "over 4096 frames (profiles/r02_big_ncu_summary.txt) = 381789 B/frame") + ", scaled to this launch's frames",
                         "algorithmic_bytes_per_frame": RX_BYTES_PER_FRAME, "tx_kernel_gbs": tx_gbs, "tx_frac": tx_gbs / peak},
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "frames_per_step": E, "ms_per_step": e_dt * 1e3, "payload_roundtrip_ok": e_ok, "host_sample_format": "ci16 (SDR wire format)",
                    "path": "cofdm_tx_batch || cofdm_rx_aligned_batch, COFDM_HOST pinned buffers, two handles run concurrently (double-buffered chunked H2D/kernel/D2H each)"},
            "gpu_launches": launches, "wall_ms_per_step": t_wall / args.steps * 1e3, "clocks": clocks,
        }
        if not args.no_cpu and world == 1:
            cores = n_cores()
            cf = args.cpu_frames or (10000 if WORKLOAD == "default" else 1200) * cores
            cpu_run(max(cores, cf // 8), cores)
            dt, kind, n_done, errs, cs = cpu_run(cf, cores)
            dt1, _, n1, errs1, _ = cpu_run(2000 if WORKLOAD == "default" else 250, 1)                       # (i) one thread, (ii) every core
            line["cpu_baseline"] = {"value": 2 * n_done * cs.output_size / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind,
                                    "sample": f"{n_done} frames tx+rx on {cores} threads in {dt:.1f} s; FFT = stand-in mixed-radix (FFTW3 not installed)",
                                    "payload_bytes_wrong": errs + errs1,
                                    "one_thread": {"value": 2 * n1 * cs.output_size / dt1 / 1e6, "unit": "Msamples/s",
                                                   "us_per_frame_tx_plus_rx": dt1 / n1 * 1e6, "sample": f"{n1} frames on 1 thread in {dt1:.1f} s"},
                                    "reference_authors_own_figure": "LOG.txt: 238 us per received frame with FFTW3 on the author's CPU (rx only)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="default", choices=["default", "big"],
                    help="default: BASELINE configs[2] (the headline, shipped config.txt); big: configs[4] (fft 4096, cp 1024, 1920 + 128 sub-carriers, 64-QAM)")
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (default 1 Mi; big workload: 16384 = 6 GB of samples)")
    ap.add_argument("--e2e-frames", type=int, default=0)
    ap.add_argument("--cpu-frames", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--oracle-frames", type=int, default=1024, help="frames of the timed batch re-decoded by the oracle (0: skip)")
    ap.add_argument("--mod-type", type=int, default=0, choices=[0, 1, 2, 4, 6, 8],
                    help="0: the shipped config.txt (16-QAM, the headline workload); otherwise the same geometry with this modType "
                         "(BASELINE configs[2] also names QPSK)")
    args = ap.parse_args()
    global WORKLOAD, CONFIG
    WORKLOAD = args.workload
    if not args.frames:
        args.frames = (1 << 20) if WORKLOAD == "default" else 16384
    if not args.e2e_frames:
        args.e2e_frames = (1 << 15) if WORKLOAD == "default" else 4096
    if WORKLOAD == "big":
        import tempfile
        from cofdm_b200 import synth
        rank = os.environ.get("RANK", "0")
        CONFIG = synth.write_config(os.path.join(tempfile.mkdtemp(), f"config_big_{rank}.txt"), fft_size=4096, cp_size=1024, num_data_subc=1920,
                                    num_pilot_subc=128, num_symb=8, pr_sin_len=128, modType=args.mod_type or 6)
        args.oracle_frames = min(args.oracle_frames, 64)
        args.mod_type = 0
    if args.mod_type:
        import tempfile
        from cofdm_b200 import synth
        rank = os.environ.get("RANK", "0")
        CONFIG = synth.write_config(os.path.join(tempfile.mkdtemp(), f"config_mod{args.mod_type}_{rank}.txt"), modType=args.mod_type)
    if args.impl == "reference":
        reference_arm(args)
    else:
        native_arm(args)


if __name__ == "__main__":
    main()
