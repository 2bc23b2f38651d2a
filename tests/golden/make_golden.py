"""Generate the committed golden fixtures.  Run ONLY in the build container (needs /root/reference
and oracle/_ref/libcofdm_ref.so = the unmodified reference sources):

    python tests/golden/make_golden.py

  ref_capture.npz  the reference's own recorded artefacts (reference data/*.bin, data.txt): the
                   tx frame of main.cpp:74, the 246 656-sample PlutoSDR capture (integer valued ->
                   stored as int16), and the three dumps of main.cpp:76-78.
  ref_vectors.npz  outputs of the compiled reference on seeded synthetic inputs for every modType:
                   tx frames, the full aligned rx chain with all taps, find_corr, chan_char, read,
                   mod/demod incl. exact decision-boundary ties, and the rx.cpp-style stream loop
                   (on a config with rx_buf_size = 10 so that one SDR block is 60 160 samples).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cofdm_b200  # noqa: E402,F401  (registers the package)
from cofdm_b200 import synth  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    # ---- the reference's recorded artefacts ---------------------------------------------------------
    data = np.fromfile(f"{REF}/data/data.bin", dtype=np.float64)
    assert np.all(data == np.round(data)) and np.abs(data).max() < 32768
    txt = open(f"{REF}/WARANDPEACE.txt", "rb").read()
    np.savez_compressed(
        f"{OUT}/ref_capture.npz",
        source_i16=np.fromfile(f"{REF}/data/source.bin", dtype=np.int16),
        capture_i16=data.astype(np.int16).reshape(-1, 2),
        t2_sin_corr=np.fromfile(f"{REF}/data/t2_sin_corr.bin", dtype=np.float64),
        phases=np.fromfile(f"{REF}/data/phases.bin", dtype=np.float64).view(np.complex128),
        constell=np.fromfile(f"{REF}/data/constell.bin", dtype=np.float64).view(np.complex128),
        data_txt=np.frombuffer(open(f"{REF}/data.txt", "rb").read(), dtype=np.uint8),
        mac_frame=np.frombuffer(bytes([1, 0, 0, 0, 0, 0, 126, 87]) + txt[:248], dtype=np.uint8),
    )

    # ---- compiled-reference outputs on seeded inputs ----------------------------------------------
    vec = {}
    tmp = "/tmp/cofdm_golden_cfg.txt"
    for mt in (1, 2, 4, 6, 8):
        synth.write_config(tmp, base=f"{REF}/config/config.txt", modType=mt)
        R = Oracle("reference", tmp)
        s = R.sizes
        rng = np.random.default_rng(100 + mt)
        pay = synth.payloads(2, s.usefull_size, seed=mt)
        vec[f"m{mt}_payload"] = pay
        tx16 = np.stack([R.tx(p)[1] for p in pay]).reshape(2, -1, 2)
        vec[f"m{mt}_tx_i16"] = tx16
        vec[f"m{mt}_tx_frame0"] = R.tx(pay[0])[0]
        # impaired frames: CFO, phase, 3-tap multipath, AWGN, slightly early timing
        rx = synth.channel(tx16, seed=mt, cfo=[0.0007, -0.0021], phase=[0.1, 0.45],
                           taps=(1.0, 0.2 - 0.1j, 0.05j), noise_sigma=1.5)
        off = np.array([0, 2])
        rec = np.stack([rx[i, s.t2sin_size - off[i]: s.t2sin_size - off[i] + s.preamble_size + s.message_size] for i in range(2)])
        vec[f"m{mt}_rx_in_i16"] = synth.to_i16(rec)
        outs = [R.rx_aligned(r) for r in rec]
        for k in ("scal", "synced", "grid", "chan", "constell", "bytes"):
            if k in ("synced", "grid") and mt != 4:
                continue                                   # the big taps are kept for the default modType only
            vec[f"m{mt}_rx_{k}"] = np.stack([o[k] for o in outs])
        vec[f"m{mt}_chan_char"] = R.chan_char(rec[0][: s.preamble_size])
        b, restored = R.read(R.tx(pay[1])[0])
        vec[f"m{mt}_read_bytes"], vec[f"m{mt}_read_restored"] = b, restored
        # mod / demod known answers, including points exactly on decision boundaries
        raw = rng.integers(0, 256, 97, dtype=np.uint8)
        vec[f"m{mt}_mod_in"], vec[f"m{mt}_mod_out"] = raw, R.mod(mt, raw)
        L = 1 << (mt // 2) if mt > 1 else 2
        grid = np.linspace(-1.5, 1.5, 61)
        ties = np.array([-1 + (2 * k + 1) / (L - 1) for k in range(L - 1)]) if mt > 1 else np.array([0.0])
        pts = np.concatenate([rng.normal(0, 0.7, 400) + 1j * rng.normal(0, 0.7, 400),
                              (grid[:, None] + 1j * grid[None, :]).ravel(),
                              (ties[:, None] + 1j * ties[None, :]).ravel(), np.array([1e-17 - 1e-17j, 0j])])
        pts = pts[: len(pts) // 8 * 8]
        db, clamped = R.demod(mt, pts)
        vec[f"m{mt}_demod_in"], vec[f"m{mt}_demod_out"], vec[f"m{mt}_demod_clamped"] = pts, db, clamped
        R.close()
    # sync functions + stream loop on a synthetic capture (default modType)
    synth.write_config(tmp, base=f"{REF}/config/config.txt", rx_buf_size=10)   # small SDR block: keeps the fixture small
    R = Oracle("reference", tmp)
    s = R.sizes
    pay = synth.payloads(6, s.usefull_size, seed=77)
    tx16 = np.stack([R.tx(p)[1] for p in pay]).reshape(6, -1, 2)
    fr = synth.channel(tx16, seed=5, cfo=np.linspace(-0.003, 0.003, 6), phase=np.linspace(0, 1, 6), noise_sigma=1.0)
    cap, starts = synth.capture(fr, gaps=[1500, 300, 4097, 260, 9000, 777], noise_sigma=3.0, seed=9, tail=s.output_size * 10)
    cap = cap[: s.output_size * s.rx_buf_size]                       # one SDR block
    capc = cap[:, 0].astype(np.float64) + 1j * cap[:, 1]
    vec["sync_capture_i16"], vec["sync_frame_starts"], vec["sync_payload"] = cap, starts, pay
    vec["sync_t2corr"] = R.t2sin_corr(capc)
    t2 = [R.find_t2sin(capc, int(st)) for st in (0, 1000, 7000, 20000)]
    vec["sync_find_t2_starts"], vec["sync_find_t2"] = np.array([0, 1000, 7000, 20000]), np.array(t2)
    pst = np.array([t for t in t2 if t >= 0] + [0, 12345])
    vec["sync_pre_starts"] = pst
    vec["sync_find_pre"] = np.array([R.find_preamble(capc, int(p)) for p in pst])
    vec["sync_find_corr"] = np.stack([R.find_corr(capc, int(p)) for p in pst])
    pos, by = R.rx_stream(cap)
    vec["sync_stream_pos"], vec["sync_stream_bytes"] = pos, by
    np.savez_compressed(f"{OUT}/ref_vectors.npz", **vec)
    print("frames found by the reference stream loop:", pos, "payload ok:", [(b == p).all() for b, p in zip(by, pay)])
    for f in ("ref_capture.npz", "ref_vectors.npz"):
        print(f, os.path.getsize(f"{OUT}/{f}") // 1024, "KiB")


if __name__ == "__main__":
    main()
