"""Parity assertions shared by the emulator tests (CPU, `not gpu`) and the GPU tests (`gpu`, through
the C ABI).  `m` is a cofdm_b200.Modem or an EmuModem; `o` is the oracle (oracle.Oracle) of the same
configuration.  Tolerances (north star): complex samples within 1e-5 relative L2 of the reference's
fp64 values (the kernels compute in fp32); bits exact, except symbols the oracle itself places within
AMBIG_MARGIN of a decision boundary, which are counted and reported, never silently accepted."""
import numpy as np

from cofdm_b200 import synth

TOL = 1e-5            # relative L2, fp32 kernels vs fp64 oracle
AMBIG_MARGIN = 2e-4   # level units, same as kAmbigMargin in c-ofdm_b200/csrc/modem.cuh


def to_np(x):
    return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)


def rel_l2(a, b):
    a, b = to_np(a).astype(np.complex128).ravel(), to_np(b).astype(np.complex128).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def cplx(i16):
    i16 = to_np(i16)
    return i16[..., 0].astype(np.float64) + 1j * i16[..., 1].astype(np.float64)


def ambiguous_symbols(points, mod):
    """bool mask: oracle points within AMBIG_MARGIN of a hard-decision boundary (modulation.cpp:62,75-78)"""
    p = np.asarray(points)
    if mod == 1:
        return np.abs(p.real + p.imag) < AMBIG_MARGIN
    L = 1 << (mod // 2)
    out = np.zeros(p.shape, bool)
    for v in (p.real, p.imag):
        u = (np.clip(v, -1, 1) + 1) * (L - 1) / 2 + 0.5
        r = np.rint(u)
        out |= (np.abs(u - r) < AMBIG_MARGIN) & (r >= 1) & (r <= L - 1)
    return out


def symbols_of(bytes_, mod):
    bits = np.unpackbits(to_np(bytes_).astype(np.uint8).ravel())
    n = len(bits) // mod
    return bits[: n * mod].reshape(n, mod) @ (1 << np.arange(mod - 1, -1, -1))


def assert_bytes_match(got, want, oracle_points, mod, what=""):
    """bit-exact except at oracle-ambiguous symbols; returns the number of such symbols that differ"""
    got, want = to_np(got).ravel(), to_np(want).ravel()
    if np.array_equal(got, want):
        return 0
    sg, sw = symbols_of(got, mod), symbols_of(want, mod)
    diff = sg != sw
    amb = ambiguous_symbols(np.asarray(oracle_points).ravel()[: len(diff)], mod)
    assert not np.any(diff & ~amb), f"{what}: {int(np.sum(diff & ~amb))} symbol decisions differ away from any boundary"
    return int(np.sum(diff))


# ---------------------------------------------------------------------------------------------------
def check_mod_demod(m, o, mod, seed=0):
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, 1023, dtype=np.uint8)
    pts = m.mod(raw, mod)
    ref = o.mod(mod, raw)
    assert np.array_equal(to_np(pts), ref.astype(np.complex64)), "QAM map must be the fp32 rounding of the reference table"
    L = 1 << (mod // 2) if mod > 1 else 2
    ties = np.array([-1 + (2 * k + 1) / (L - 1) for k in range(L - 1)]) if mod > 1 else np.array([0.0])
    p = np.concatenate([rng.normal(0, 0.7, 4000) + 1j * rng.normal(0, 0.7, 4000),
                        (ties[:, None] + 1j * ties[None, :]).ravel()])
    p = p[: len(p) // 8 * 8].astype(np.complex64)           # the inputs themselves are fp32 here
    got, n_amb = m.demod(p, mod)
    want, _ = o.demod(mod, p.astype(np.complex128))
    nd = assert_bytes_match(got, want, p.astype(np.complex128), mod, "demod")
    assert n_amb >= nd
    # round trip: demod(mod(x)) == x
    back, _ = m.demod(m.mod(raw, mod), mod)
    assert np.array_equal(to_np(back)[: len(raw)], raw)
    return dict(ambiguous=n_amb, differing=nd)


def check_tx(m, o, n_frames=3, seed=1):
    s = o.sizes
    pay = synth.payloads(n_frames, s.usefull_size, seed=seed)
    f32 = to_np(m.tx_batch(pay, 0))
    i16 = to_np(m.tx_batch(pay, 1))
    worst, flips = 0.0, 0
    for i in range(n_frames):
        ref, q = o.tx(pay[i])
        worst = max(worst, rel_l2(f32[i], ref))
        d = i16[i].reshape(-1).astype(np.int32) - q.astype(np.int32)
        # int16 = trunc(x*mult): an fp32 sample may sit on the other side of an integer than the fp64 one -- but only where the
        # fp64 value lies closer to that integer than the fp32 error of this very frame (measured on its cf32 form, x2 margin)
        assert np.abs(d).max() <= 1
        frac = np.abs(np.stack([ref.real, ref.imag], -1).reshape(-1) * s.mult)
        err = np.abs(f32[i].astype(np.complex128) - ref)
        window = 2.0 * s.mult * float(max(err.real.max(), err.imag.max(), np.abs(err).max()))
        assert window < 1e-3, window
        near = np.abs(frac - np.rint(frac)) < window
        assert not np.any((d != 0) & ~near), "int16 frame differs away from a truncation boundary"
        flips += int(np.count_nonzero(d))
    assert worst < TOL, worst
    return dict(rel_l2=worst, int16_boundary_flips=flips)


def coarse_cfo_near_tie(rec_c, s, rel_gap=1e-5):
    """pilot_freq_sinh (Frame.hpp:285-337) restated in numpy fp64 on one received preamble: True when, in at least one of
    the arg-max windows, the runner-up |X| lies within `rel_gap` of the maximum -- the only situation in which an fp32
    spectrum may legitimately pick another bin than the reference's fp64 one"""
    size = s.preamble_size
    x = np.fft.fftshift(np.abs(np.fft.fft(np.asarray(rec_c[:size], dtype=np.complex128))))
    rel_bw = (s.num_data_subc + s.num_pilot_subc) / s.fft_size
    rel_pw = rel_bw / s.num_pilot_subc
    w = int(size * rel_pw)
    b0 = int((1.0 - rel_bw - rel_pw) / 2.0 * size)
    for win in range(s.num_pilot_subc + 1):
        if win == s.num_pilot_subc // 2:
            continue
        seg = np.sort(x[max(0, b0 + win * w): b0 + (win + 1) * w])
        if seg[-1] > 0 and (seg[-1] - seg[-2]) <= rel_gap * seg[-1]:
            return True
    return False


def impaired_records(o, n_frames, seed, cfo_max=0.003, noise=1.5, taps=(1.0, 0.2 - 0.1j, 0.05j), early=2):
    """payloads + int16 rx records [n, rx_len, 2] (preamble-aligned up to `early` samples early)"""
    s = o.sizes
    rng = np.random.default_rng(seed)
    pay = synth.payloads(n_frames, s.usefull_size, seed=seed + 1)
    tx16 = np.stack([o.tx(p)[1] for p in pay]).reshape(n_frames, -1, 2)
    rx = synth.channel(tx16, seed=seed + 2, cfo=rng.uniform(-cfo_max, cfo_max, n_frames), phase=rng.uniform(0, 1, n_frames),
                       taps=taps, noise_sigma=noise)
    rxl = s.preamble_size + s.message_size
    early = min(early, s.t2sin_size)
    off = rng.integers(0, early + 1, n_frames)
    rec = np.stack([rx[i, s.t2sin_size - off[i]: s.t2sin_size - off[i] + rxl] for i in range(n_frames)])
    return pay, synth.to_i16(rec)


def check_rx_against_oracle(m, o, rec_i16, fmt="i16", want=None):
    """full fused chain with every tap; rec_i16 [n, rx_len, 2].  `want` = precomputed oracle outputs."""
    s = o.sizes
    rec_c = cplx(rec_i16)
    x = np.ascontiguousarray(rec_i16) if fmt == "i16" else rec_c.astype(np.complex64)
    out, taps, n_amb = m.rx_aligned_batch(x, taps=True)
    out = to_np(out)
    taps = {k: to_np(v) for k, v in taps.items()}
    stats = dict(synced=0.0, grid=0.0, chan=0.0, constell=0.0, differing=0, shift_mismatch=0, ambiguous=n_amb)
    for i in range(len(rec_c)):
        r = want[i] if want is not None else o.rx_aligned(rec_c[i])
        if taps["scal"][i, 0] != np.float32(r["scal"][0]):
            # The coarse-CFO arg-max landed on another bin.  Legitimate ONLY for a near-tie between two magnitudes of the
            # reference's own fp64 spectrum; anything else is a bug and fails here.  Such a frame is counted; its later
            # stages follow another (equally valid) estimate, so they are not comparable sample by sample -- the decoded
            # bytes still are (the fine CFO stage absorbs one step of the coarse grid), up to boundary-ambiguous symbols.
            assert coarse_cfo_near_tie(rec_c[i], s), f"frame {i}: coarse CFO {taps['scal'][i, 0]} != {r['scal'][0]} without a near-tie"
            stats["shift_mismatch"] += 1
            stats["differing"] += assert_bytes_match(out[i], r["bytes"], r["constell"], s.mod_type, f"rx frame {i} (coarse near-tie)")
            continue
        for k in ("synced", "grid", "chan", "constell"):
            if k not in taps:
                continue                                   # the generic path has no time-domain / grid taps
            g = taps[k][i]
            e = rel_l2(g, r[k])
            stats[k] = max(stats[k], e)
            assert e < TOL, (i, k, e)
        stats["differing"] += assert_bytes_match(out[i], r["bytes"], r["constell"], s.mod_type, f"rx frame {i}")
    return stats


def check_read_and_chan_char(m, o, seed=9):
    """FRAME_FORM::read (sync-less demodulation of whole frames) and PREAMBLE_FORM::chan_char"""
    s = o.sizes
    pay = synth.payloads(3, s.usefull_size, seed=seed)
    frames, i16 = [], []
    for p in pay:
        f, q = o.tx(p)
        frames.append(f)
        i16.append(q.reshape(-1, 2))
    # a static gain only: read() has no synchronisation at all; the pilot normalisation absorbs the gain
    frames = np.stack(frames) * 0.8
    out, restored, chan, amb = m.read_batch(frames.astype(np.complex64), taps=True)
    worst = 0.0
    for i in range(3):
        want_b, want_r = o.read(frames[i])
        worst = max(worst, rel_l2(to_np(restored)[i], want_r))
        assert_bytes_match(to_np(out)[i], want_b, want_r, s.mod_type, "read")
        assert np.array_equal(want_b, pay[i])
        cc = o.chan_char(frames[i][s.t2sin_size: s.t2sin_size + s.preamble_size])
        worst = max(worst, rel_l2(to_np(chan)[i], cc))
    assert worst < TOL, worst
    out16, _ = m.read_batch(np.stack(i16))
    assert np.array_equal(to_np(out16), pay)
    return dict(rel_l2=worst)
