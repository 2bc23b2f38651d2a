"""Seeded synthetic workloads (host side, numpy): payloads, a baseband channel and long captures.

The reference has no channel model (python_code/channel.py drives a PlutoSDR) and its WAV payload is
not shipped; these generators stand in for both.  Everything is deterministic in
the seed so the oracle and the CUDA path see bit-identical inputs.
"""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_CONFIG = os.path.join(ROOT, "config", "config.txt")


def write_config(path, base=DEFAULT_CONFIG, **overrides):
    """Copy `base` to `path`, replacing `key = value` for every override (keys as in config.txt)."""
    text = open(base).read()
    for k, v in overrides.items():
        pat = re.compile(rf"^{re.escape(k)}\s*=.*$", re.M)
        line = f"{k} = {int(v)}"
        text = pat.sub(line, text) if pat.search(text) else text + "\n" + line + "\n"
    with open(path, "w") as f:
        f.write(text)
    return path


def payloads(n_frames, bytes_per_frame, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (n_frames, bytes_per_frame), dtype=np.uint8)


def wav_payload(seconds=2.0, rate=44100, seed=7):
    """Deterministic mono 16-bit PCM WAV file image (header + samples) as bytes: a stand-in for the
    reference's missing FlyMeToTheMoon_mono.wav (tx.cpp:30 streams the raw file bytes)."""
    rng = np.random.default_rng(seed)
    t = np.arange(int(seconds * rate)) / rate
    x = sum(a * np.sin(2 * np.pi * f * t + p) for a, f, p in zip((0.4, 0.25, 0.15), (220.0, 277.18, 329.63), rng.uniform(0, 6.28, 3)))
    pcm = np.round(x / np.abs(x).max() * 30000).astype("<i2").tobytes()
    hdr = b"RIFF" + (36 + len(pcm)).to_bytes(4, "little") + b"WAVEfmt " + (16).to_bytes(4, "little") + \
        (1).to_bytes(2, "little") + (1).to_bytes(2, "little") + rate.to_bytes(4, "little") + \
        (rate * 2).to_bytes(4, "little") + (2).to_bytes(2, "little") + (16).to_bytes(2, "little") + \
        b"data" + len(pcm).to_bytes(4, "little")
    return hdr + pcm


def channel(frames_i16, seed=0, cfo=0.0, phase=0.0, taps=(1.0,), noise_sigma=0.0, gain=1.0):
    """Apply CFO (cycles/sample), a constant phase, an FIR multipath, gain and AWGN to int16 frames
    [n, L, 2] (or [L, 2]); returns complex128 rounded to the int16 grid (what an ADC would deliver)."""
    x = np.asarray(frames_i16)
    x = (x[..., 0] + 1j * x[..., 1]).astype(np.complex128)
    single = x.ndim == 1
    x = np.atleast_2d(x)
    rng = np.random.default_rng(seed)
    n = np.arange(x.shape[1])
    cfo = np.broadcast_to(np.asarray(cfo, dtype=np.float64).reshape(-1, 1), (x.shape[0], 1))
    phase = np.broadcast_to(np.asarray(phase, dtype=np.float64).reshape(-1, 1), (x.shape[0], 1))
    y = np.zeros_like(x)
    for d, h in enumerate(taps):
        y[:, d:] += h * x[:, :x.shape[1] - d]
    y = gain * y * np.exp(2j * np.pi * (cfo * n + phase))
    if noise_sigma > 0:
        y = y + rng.normal(0, noise_sigma, y.shape) + 1j * rng.normal(0, noise_sigma, y.shape)
    y = np.clip(np.round(y.real), -32768, 32767) + 1j * np.clip(np.round(y.imag), -32768, 32767)
    return y[0] if single else y


def to_i16(x):
    x = np.asarray(x)
    return np.stack([x.real, x.imag], axis=-1).astype(np.int16)


def capture(frames_c, gaps, noise_sigma=3.0, seed=0, tail=0):
    """Place complex frames [n, L] into a long capture with `gaps[i]` noise-only samples before frame i
    (noise floor like the reference's data/data.bin).  Returns (int16 capture [N, 2], frame starts)."""
    rng = np.random.default_rng(seed)
    frames_c = np.asarray(frames_c)
    L = frames_c.shape[1]
    total = int(np.sum(gaps)) + L * len(frames_c) + int(tail)
    cap = rng.normal(0, noise_sigma, total) + 1j * rng.normal(0, noise_sigma, total)
    starts, pos = [], 0
    for f, g in zip(frames_c, gaps):
        pos += int(g)
        cap[pos:pos + L] += f
        starts.append(pos)
        pos += L
    return to_i16(np.round(cap.real) + 1j * np.round(cap.imag)), np.array(starts, dtype=np.int64)


# ---- MAC framing stand-in (the reference's mac/mac_frame.hpp is missing from its tree; layout recovered in
# 8-byte little-endian header {u16 tx_id, rx_id, seq_num, cs}, cs = 16-bit sum of
# every byte of the frame taken with cs = 0).  Host-side byte shuffling for the loopback harnesses only.
def mac_write(payload, tx_id=1, rx_id=0, seq=0):
    payload = np.asarray(payload, dtype=np.uint8)
    hdr = np.array([tx_id & 255, tx_id >> 8, rx_id & 255, rx_id >> 8, seq & 255, (seq >> 8) & 255, 0, 0], dtype=np.uint8)
    cs = (int(hdr.sum()) + int(payload.sum())) & 0xFFFF
    hdr[6], hdr[7] = cs & 255, cs >> 8
    return np.concatenate([hdr, payload])


def mac_read(frame):
    """-> (payload, tx_id, rx_id, seq, checksum_ok)"""
    frame = np.asarray(frame, dtype=np.uint8)
    tx_id, rx_id, seq, cs = (int(frame[2 * i]) | int(frame[2 * i + 1]) << 8 for i in range(4))
    ok = ((int(frame[:6].sum()) + int(frame[8:].sum())) & 0xFFFF) == cs
    return frame[8:], tx_id, rx_id, seq, ok


def text_payload(n_bytes, seed=3):
    """deterministic ASCII text standing in for the reference's WARANDPEACE.txt (not shipped here)"""
    rng = np.random.default_rng(seed)
    words = ["peace", "war", "prince", "andrew", "natasha", "pierre", "moscow", "the", "and", "of", "a", "said", "was", "his", "her"]
    out = []
    size = 0
    while size < n_bytes:
        w = words[int(rng.integers(len(words)))] + (" " if rng.random() > 0.08 else ".\n")
        out.append(w)
        size += len(w)
    return np.frombuffer("".join(out).encode()[:n_bytes], dtype=np.uint8).copy()
