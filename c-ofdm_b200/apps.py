"""File-level look-alikes of the reference apps on top of the C ABI.

The reference's `main`, `tx` and `rx` talk to two PlutoSDRs; everything between the radio and the files they
read or write is the hot path this package provides.  These helpers replace the radio by an int16 capture file
(interleaved I,Q -- the SDR wire format, also what python_code/channel.py reads and writes) and keep every
artefact format the reference's Python tooling expects:

  * dumps of `main.cpp:74-78,106-108` in the formats of `io/io.hpp:15-79` -- `data/source.bin` (int16 I,Q),
    `data/data.bin`, `data/phases.bin`, `data/constell.bin` (interleaved float64 re,im), `data/t2_sin_corr.bin`
    (float64) and `data.txt` -- so `python_code/ofdm.py` plots a GPU run unchanged;
  * the per-iteration trace of `rx.cpp:25-43,129-235` (`KEY:value` pairs, one line per loop turn) so
    `python_code/timetrace.py::parse_log_file` reads it unchanged.  Stage times are amortised over the batch the GPU
    processed at once (there is no per-frame stage boundary on the device).

Host-side file shuffling only; all DSP goes through `Modem`.
"""
import os
import time

import numpy as np

from . import synth


# ---- io/io.hpp formats ------------------------------------------------------------------------------
def write_complex(path, data):
    """write_complex_to_file (io/io.hpp:15-52): re, im of every element in the element type's own width;
    complex arrays go out as float64 pairs, int16 [n, 2] arrays as int16 pairs."""
    a = np.asarray(data)
    if a.dtype == np.int16:
        a.reshape(-1, 2).tofile(path)
    else:
        np.ascontiguousarray(a, dtype=np.complex128).view(np.float64).tofile(path)


def read_complex(path, dtype=np.float64):
    """read_complex_from_file (io/io.hpp:54-69)"""
    raw = np.fromfile(path, dtype=dtype)
    if raw.size % 2:
        raise ValueError("File corrupted: incomplete complex number")
    return raw.reshape(-1, 2) if dtype == np.int16 else raw[0::2] + 1j * raw[1::2]


def write_double(path, data):
    """write_double_to_file (io/io.hpp:72-79)"""
    np.ascontiguousarray(data, dtype=np.float64).tofile(path)


def config_value(path, key):
    """one key of config.txt with the semantics of config/parser.cpp:4-33: `key = long`, lines starting with `#`
    are comments, white space is ignored, a missing key reads 0 (ConfigMap::operator[])"""
    import re
    val = 0
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line[0] == "#" or "=" not in line:
                continue
            k, v = line.split("=", 1)
            if "".join(k.split()) == key:
                m = re.match(r"[-+]?\d+", "".join(v.split()))       # std::stol: leading integer
                if m is None:
                    raise ValueError("stol")
                val = int(m.group(0))
    return val


# ---- main.cpp ---------------------------------------------------------------------------------------
def main_dump(modem, capture_i16, out_dir, tx_int16=None, mac=True):
    """The receive half of main.cpp:46-108 on one SDR block, with the same dump files.
    Returns a dict with t2_sin_begin, pr_begin, shift, payload (MAC body if mac=True), and the MAC header fields."""
    s = modem.sizes
    cap = np.ascontiguousarray(capture_i16, dtype=np.int16).reshape(-1, 2)
    os.makedirs(os.path.join(out_dir, "data"), exist_ok=True)
    rel = modem.t2sin_metric(cap)                                            # main.cpp:48 t2sin.corr
    corr = np.where(rel > np.float32(config_value(modem.config_path, "T2_sin_level") / 1000), rel, 0.0)   # Frame.hpp:140-142
    t2_begin = modem.find_t2sin(cap, 0)                                      # main.cpp:49
    pr_begin = int(modem.preamble_search(cap, np.array([t2_begin], dtype=np.int64))[0]) + 1   # main.cpp:51
    payload, taps, _ = modem.rx_aligned_batch(cap, n_frames=1, offset=pr_begin, taps=True)     # main.cpp:60-80
    if tx_int16 is not None:
        write_complex(os.path.join(out_dir, "data", "source.bin"), np.asarray(tx_int16, dtype=np.int16))
    write_complex(os.path.join(out_dir, "data", "data.bin"), cap[:, 0].astype(np.float64) + 1j * cap[:, 1])
    write_double(os.path.join(out_dir, "data", "t2_sin_corr.bin"), corr)
    write_complex(os.path.join(out_dir, "data", "phases.bin"), taps["chan"][0])
    write_complex(os.path.join(out_dir, "data", "constell.bin"), taps["constell"][0].reshape(-1))
    shift = float(taps["scal"][0][5]) / (s.num_pilot_subc * s.ofdm_len * s.num_pr_symb)      # kc / (NP * 640), exact
    res = {"t2_sin_begin": int(t2_begin), "pr_begin": pr_begin, "shift": shift, "frame_bytes": payload[0]}
    body = payload[0]
    if mac:
        body, tx_id, rx_id, seq, ok = synth.mac_read(payload[0])
        res.update(tx_id=tx_id, rx_id=rx_id, seq=seq, checksum_ok=ok)
    res["payload"] = body
    np.asarray(body, dtype=np.uint8).tofile(os.path.join(out_dir, "data.txt"))   # main.cpp:106-108
    return res


# ---- tx.cpp -----------------------------------------------------------------------------------------
def tx_file(modem, in_path, capture_path, tx_id=1, rx_id=0, gap=700, noise_sigma=0.0, seed=0, batch=4096):
    """tx.cpp:29-40 with the radio replaced by a capture file: the input file is cut into MAC payloads
    (usefull_size - 8 bytes, the last one zero-padded like fread into a reused buffer leaves it), each is framed
    (`mac.write`), modulated (`FRAME_FORM::write` + `get_int16`) and appended to `capture_path` as int16 I,Q with
    `gap` idle samples in front of every frame.  Returns the number of frames."""
    s = modem.sizes
    body = s.usefull_size - 8
    data = np.fromfile(in_path, dtype=np.uint8)
    n = (data.size + body - 1) // body
    rng = np.random.default_rng(seed)
    prev = np.zeros(body, dtype=np.uint8)
    with open(capture_path, "wb") as f:
        for f0 in range(0, n, batch):
            k = min(batch, n - f0)
            frames = np.zeros((k, s.usefull_size), dtype=np.uint8)
            for i in range(k):
                chunk = data[(f0 + i) * body:(f0 + i + 1) * body]
                prev = prev.copy()
                prev[:chunk.size] = chunk                    # a short last read leaves the tail of the previous payload
                frames[i] = synth.mac_write(prev, tx_id, rx_id, (f0 + i) & 0xFFFF)
            tx16 = modem.tx_batch(frames, fmt=1)             # COFDM_CI16: [k, output_size, 2]
            rec = np.zeros((k, gap + s.output_size, 2), dtype=np.int16)
            rec[:, gap:] = tx16
            if noise_sigma > 0:
                rec = (rec + np.rint(rng.normal(0, noise_sigma, rec.shape))).astype(np.int16)
            rec.tofile(f)
        # idle tail so that the stream receiver sees whole SDR blocks after the last frame
        tail = np.zeros((s.output_size * (s.rx_buf_size + 1), 2), dtype=np.int16)
        if noise_sigma > 0:
            tail = np.rint(rng.normal(0, noise_sigma, tail.shape)).astype(np.int16)
        tail.tofile(f)
    return n


# ---- rx.cpp -----------------------------------------------------------------------------------------
def rx_file(modem, capture_path, out_path, log_path=None, shards=1, n_bytes=None):
    """rx.cpp:101-235 with the radio replaced by a capture file: stream-receive every frame, `mac.read` it,
    append the payloads to `out_path`; optionally write a LOG.txt-format trace.  Returns a dict of counters."""
    s = modem.sizes
    cap = np.fromfile(capture_path, dtype=np.int16).reshape(-1, 2)
    modem.enable_timing(True)
    t0 = time.perf_counter()
    pos, frames = modem.rx_stream(cap, shards=shards)
    t_rx = time.perf_counter() - t0
    stages = modem.last_stage_ms()
    modem.enable_timing(False)
    t1 = time.perf_counter()
    bodies, seqs, bad_cs = [], [], 0
    for fr in frames:
        body, _, _, seq, ok = synth.mac_read(fr)
        bodies.append(body)
        seqs.append(seq)
        bad_cs += 0 if ok else 1
    t_mac = time.perf_counter() - t1
    out = np.concatenate(bodies) if bodies else np.zeros(0, np.uint8)
    if n_bytes is not None:
        out = out[:n_bytes]
    out.tofile(out_path)
    if log_path is not None:
        write_trace(log_path, np.asarray(pos, dtype=np.int64), stages, t_rx, t_mac, seqs, block_samples=s.output_size * max(1, s.rx_buf_size))
    return {"frames": len(frames), "bad_checksums": bad_cs, "positions": pos, "seq": np.array(seqs, dtype=np.int64),
            "seconds_rx": t_rx, "stage_ms": stages}


def write_trace(path, positions, stage_ms, t_rx, t_mac, seqs, block_samples):
    """LOG.txt in the format rx.cpp:32-36,128-235 prints (one line per frame, `KEY:seconds ` pairs; python_code/timetrace.py
    parses any keys).  Every value is MEASURED on this run -- nothing is taken from the reference's own LOG.txt:
      T2SIN            device time of the stream scanner (sync-tone detector + preamble search, stream.cuh) / frames
      PILOT_SINH       device time of the acquire kernel (pilot_freq_sinh, the preamble's cp_freq_sinh + pr_phase_sinh, chan_char_lq) / frames
      FREQ_PHASE_SINH  0: the per-symbol CFO / phase corrections are fused into the two neighbouring kernels and cannot be timed apart
      PFC              device time of the demod kernel (cp_freq_sinh of the message symbols, FFT, pilots, equaliser, demap) / frames
      MAC              host time of MAC::read / frames
      GATHER, MERGE, D2H   frame gather kernel (0 when every frame is demodulated in place), host-side list read-back + shard merge,
                           payload copy-back, / frames
      CONVERT          upload of the capture (the device-side form_int16_to_double), on the first frame of each SDR block, per block
      FR_IN_BUF        ordinal of the frame inside its SDR block, from the detected positions
      TIME             wall time of the whole call / frames (includes what the stages above do not cover: launches, synchronisation)
    The GPU processes a capture as one batch, so per-frame values are batch times divided by the number of frames."""
    n = len(positions)
    per = lambda key: stage_ms.get(key, 0.0) * 1e-3 / max(1, n)
    blocks = positions // max(1, block_samples)
    n_blocks = int(blocks.max()) + 1 if n else 1
    convert = stage_ms.get("upload", 0.0) * 1e-3 / n_blocks
    tot = (t_rx + t_mac) / max(1, n)
    with open(path, "w") as f:
        g, in_buf, prev_blk = 0.0, 0, -1
        for i in range(n):
            in_buf = in_buf + 1 if blocks[i] == prev_blk else 1
            prev_blk = blocks[i]
            extra = f"CONVERT:{convert:.6g} " if in_buf == 1 else ""
            f.write(f"ITER:{i} GLOBAL:{g:.6g} T2SIN:{per('scan'):.6g} {extra}PILOT_SINH:{per('acquire'):.6g} FREQ_PHASE_SINH:{0.0:.6g} "
                    f"PFC:{per('demod'):.6g} MAC:{t_mac / max(1, n):.6g} GATHER:{per('gather'):.6g} MERGE:{per('merge'):.6g} D2H:{per('d2h'):.6g} "
                    f"SEQ:{int(seqs[i])} DET:{i} FR_IN_BUF:{in_buf} TIME:{tot:.6g}\n")
            g += tot
