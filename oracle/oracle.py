"""ctypes loader for the two CPU checkers declared in oracle/oracle_api.h.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

    Oracle("port")       -> oracle/libcofdm_oracle.so   (C restatement, builds anywhere)
    Oracle("reference")  -> oracle/_ref/libcofdm_ref.so (unmodified reference sources + stand-in FFT)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBS = {"port": os.path.join(HERE, "libcofdm_oracle.so"),
        "reference": os.path.join(HERE, "_ref", "libcofdm_ref.so")}


class Sizes(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "fft_size", "num_data_subc", "num_pilot_subc", "cp_size", "num_symb", "num_pr_symb",
        "pr_sin_len", "pr_seed", "t2sin_size", "t2_f1", "t2_f2", "smooth", "mod_type",
        "ofdm_len", "preamble_size", "message_size", "output_size", "usefull_size",
        "constell_size", "mult", "rx_buf_size", "iterations", "cor_size")] + [
        (n, C.c_double) for n in ("t2_level", "pr_level", "pilot_ampl")]


def build(kind="port"):
    """(Re)build a checker with oracle/Makefile; `reference` needs /root/reference."""
    target = "libcofdm_oracle.so" if kind == "port" else "ref"
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


def available(kind):
    return os.path.exists(LIBS[kind])


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _c128(n):
    return np.zeros(n, dtype=np.complex128)


class Oracle:
    def __init__(self, kind="port", config_path=None):
        if not os.path.exists(LIBS[kind]):
            raise FileNotFoundError(f"{LIBS[kind]} not built (make -C oracle)")
        self.kind = kind
        self.lib = lib = C.CDLL(LIBS[kind])
        lib.oc_kind.restype = C.c_char_p
        lib.oc_last_error.restype = C.c_char_p
        lib.oc_create.restype = C.c_void_p
        lib.oc_create.argtypes = [C.c_char_p]
        lib.oc_destroy.argtypes = [C.c_void_p]
        lib.oc_get_sizes.argtypes = [C.c_void_p, C.POINTER(Sizes)]
        lib.oc_get_constants.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        lib.oc_bit_stream_converter.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        lib.oc_mod.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        lib.oc_demod.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        lib.oc_tx.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        lib.oc_t2sin_corr.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        lib.oc_find_t2sin.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int]
        lib.oc_find_corr.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p]
        lib.oc_find_preamble.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int]
        lib.oc_rx_aligned.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        lib.oc_read.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        lib.oc_chan_char.argtypes = [C.c_void_p] + [C.c_void_p] * 2
        lib.oc_rx_stream.restype = C.c_int
        lib.oc_rx_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p]
        lib.oc_txrx_loop.restype = C.c_long
        lib.oc_txrx_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        assert lib.oc_kind().decode() == kind
        self.h = None
        if config_path is not None:
            self.h = lib.oc_create(os.fsencode(config_path))
            if not self.h:
                raise RuntimeError(lib.oc_last_error().decode())
            s = Sizes()
            lib.oc_get_sizes(self.h, C.byref(s))
            self.sizes = s

    def close(self):
        if self.h:
            self.lib.oc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- constants ------------------------------------------------------------------------------
    def constants(self):
        s = self.sizes
        out = dict(t2sin_tone=_c128(s.t2sin_size), t2_mask=np.zeros(s.t2sin_size),
                   preamble_bytes=np.zeros(s.num_data_subc * s.num_pr_symb // 8, dtype=np.uint8),
                   ofdm_preamble=_c128(s.preamble_size), mod_preamble=_c128(s.num_data_subc * s.num_pr_symb),
                   matched=_c128(s.pr_sin_len), constell=_c128(1 << s.mod_type))
        self.lib.oc_get_constants(self.h, *[v.ctypes.data for v in out.values()])
        return out

    # ---- modulation -----------------------------------------------------------------------------
    def bit_stream_converter(self, out_bits, in_bits, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        n_out = (len(data) * in_bits + out_bits - 1) // out_bits
        out = np.zeros(max(n_out, 1), dtype=np.uint8)
        n = self.lib.oc_bit_stream_converter(out_bits, in_bits, data.ctypes.data, len(data), out.ctypes.data)
        return out[:n]

    def mod(self, mod_type, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        n_out = (len(data) * 8 + mod_type - 1) // mod_type
        pts = _c128(max(n_out, 1))
        n = self.lib.oc_mod(mod_type, data.ctypes.data, len(data), pts.ctypes.data)
        return pts[:n]

    def demod(self, mod_type, points):
        """-> (bytes, clamped_points) ; the reference clamps its argument in place."""
        pts = np.array(points, dtype=np.complex128)
        out = np.zeros((len(pts) * mod_type + 7) // 8 + 1, dtype=np.uint8)
        n = self.lib.oc_demod(mod_type, pts.ctypes.data, len(pts), out.ctypes.data)
        return out[:n], pts

    # ---- tx -------------------------------------------------------------------------------------
    def tx(self, payload, int16=True):
        s = self.sizes
        payload = np.ascontiguousarray(payload, dtype=np.uint8)
        assert len(payload) == s.usefull_size
        frame = _c128(s.output_size)
        q = np.zeros(2 * s.output_size, dtype=np.int16) if int16 else None
        self.lib.oc_tx(self.h, payload.ctypes.data, frame.ctypes.data, q.ctypes.data if int16 else None)
        return (frame, q) if int16 else frame

    # ---- sync -----------------------------------------------------------------------------------
    def t2sin_corr(self, sig):
        sig = np.ascontiguousarray(sig, dtype=np.complex128)
        out = np.zeros(len(sig) // self.sizes.t2sin_size)
        self.lib.oc_t2sin_corr(self.h, sig.ctypes.data, len(sig), out.ctypes.data)
        return out

    def find_t2sin(self, sig, start):
        sig = np.ascontiguousarray(sig, dtype=np.complex128)
        return self.lib.oc_find_t2sin(self.h, sig.ctypes.data, len(sig), start)

    def find_corr(self, sig, start):
        sig = np.ascontiguousarray(sig, dtype=np.complex128)
        out = np.zeros(self.sizes.cor_size)
        self.lib.oc_find_corr(self.h, sig.ctypes.data, len(sig), start, out.ctypes.data)
        return out

    def find_preamble(self, sig, start):
        sig = np.ascontiguousarray(sig, dtype=np.complex128)
        return self.lib.oc_find_preamble(self.h, sig.ctypes.data, len(sig), start)

    # ---- rx -------------------------------------------------------------------------------------
    def rx_aligned(self, samples):
        s = self.sizes
        samples = np.ascontiguousarray(samples, dtype=np.complex128)
        assert len(samples) == s.preamble_size + s.message_size
        r = dict(scal=np.zeros(4), synced=_c128(len(samples)), grid=_c128(s.num_symb * s.fft_size),
                 chan=_c128(s.num_data_subc), constell=_c128(s.constell_size),
                 bytes=np.zeros(s.usefull_size, dtype=np.uint8))
        self.lib.oc_rx_aligned(self.h, samples.ctypes.data, *[v.ctypes.data for v in r.values()])
        return r

    def read(self, frame):
        s = self.sizes
        frame = np.ascontiguousarray(frame, dtype=np.complex128)
        assert len(frame) == s.output_size
        restored = _c128(s.constell_size)
        out = np.zeros(s.usefull_size, dtype=np.uint8)
        self.lib.oc_read(self.h, frame.ctypes.data, restored.ctypes.data, out.ctypes.data)
        return out, restored

    def chan_char(self, samples):
        samples = np.ascontiguousarray(samples, dtype=np.complex128)
        out = _c128(self.sizes.num_data_subc)
        self.lib.oc_chan_char(self.h, samples.ctypes.data, out.ctypes.data)
        return out

    def rx_stream(self, capture_i16, max_frames=1 << 20):
        s = self.sizes
        cap = np.ascontiguousarray(capture_i16, dtype=np.int16).reshape(-1)
        n = len(cap) // 2
        max_frames = min(max_frames, n // s.message_size + 2)
        pos = np.zeros(max_frames, dtype=np.int64)
        out = np.zeros(max_frames * s.usefull_size, dtype=np.uint8)
        k = self.lib.oc_rx_stream(self.h, cap.ctypes.data, n, max_frames, pos.ctypes.data, out.ctypes.data)
        return pos[:k].copy(), out[:k * s.usefull_size].reshape(k, s.usefull_size).copy()

    def txrx_loop(self, payloads, want_bytes=False):
        """bench.py CPU baseline: tx -> int16 -> double -> aligned rx for every payload row, in C."""
        payloads = np.ascontiguousarray(payloads, dtype=np.uint8).reshape(-1, self.sizes.usefull_size)
        out = np.zeros_like(payloads) if want_bytes else None
        bad = self.lib.oc_txrx_loop(self.h, payloads.ctypes.data, len(payloads), out.ctypes.data if want_bytes else None)
        return (bad, out) if want_bytes else bad
