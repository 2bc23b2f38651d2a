"""Modem-shaped wrapper around tests/emu/libcofdm_emu.so (the real kernel source run under the CPU
thread emulator).  Same method names and return shapes as cofdm_b200.Modem so tests/parity_checks.py
runs unchanged against either.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))
import build_emu  # noqa: E402

CF32, CI16 = 0, 1


class EmuSizes:
    pass


class EmuModem:
    is_emulator = True

    def __init__(self, config_path, oracle_sizes):
        # COFDM_EMU_LIB: use a pre-built variant of the emulator library (the ASan build of tests/test_emu_asan.py)
        self.lib = L = C.CDLL(os.environ.get("COFDM_EMU_LIB") or build_emu.build())
        vp = C.c_void_p
        L.emu_create.restype = vp
        L.emu_create.argtypes = [C.c_char_p]
        L.emu_destroy.argtypes = [vp]
        L.emu_rx_fused512.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_longlong] + [vp] * 7
        L.emu_tx512.argtypes = [vp, vp, C.c_int, vp, C.c_int]
        L.emu_rx_fused512_notaps.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_longlong, vp, vp]
        L.emu_rx_fused512_mode.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int] + [vp] * 7
        L.emu_t2sin_metric.argtypes = [vp, vp, C.c_int, C.c_longlong, C.c_longlong, vp]
        L.emu_preamble_corr.argtypes = [vp, vp, C.c_int, C.c_longlong, vp, C.c_int, vp, vp]
        L.emu_rx_generic.argtypes = [vp, vp, C.c_int, C.c_int, C.c_longlong] + [vp] * 5
        L.emu_tx_generic.argtypes = [vp, vp, C.c_int, vp, C.c_int]
        L.emu_rx_generic_mode.argtypes = [vp, vp, C.c_int, C.c_int, C.c_longlong, C.c_int] + [vp] * 5
        L.emu_rx_big.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int] + [vp] * 5
        L.emu_big_ok.argtypes = [vp]
        L.emu_fused_ok.argtypes = [vp]
        L.emu_mod.argtypes = [vp, C.c_int, vp, C.c_longlong, vp, C.c_longlong]
        L.emu_demod.argtypes = [vp, C.c_int, vp, C.c_longlong, vp, C.c_longlong, vp]
        self.h = L.emu_create(os.fsencode(config_path))
        assert self.h
        o = oracle_sizes
        s = self.sizes = EmuSizes()
        for n in ("fft_size", "num_data_subc", "num_pilot_subc", "cp_size", "num_symb", "num_pr_symb", "pr_sin_len",
                  "t2sin_size", "mod_type", "ofdm_len", "output_size", "usefull_size", "constell_size", "cor_size",
                  "rx_buf_size", "iterations"):
            setattr(s, n, getattr(o, n))
        s.rx_len = o.preamble_size + o.message_size
        self.use_tma = 1
        L.emu_stream_scan.argtypes = [vp, vp, vp, vp, C.c_int, vp, C.c_int, vp]
        self.fused = bool(L.emu_fused_ok(self.h))
        self.big = bool(L.emu_big_ok(self.h))         # fft-4096 cluster kernels (big.cuh)
        self.big_mode = 0
        s.fused_path = 1 if self.fused else 0

    def set_pc_plain(self, on):
        """1: the one-lag-per-slot preamble search kernel (fallback for odd sizes), 0: the 4-lags-per-thread one"""
        self.lib.emu_set_pc_plain.argtypes = [C.c_int]
        self.lib.emu_set_pc_plain(int(on))

    def set_big_lay(self, on):
        """1: fft-4096 configurations whose map allows it run the layout-specialised demod instance (the product default)"""
        self.lib.emu_set_big_lay.argtypes = [C.c_int]
        self.lib.emu_set_big_lay(int(on))

    def big_lay(self):
        self.lib.emu_big_lay.argtypes = [C.c_void_p]
        return int(self.lib.emu_big_lay(self.h))

    def set_tx_bulk(self, on):
        """1: tx symbols leave as TMA bulk stores of linear images (the product default), 0: register stores"""
        self.lib.emu_set_tx_bulk.argtypes = [C.c_int]
        self.lib.emu_set_tx_bulk(int(on))

    def close(self):
        if self.h:
            self.lib.emu_destroy(self.h)
            self.h = None

    def mod(self, data, mod_type=None):
        mod_type = mod_type or self.sizes.mod_type
        data = np.ascontiguousarray(data, np.uint8)
        n_pts = (data.size * 8 + mod_type - 1) // mod_type
        out = np.zeros(n_pts, np.complex64)
        assert self.lib.emu_mod(self.h, mod_type, data.ctypes.data, data.size, out.ctypes.data, n_pts) == 0
        return out

    def demod(self, points, mod_type=None):
        mod_type = mod_type or self.sizes.mod_type
        points = np.ascontiguousarray(points, np.complex64)
        nb = (points.size * mod_type + 7) // 8
        out = np.zeros(nb, np.uint8)
        amb = np.zeros(1, np.uint64)
        assert self.lib.emu_demod(self.h, mod_type, points.ctypes.data, points.size, out.ctypes.data, nb, amb.ctypes.data) == 0
        return out, int(amb[0])

    def tx_batch(self, payload, fmt=CF32, out=None):
        s = self.sizes
        payload = np.ascontiguousarray(payload, np.uint8)
        n = payload.size // s.usefull_size
        out = np.zeros((n, s.output_size), np.complex64) if fmt == CF32 else np.zeros((n, s.output_size, 2), np.int16)
        fn = self.lib.emu_tx512 if self.fused else self.lib.emu_tx_generic
        assert fn(self.h, payload.ctypes.data, n, out.ctypes.data, fmt) == 0
        return out

    def rx_aligned_batch(self, samples, n_frames=None, frame_stride=None, offset=0, out=None, taps=False, count_ambiguous=True):
        s = self.sizes
        fmt = CI16 if samples.dtype == np.int16 else CF32
        samples = np.ascontiguousarray(samples)
        total = samples.size // 2 if fmt == CI16 else samples.size
        frame_stride = frame_stride or s.rx_len
        if n_frames is None:
            n_frames = (total - offset - s.rx_len) // frame_stride + 1
        out = np.zeros((n_frames, s.usefull_size), np.uint8)
        amb = np.zeros(1, np.uint64)
        t = dict(scal=np.zeros((n_frames, 48), np.float32), grid=np.zeros((n_frames, s.num_symb * s.fft_size), np.complex64),
                 chan=np.zeros((n_frames, s.num_data_subc), np.complex64), constell=np.zeros((n_frames, s.constell_size), np.complex64),
                 synced=np.zeros((n_frames, s.rx_len), np.complex64))
        base = samples.ctypes.data + offset * (4 if fmt == CI16 else 8)
        ptrs = [v.ctypes.data for v in t.values()] if taps else [None] * 5
        if self.fused and not taps:
            assert self.lib.emu_rx_fused512_notaps(self.h, base, fmt, self.use_tma, n_frames, frame_stride, out.ctypes.data,
                                                   amb.ctypes.data if count_ambiguous else None) == 0
            return out, int(amb[0])
        if not self.fused and self.big:
            assert self.lib.emu_rx_big(self.h, base, fmt, self.use_tma, n_frames, frame_stride, self.big_mode, out.ctypes.data,
                                       amb.ctypes.data if count_ambiguous else None, ptrs[0], ptrs[2], ptrs[3]) == 0
            t.pop("grid"); t.pop("synced")
            return (out, t, int(amb[0])) if taps else (out, int(amb[0]))
        if not self.fused:
            assert self.lib.emu_rx_generic(self.h, base, fmt, n_frames, frame_stride, out.ctypes.data, amb.ctypes.data,
                                           ptrs[0], ptrs[2], ptrs[3]) == 0
            t.pop("grid"); t.pop("synced")
            return (out, t, int(amb[0])) if taps else (out, int(amb[0]))
        assert self.lib.emu_rx_fused512(self.h, base, fmt, self.use_tma, n_frames, frame_stride, out.ctypes.data,
                                        amb.ctypes.data, *ptrs) == 0
        return (out, t, int(amb[0])) if taps else (out, int(amb[0]))

    def read_batch(self, frames, taps=False):
        s = self.sizes
        fmt = CI16 if frames.dtype == np.int16 else CF32
        frames = np.ascontiguousarray(frames)
        n = (frames.size // 2 if fmt == CI16 else frames.size) // s.output_size
        out = np.zeros((n, s.usefull_size), np.uint8)
        amb = np.zeros(1, np.uint64)
        restored = np.zeros((n, s.constell_size), np.complex64)
        chan = np.zeros((n, s.num_data_subc), np.complex64)
        base = frames.ctypes.data + s.t2sin_size * (4 if fmt == CI16 else 8)
        if not self.fused:
            # the any-size kernels in their sync-less form (no chan_char output there)
            assert self.lib.emu_rx_generic_mode(self.h, base, fmt, n, s.output_size, 1, out.ctypes.data, amb.ctypes.data,
                                                None, None, restored.ctypes.data) == 0
            return (out, restored, None, int(amb[0])) if taps else (out, int(amb[0]))
        assert self.lib.emu_rx_fused512_mode(self.h, base, fmt, self.use_tma, n, s.output_size, 1, out.ctypes.data, amb.ctypes.data,
                                             None, None, chan.ctypes.data, restored.ctypes.data, None) == 0
        return (out, restored, chan, int(amb[0])) if taps else (out, int(amb[0]))

    def t2sin_metric(self, samples, start=0):
        fmt = CI16 if samples.dtype == np.int16 else CF32
        samples = np.ascontiguousarray(samples)
        n = samples.size // 2 if fmt == CI16 else samples.size
        nb = max(0, (n - start) // self.sizes.t2sin_size)
        out = np.zeros(nb, np.float32)
        assert self.lib.emu_t2sin_metric(self.h, samples.ctypes.data, fmt, start, nb, out.ctypes.data) == 0
        return out

    def find_t2sin(self, samples, start=0):
        rel = self.t2sin_metric(samples, start)
        hit = np.nonzero(rel > np.float32(0.8))[0]
        return int(start + hit[0] * self.sizes.t2sin_size) if len(hit) else -1

    def preamble_search(self, samples, starts, want_cor=False):
        fmt = CI16 if samples.dtype == np.int16 else CF32
        samples = np.ascontiguousarray(samples)
        n = samples.size // 2 if fmt == CI16 else samples.size
        starts = np.ascontiguousarray(starts, np.int64)
        first = np.zeros(len(starts), np.int64)
        cor = np.zeros((len(starts), self.sizes.cor_size), np.float32)
        assert self.lib.emu_preamble_corr(self.h, samples.ctypes.data, fmt, n, starts.ctypes.data, len(starts),
                                          cor.ctypes.data if want_cor else None, first.ctypes.data) == 0
        return (first, cor) if want_cor else first

    def stream_scan(self, capture_i16, shards):
        """device-side acquisition loop; shards = [(first_sample, n_blocks), ...] -> list of position arrays"""
        cap = np.ascontiguousarray(capture_i16, dtype=np.int16)
        first = np.array([a for a, _ in shards], dtype=np.int64)
        nblk = np.array([b for _, b in shards], dtype=np.int64)
        s = self.sizes
        max_per = int(max(nblk) * s.output_size * s.rx_buf_size // (s.ofdm_len * s.num_symb) + 2)
        pos = np.zeros((len(shards), max_per), dtype=np.int64)
        cnt = np.zeros(len(shards), dtype=np.int32)
        assert self.lib.emu_stream_scan(self.h, cap.ctypes.data, first.ctypes.data, nblk.ctypes.data, len(shards),
                                        pos.ctypes.data, max_per, cnt.ctypes.data) == 0
        return [pos[i, :cnt[i]].copy() for i in range(len(shards))]
