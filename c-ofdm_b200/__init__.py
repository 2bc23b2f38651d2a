"""cofdm_b200 -- ctypes binding of libcofdm_b200.so (include/cofdm.h).

The product is the CUDA library; this module is the thin Python shim the north star asks for (what a
new-style python_code/ofdm.py imports).  It never computes anything itself and has NO CPU fallback: if
the shared library is missing or no sm_100 GPU is usable, calls raise.

Buffers: numpy arrays are host memory (COFDM_HOST: the library does H2D/D2H itself); torch CUDA
tensors are device memory (COFDM_DEVICE: kernels are enqueued on the current torch stream).
Sample format follows the dtype: complex64 -> cf32, int16 (..., 2) -> ci16.

The directory is named `c-ofdm_b200` after the reference; import it as `cofdm_b200` (cofdm_b200.py at
the repo root registers it under that name).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcofdm_b200.so")

HOST, DEVICE, DEVICE_IN = 0, 1, 2
CF32, CI16 = 0, 1
NOT_FOUND_T2SIN, NOT_FOUND_PREAMBLE = -1, -10


class CofdmError(RuntimeError):
    pass


class Sizes(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "fft_size", "num_data_subc", "num_pilot_subc", "cp_size", "num_symb", "num_pr_symb",
        "pr_sin_len", "t2sin_size", "mod_type", "ofdm_len", "rx_len", "output_size", "usefull_size",
        "constell_size", "cor_size", "mult", "rx_buf_size", "iterations", "fused_path", "device")]


class RxTaps(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("scal", "grid", "chan", "constell", "synced")]


_lib = None


def load_library():
    """dlopen the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    # COFDM_LIB_PATH: a developer hook for A/B runs of an experimental build of the SAME library (build.py --out)
    path = os.environ.get("COFDM_LIB_PATH") or LIB_PATH
    if not os.path.exists(path):
        raise CofdmError(f"{path} is missing: build it with `python c-ofdm_b200/build.py` "
                         "(__graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(path)
    vp, sz, ci = C.c_void_p, C.c_size_t, C.c_int
    lib.cofdm_last_error.restype = C.c_char_p
    lib.cofdm_version.restype = C.c_char_p
    lib.cofdm_create.argtypes = [C.c_char_p, ci, C.POINTER(vp)]
    lib.cofdm_destroy.argtypes = [vp]
    lib.cofdm_query.argtypes = [vp, C.POINTER(Sizes)]
    lib.cofdm_set_stream.argtypes = [vp, vp]
    lib.cofdm_synchronize.argtypes = [vp]
    lib.cofdm_own_stream.argtypes = [vp]
    lib.cofdm_own_stream.restype = vp
    lib.cofdm_get_constants.argtypes = [vp] + [vp] * 6
    lib.cofdm_mod.argtypes = [vp, ci, vp, sz, vp, ci]
    lib.cofdm_demod.argtypes = [vp, ci, vp, sz, vp, vp, ci]
    lib.cofdm_tx_batch.argtypes = [vp, vp, sz, vp, ci, ci]
    lib.cofdm_rx_aligned_batch.argtypes = [vp, vp, ci, sz, sz, vp, vp, C.POINTER(RxTaps), ci]
    lib.cofdm_read_batch.argtypes = [vp, vp, ci, sz, vp, vp, vp, vp, ci]
    lib.cofdm_t2sin_metric.argtypes = [vp, vp, ci, sz, sz, vp, ci]
    lib.cofdm_find_t2sin.argtypes = [vp, vp, ci, sz, sz, C.POINTER(C.c_longlong), ci]
    lib.cofdm_preamble_search.argtypes = [vp, vp, ci, sz, vp, sz, vp, vp, ci]
    lib.cofdm_i16_to_cf32.argtypes = [vp, vp, vp, sz, ci]
    lib.cofdm_rx_stream.argtypes = [vp, vp, sz, sz, vp, vp, C.POINTER(sz)]
    lib.cofdm_rx_stream_sharded.argtypes = [vp, vp, sz, C.c_int, C.c_int, sz, vp, vp, C.POINTER(sz), C.POINTER(sz)]
    lib.cofdm_ring_load.argtypes = [vp, vp, sz, C.POINTER(vp)]
    lib.cofdm_allreduce_counters.argtypes = [vp, vp, vp, sz, vp, sz]
    lib.cofdm_enable_timing.argtypes = [vp, ci]
    lib.cofdm_last_stage_ms.argtypes = [vp, vp, ci]
    lib.cofdm_last_kernel_ms.argtypes = [vp]
    lib.cofdm_last_kernel_ms.restype = C.c_float
    lib.cofdm_launch_count.argtypes = [vp]
    lib.cofdm_launch_count.restype = C.c_ulonglong
    _lib = lib
    return lib


def _is_torch(x):
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda")


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        return x.data_ptr()
    return x.ctypes.data


def _space(*xs):
    kinds = {bool(_is_torch(x) and x.is_cuda) for x in xs if x is not None}
    if len(kinds) != 1:
        raise CofdmError("all buffers of one call must live in the same memory space")
    return DEVICE if kinds.pop() else HOST


def _fmt_of(x):
    name = str(x.dtype)
    if "complex64" in name or "float32" in name:
        return CF32
    if "int16" in name:
        return CI16
    raise CofdmError(f"samples must be complex64 or int16 (I,Q), not {name}")


def _n_samples(x, fmt):
    n = int(np.prod(tuple(x.shape)))
    name = str(x.dtype)
    if fmt == CI16 or "float32" in name:
        if n % 2:
            raise CofdmError("interleaved sample arrays need an even element count")
        return n // 2
    return n


def _host_buffer(shape):
    """uint8 host array for results; page-locked (through torch's caching host allocator) when torch is there,
    so that the device-to-host copy of large results runs at PCIe speed"""
    n = int(np.prod(shape))
    if n >= (1 << 20):
        try:
            import torch
            if torch.cuda.is_available():
                return torch.empty(shape, dtype=torch.uint8, pin_memory=True).numpy()
        except Exception:
            pass
    return np.zeros(shape, dtype=np.uint8)


class Modem:
    """One FRAME_FORM-equivalent handle: owns the config, the device tables and a stream."""

    def __init__(self, config_path, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        self.config_path = str(config_path)
        rc = self.lib.cofdm_create(os.fsencode(config_path), int(device), C.byref(h))
        if rc != 0:
            raise CofdmError(f"cofdm_create failed ({rc}): {self.lib.cofdm_last_error().decode()}")
        self.h = h
        self.sizes = Sizes()
        self._chk(self.lib.cofdm_query(self.h, C.byref(self.sizes)))
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.cofdm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise CofdmError(f"cofdm call failed ({rc}): {self.lib.cofdm_last_error().decode()}")

    # ---- streams / timing ---------------------------------------------------------------------------
    def use_torch_stream(self):
        import torch
        self._chk(self.lib.cofdm_set_stream(self.h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def use_own_stream(self):
        self._chk(self.lib.cofdm_set_stream(self.h, C.c_void_p(self.lib.cofdm_own_stream(self.h))))

    def set_stream(self, cuda_stream_ptr):
        self._chk(self.lib.cofdm_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._chk(self.lib.cofdm_synchronize(self.h))

    def enable_timing(self, on=True):
        self._chk(self.lib.cofdm_enable_timing(self.h, int(on)))

    def last_kernel_ms(self):
        return float(self.lib.cofdm_last_kernel_ms(self.h))

    def last_stage_ms(self):
        """measured per-stage milliseconds of the last stream call / of the rx launches since enable_timing (see cofdm.h)"""
        out = (C.c_float * 7)()
        self._chk(self.lib.cofdm_last_stage_ms(self.h, out, 7))
        return dict(zip(("upload", "scan", "merge", "gather", "acquire", "demod", "d2h"), [float(v) for v in out]))

    def launch_count(self):
        return int(self.lib.cofdm_launch_count(self.h))

    def _new(self, like, shape, dtype):
        if _is_torch(like):
            import torch
            if like.is_cuda:
                # device-space calls are stream-ordered with the caller's torch work
                self.lib.cofdm_set_stream(self.h, C.c_void_p(torch.cuda.current_stream(like.device).cuda_stream))
            return torch.empty(shape, dtype=getattr(torch, dtype), device=like.device)
        return np.empty(shape, dtype=getattr(np, dtype))

    def _follow(self, x):
        """enqueue on torch's current stream when `x` is a CUDA tensor (no-op for host arrays)"""
        if _is_torch(x) and x.is_cuda:
            import torch
            self.lib.cofdm_set_stream(self.h, C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))

    # ---- constants ------------------------------------------------------------------------------------
    def constants(self):
        s = self.sizes
        out = dict(t2sin_tone=np.zeros(s.t2sin_size, np.complex128),
                   preamble_bytes=np.zeros(s.num_data_subc * s.num_pr_symb // 8, np.uint8),
                   ofdm_preamble=np.zeros(s.ofdm_len * s.num_pr_symb, np.complex128),
                   mod_preamble=np.zeros(s.num_data_subc * s.num_pr_symb, np.complex128),
                   matched=np.zeros(s.pr_sin_len, np.complex128),
                   constell=np.zeros(1 << s.mod_type, np.complex128))
        self._chk(self.lib.cofdm_get_constants(self.h, *[v.ctypes.data for v in out.values()]))
        return out

    # ---- Modulation::mod / demod ------------------------------------------------------------------------
    def mod(self, data, mod_type=None):
        self._follow(data)
        mod_type = mod_type or self.sizes.mod_type
        n_bytes = int(np.prod(tuple(data.shape)))
        n_pts = (n_bytes * 8 + mod_type - 1) // mod_type
        out = self._new(data, (n_pts,), "complex64")
        self._chk(self.lib.cofdm_mod(self.h, mod_type, _ptr(data), n_bytes, _ptr(out), _space(data, out)))
        return out

    def demod(self, points, mod_type=None):
        self._follow(points)
        """-> (bytes, ambiguous_count)"""
        mod_type = mod_type or self.sizes.mod_type
        n_pts = _n_samples(points, CF32)
        out = self._new(points, ((n_pts * mod_type + 7) // 8,), "uint8")
        amb = C.c_ulonglong(0)
        self._chk(self.lib.cofdm_demod(self.h, mod_type, _ptr(points), n_pts, _ptr(out), C.addressof(amb), _space(points, out)))
        return out, int(amb.value)

    # ---- FRAME_FORM::write + get / get_int16 ------------------------------------------------------------
    def tx_batch(self, payload, fmt=CF32, out=None):
        self._follow(payload)
        s = self.sizes
        n_bytes = int(np.prod(tuple(payload.shape)))
        if n_bytes % s.usefull_size:
            raise CofdmError(f"payload must be a multiple of usefull_size={s.usefull_size} bytes")
        n = n_bytes // s.usefull_size
        if out is None:
            out = self._new(payload, (n, s.output_size), "complex64") if fmt == CF32 else self._new(payload, (n, s.output_size, 2), "int16")
        self._chk(self.lib.cofdm_tx_batch(self.h, _ptr(payload), n, _ptr(out), fmt, _space(payload, out)))
        return out

    # ---- fused aligned rx chain ---------------------------------------------------------------------------
    def rx_aligned_batch(self, samples, n_frames=None, frame_stride=None, offset=0, out=None, taps=False, count_ambiguous=True):
        """samples: [n_frames, rx_len] records (or any flat layout with frame_stride/offset in samples).
        -> bytes [n_frames, usefull_size]  (+ dict of taps, + ambiguous count)"""
        s = self.sizes
        self._follow(samples)
        fmt = _fmt_of(samples)
        total = _n_samples(samples, fmt)
        frame_stride = frame_stride or s.rx_len
        if n_frames is None:
            n_frames = (total - offset - s.rx_len) // frame_stride + 1 if total - offset >= s.rx_len else 0
        if n_frames and offset + (n_frames - 1) * frame_stride + s.rx_len > total:
            raise CofdmError("sample buffer too small for n_frames")
        if out is None:
            out = self._new(samples, (n_frames, s.usefull_size), "uint8")
        space = _space(samples, out)
        tap_bufs, tp = None, None
        if taps:
            tap_bufs = dict(scal=self._new(samples, (n_frames, 48), "float32"),
                            grid=self._new(samples, (n_frames, s.num_symb * s.fft_size), "complex64"),
                            chan=self._new(samples, (n_frames, s.num_data_subc), "complex64"),
                            constell=self._new(samples, (n_frames, s.constell_size), "complex64"),
                            synced=self._new(samples, (n_frames, s.rx_len), "complex64"))
            if s.fused_path != 1:                      # the generic path has no time-domain / grid taps
                tap_bufs["grid"] = tap_bufs["synced"] = None
            tp = RxTaps(*[_ptr(v) for v in tap_bufs.values()])
            tap_bufs = {k: v for k, v in tap_bufs.items() if v is not None}
        amb = C.c_ulonglong(0)
        base = _ptr(samples) + offset * (4 if fmt == CI16 else 8)
        self._chk(self.lib.cofdm_rx_aligned_batch(self.h, base, fmt, n_frames, frame_stride, _ptr(out),
                                                  C.addressof(amb) if count_ambiguous else None,
                                                  C.byref(tp) if tp is not None else None, space))
        if taps:
            return out, tap_bufs, int(amb.value)
        return out, int(amb.value)

    # ---- FRAME_FORM::read (sync-less) + PREAMBLE_FORM::chan_char ---------------------------------------------
    def read_batch(self, frames, taps=False):
        """frames: [n, output_size] whole frames -> bytes [n, usefull_size] (+ restored points, chan_char)"""
        s = self.sizes
        self._follow(frames)
        fmt = _fmt_of(frames)
        n = _n_samples(frames, fmt) // s.output_size
        out = self._new(frames, (n, s.usefull_size), "uint8")
        restored = self._new(frames, (n, s.constell_size), "complex64") if taps else None
        chan = self._new(frames, (n, s.num_data_subc), "complex64") if (taps and s.fused_path == 1) else None   # chan_char: fft-512 kernels only
        amb = C.c_ulonglong(0)
        self._chk(self.lib.cofdm_read_batch(self.h, _ptr(frames), fmt, n, _ptr(out), C.addressof(amb), _ptr(restored), _ptr(chan), _space(frames, out)))
        return (out, restored, chan, int(amb.value)) if taps else (out, int(amb.value))

    # ---- sync ------------------------------------------------------------------------------------------------
    def t2sin_metric(self, samples, start=0):
        self._follow(samples)
        fmt = _fmt_of(samples)
        n = _n_samples(samples, fmt)
        nb = max(0, (n - start) // self.sizes.t2sin_size)
        out = self._new(samples, (nb,), "float32")
        self._chk(self.lib.cofdm_t2sin_metric(self.h, _ptr(samples), fmt, n, start, _ptr(out), _space(samples, out)))
        return out

    def find_t2sin(self, samples, start=0):
        self._follow(samples)
        fmt = _fmt_of(samples)
        n = _n_samples(samples, fmt)
        pos = C.c_longlong(0)
        self._chk(self.lib.cofdm_find_t2sin(self.h, _ptr(samples), fmt, n, start, C.byref(pos), _space(samples)))
        return int(pos.value)

    def preamble_search(self, samples, starts, want_cor=False):
        self._follow(samples)
        fmt = _fmt_of(samples)
        n = _n_samples(samples, fmt)
        ns = int(np.prod(tuple(starts.shape)))
        first = self._new(samples, (ns,), "int64")
        cor = self._new(samples, (ns, self.sizes.cor_size), "float32") if want_cor else None
        self._chk(self.lib.cofdm_preamble_search(self.h, _ptr(samples), fmt, n, _ptr(starts), ns, _ptr(cor), _ptr(first),
                                                 _space(samples, starts, first)))
        return (first, cor) if want_cor else first

    def rx_stream(self, capture_i16, max_frames=None, shards=1, want_bytes=True, return_unmerged=False, bytes_on_device=False):
        """rx.cpp's acquisition loop over an int16 capture [N, 2] (numpy = host, torch cuda tensor = device)
        -> (preamble positions, payload bytes).  shards > 1: the capture is cut into that many ranges of whole
        SDR blocks, each scanned by its own CTA, the chains merged (cofdm_rx_stream_sharded).
        bytes_on_device (device captures only): the payloads come back as a torch cuda tensor and never cross PCIe."""
        s = self.sizes
        if isinstance(capture_i16, np.ndarray):
            cap = np.ascontiguousarray(capture_i16, dtype=np.int16).reshape(-1)
            n, space, ptr = cap.size // 2, HOST, cap.ctypes.data
        else:
            self._follow(capture_i16)
            cap = capture_i16.contiguous()
            n, space, ptr = cap.numel() // 2, DEVICE_IN, cap.data_ptr()
        if max_frames is None:
            max_frames = n // (s.ofdm_len * s.num_symb) + 2
        pos = np.zeros(max_frames, dtype=np.int64)
        k, um = C.c_size_t(0), C.c_size_t(0)
        if bytes_on_device and want_bytes and space == DEVICE_IN:
            import torch
            dout = torch.empty((max_frames, s.usefull_size), dtype=torch.uint8, device=capture_i16.device)
            self._chk(self.lib.cofdm_rx_stream_sharded(self.h, ptr, n, DEVICE, int(shards), max_frames, pos.ctypes.data,
                                                       dout.data_ptr(), C.byref(k), C.byref(um)))
            res = (pos[:k.value].copy(), dout[:k.value])
            return res + (um.value,) if return_unmerged else res
        out = _host_buffer((max_frames if want_bytes else 0, s.usefull_size))
        self._chk(self.lib.cofdm_rx_stream_sharded(self.h, ptr, n, space, int(shards), max_frames, pos.ctypes.data,
                                                   out.ctypes.data if want_bytes else None, C.byref(k), C.byref(um)))
        res = (pos[:k.value].copy(), out[:k.value] if want_bytes else None)   # a view: no second pass over the payload
        return res + (um.value,) if return_unmerged else res

    def ring_load(self, ring_i16):
        """FRAME_FORM::form_int16_to_double on the receiver's ring: upload a host int16 ring [n, 2] once; returns the
        device address, to be passed with fmt CI16 / space DEVICE_IN (see include/cofdm.h)"""
        import numpy as np
        a = np.ascontiguousarray(ring_i16, dtype=np.int16)
        dev = C.c_void_p()
        self._chk(self.lib.cofdm_ring_load(self.h, a.ctypes.data, a.size // 2, C.byref(dev)))
        return dev.value

    def ring_find(self, ring_dev, n_samples, start=0):
        """find_t2sin then find_preamble (+1) on the resident ring; only the two 8-byte results cross PCIe"""
        pos = C.c_longlong(-1)
        self._chk(self.lib.cofdm_find_t2sin(self.h, ring_dev, CI16, n_samples, start, C.byref(pos), DEVICE_IN))
        if pos.value < 0:
            return pos.value, None
        st, first = C.c_longlong(pos.value), C.c_longlong(-10)
        self._chk(self.lib.cofdm_preamble_search(self.h, ring_dev, CI16, n_samples, C.byref(st), 1, None, C.byref(first), DEVICE_IN))
        return pos.value, first.value

    def allreduce_counters(self, nccl_comm, sums, maxes):
        """SUM of integer counters and MAX of float values over the ranks of the caller's ncclComm_t (in place)"""
        import numpy as np
        s_ = np.ascontiguousarray(sums, dtype=np.uint64)
        m_ = np.ascontiguousarray(maxes, dtype=np.float64)
        self._chk(self.lib.cofdm_allreduce_counters(self.h, nccl_comm, s_.ctypes.data if s_.size else None, s_.size,
                                                    m_.ctypes.data if m_.size else None, m_.size))
        return s_, m_

    def i16_to_cf32(self, samples):
        self._follow(samples)
        n = _n_samples(samples, CI16)
        out = self._new(samples, (n,), "complex64")
        self._chk(self.lib.cofdm_i16_to_cf32(self.h, _ptr(samples), _ptr(out), n, _space(samples, out)))
        return out
