/* cofdm.h -- C ABI of libcofdm_b200.so: the B200-native (sm_100a) implementation of the C-OFDM
 * baseband hot path.  Plain C, opaque handle, plain pointers and sizes, int status codes; no
 * exceptions, no torch/C++ types cross this boundary.
 *
 * The reference (DmSM-1/C-OFDM) has no FFI layer: its hot path is the C++ class API of
 * OFDM/Frame.hpp + OFDM/modulation.hpp compiled into main/tx/rx.  Each entry point below names the
 * reference interface it replaces (file:line relative to the reference root).  The header-compatible
 * C++ facade (c-ofdm_b200/cxx/OFDM/Frame.hpp, modulation.hpp) and the ctypes binding
 * (c-ofdm_b200/__init__.py, used by a new-style python_code/ofdm.py) both sit on exactly these calls;
 * INTEGRATION.md shows the bindings.
 *
 * Conventions
 *   status     0 = COFDM_OK, negative = error; text via cofdm_last_error() (thread local).
 *   space      every data pointer of a call lives in the memory space given by `space`:
 *              COFDM_HOST   host memory; the library copies it to device staging buffers and copies the results
 *                           back (H2D/D2H inside the call, call returns when done).  The caller's pointers go
 *                           straight to cudaMemcpyAsync: PINNED host memory (cudaHostAlloc, torch pin_memory)
 *                           gives asynchronous, full-rate, double-buffered copies; pageable memory works too,
 *                           at the driver's staged-copy rate.
 *              COFDM_DEVICE device memory of the handle's GPU; the call only enqueues kernels on the
 *                           handle's stream (cofdm_set_stream) and returns.
 *              COFDM_DEVICE_IN  the SAMPLE input is device memory (e.g. the ring of cofdm_ring_load), every other
 *                           pointer of the call (results, taps, starts) is host memory: per-frame calls on a
 *                           resident capture move only their small results over PCIe.
 *   samples    COFDM_CF32 interleaved float32 (re,im)  |  COFDM_CI16 interleaved int16 (I,Q), the
 *              SDR wire format of FRAME_FORM::get_int16 / from_sdr_int16_buf.
 *   threading  one handle per host thread / CUDA stream (the reference's FRAME_FORM is not
 *              thread-safe either); different handles are independent.
 *   no CPU fallback: every entry point that computes fails with COFDM_ERR_CUDA when no sm_100
 *              device is usable.
 */
#ifndef COFDM_B200_H
#define COFDM_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define COFDM_OK 0
#define COFDM_ERR_CONFIG (-1)      /* config file unreadable / malformed / unsupported sizes */
#define COFDM_ERR_CUDA (-2)        /* CUDA runtime error (no device, launch failure, ...) */
#define COFDM_ERR_ARG (-3)         /* bad argument */
#define COFDM_ERR_UNSUPPORTED (-4) /* configuration not covered by the kernels built so far */

#define COFDM_HOST 0
#define COFDM_DEVICE 1
#define COFDM_DEVICE_IN 2
#define COFDM_CF32 0
#define COFDM_CI16 1

#define COFDM_NOT_FOUND_T2SIN (-1)     /* T2SIN_FORM::find_t2sin sentinel, OFDM/Frame.hpp:196 */
#define COFDM_NOT_FOUND_PREAMBLE (-10) /* PREAMBLE_FORM::find_preamble sentinel, OFDM/Frame.cpp:377 */

typedef struct cofdm cofdm_t;

/* sizes of one configuration; names follow the reference's members */
typedef struct {
    int fft_size, num_data_subc, num_pilot_subc, cp_size, num_symb, num_pr_symb;
    int pr_sin_len, t2sin_size, mod_type;
    int ofdm_len;      /* OFDM_FORM::ofdm_len                          OFDM/Frame.cpp:168 */
    int rx_len;        /* message_with_preamble.size                   OFDM/Frame.cpp:169 */
    int output_size;   /* FRAME_FORM::output_size (samples per frame)  OFDM/Frame.cpp:224 */
    int usefull_size;  /* FRAME_FORM::usefull_size (bytes per frame)   OFDM/Frame.cpp:223 */
    int constell_size; /* message.usefull_size (points per frame)      OFDM/Frame.cpp:170 */
    int cor_size;      /* PREAMBLE_FORM::cor.size()                    OFDM/Frame.cpp:266 */
    int mult, rx_buf_size, iterations;
    int fused_path;    /* 1: fused fft-512 kernels; 0: generic any-size path (slower, generic.cuh); -1: tx/rx unsupported */
    int device;
} cofdm_sizes;

/* optional outputs of cofdm_rx_aligned_batch for parity checks (same `space` as the call; any
 * pointer may be NULL): per frame scal[48] = {shift, a, b, theta, g, shift*pf_den, 0, 0, ...; [16+s] = whole-bin
 * part m_s of symbol s's rotation; [32+s] = Arg of symbol s's raw CP correlation in turns},
 * grid[num_symb*fft_size], chan[num_data_subc], constell[constell_size], synced[rx_len] (complex64 each).
 * `synced` needs `scal` (a second small kernel completes it from the per-symbol scalars). */
typedef struct {
    float *scal;
    float *grid;
    float *chan;
    float *constell;
    float *synced;
} cofdm_rx_taps;

const char *cofdm_last_error(void);
const char *cofdm_version(void);

/* FRAME_FORM::FRAME_FORM(const std::string&)  OFDM/Frame.cpp:213-232 (+ parse_config,
 * config/parser.cpp:4-33): reads the SAME config.txt, builds every frame-invariant table in fp64 on
 * the host and uploads it.  device = CUDA ordinal. */
int cofdm_create(const char *config_path, int device, cofdm_t **out);
void cofdm_destroy(cofdm_t *h);
int cofdm_query(const cofdm_t *h, cofdm_sizes *out);
/* stream used for COFDM_DEVICE calls and single-stream COFDM_HOST calls: a cudaStream_t passed as
 * void*; NULL selects CUDA's default stream.  A new handle uses its own non-blocking stream, which
 * cofdm_own_stream() returns so that it can be restored. */
int cofdm_set_stream(cofdm_t *h, void *cuda_stream);
void *cofdm_own_stream(const cofdm_t *h);
int cofdm_synchronize(cofdm_t *h);

/* frame-invariant constants in fp64 as the reference holds them (host pointers, any may be NULL):
 * T2SIN tone [2*t2sin_size] (Frame.cpp:139-154), preamble bytes (Frame.cpp:269-272), ofdm_preamble
 * [2*preamble size] (Frame.cpp:282), mod_preamble (Frame.cpp:283), conjected_sinh_part [2*pr_sin_len]
 * (Frame.cpp:285-293), constellation [2<<mod_type] (modulation.cpp:23-36) */
int cofdm_get_constants(const cofdm_t *h, double *t2sin_tone, uint8_t *preamble_bytes, double *ofdm_preamble,
                        double *mod_preamble, double *matched, double *constell);

/* Modulation::mod  OFDM/modulation.cpp:39-50 : bytes -> ceil(8*n_bytes/mod_type) complex64 points */
int cofdm_mod(cofdm_t *h, int mod_type, const uint8_t *bytes, size_t n_bytes, float *points, int space);
/* Modulation::demod  OFDM/modulation.cpp:53-87 : points -> ceil(n_points*mod_type/8) bytes.  The input
 * is NOT clamped in place (the reference clamps its argument, :68-73; callers never read it back).
 * *ambiguous (host pointer, may be NULL) += points within 2e-4 level units of a decision boundary. */
int cofdm_demod(cofdm_t *h, int mod_type, const float *points, size_t n_points, uint8_t *bytes,
                unsigned long long *ambiguous, int space);

/* FRAME_FORM::write + get / get_int16  OFDM/Frame.cpp:235-256 (OFDM_FORM::write :185-198,
 * FFT_FORM::write :54-70), batched: payload[n_frames*usefull_size] -> frames[n_frames*output_size]
 * complete frames [T2SIN | preamble | message] in `fmt`.  Any sample-aligned `frames` pointer works; a 16-byte
 * aligned one (cudaMalloc, torch) lets every symbol leave the SM as one TMA bulk store. */
int cofdm_tx_batch(cofdm_t *h, const uint8_t *payload, size_t n_frames, void *frames, int fmt, int space);

/* The aligned-frame receive chain of main.cpp:60-80 / rx.cpp:200-220, batched and fused:
 * pilot_freq_sinh (Frame.hpp:285-337) -> freq_shift (:340-348) -> cp_freq_sinh (:238-263) ->
 * pr_phase_sinh (:265-274) -> chan_char_lq (:389-434) -> message.fft (:276-282, Frame.cpp:73-96) ->
 * equalise (rx.cpp:214-216) -> Mod.demod (modulation.cpp:53-87).
 * samples: n_frames records of rx_len samples starting at the preamble (what the apps copy to
 * buf + t2sin.size, rx.cpp:192-196), consecutive records frame_stride samples apart
 * (frame_stride >= rx_len; pass output_size with samples+t2sin_size to read whole frames in place).
 * bytes[n_frames*usefull_size] (4-byte aligned).  *ambiguous as in cofdm_demod.  Records may start at any sample;
 * when `samples` and the record stride are 16-byte aligned the symbols are staged by TMA bulk copies (cf32, or raw
 * int16 widened when read), otherwise by plain loads. */
int cofdm_rx_aligned_batch(cofdm_t *h, const void *samples, int fmt, size_t n_frames, size_t frame_stride,
                           uint8_t *bytes, unsigned long long *ambiguous, const cofdm_rx_taps *taps, int space);

/* FRAME_FORM::read(void*)  OFDM/Frame.cpp:239-242 -> OFDM_FORM::read :201-208, batched: whole frames
 * [n_frames*output_size] in, NO synchronisation or channel correction: CP strip, FFT, pilot normalisation and
 * segment correction (FFT_FORM::read, Frame.cpp:73-96), hard demap -> bytes[n_frames*usefull_size].
 * Optional outputs (same space, may be NULL): restored[n_frames*constell_size] complex64 = restored_buf before
 * the demap, chan_char[n_frames*num_data_subc] complex64 = PREAMBLE_FORM::chan_char() (Frame.hpp:375-385) of
 * the frame's preamble as it stands. */
int cofdm_read_batch(cofdm_t *h, const void *frames, int fmt, size_t n_frames, uint8_t *bytes,
                     unsigned long long *ambiguous, float *restored, float *chan_char, int space);

/* T2SIN_FORM::corr / find_t2sin block metric  OFDM/Frame.hpp:96-197: rel[i] = masked / total spectral
 * energy of the 256-sample block starting at start + i*t2sin_size, i < (n_samples-start)/t2sin_size;
 * 0 for blocks the reference skips (zero / NaN energy).  No threshold applied. */
int cofdm_t2sin_metric(cofdm_t *h, const void *samples, int fmt, size_t n_samples, size_t start, float *rel, int space);
/* T2SIN_FORM::find_t2sin  OFDM/Frame.hpp:150-197: first block start with rel > T2_sin_level, or -1 */
int cofdm_find_t2sin(cofdm_t *h, const void *samples, int fmt, size_t n_samples, size_t start, long long *pos, int space);

/* PREAMBLE_FORM::find_corr / find_preamble  OFDM/Frame.cpp:297-378 for n_starts candidate positions:
 * cor[n_starts*cor_size] (may be NULL) and first[n_starts] = first lag index (absolute sample index)
 * whose normalised correlation exceeds pr_level, or -10.  starts/first are int64 in `space`. */
int cofdm_preamble_search(cofdm_t *h, const void *samples, int fmt, size_t n_samples, const long long *starts,
                          size_t n_starts, float *cor, long long *first, int space);

/* The acquisition loop of the streaming receiver, rx.cpp:101-235, over an in-memory int16 capture (HOST
 * pointer; interleaved I,Q) that stands in for consecutive SDR::recv blocks of output_size*rx_buf_size
 * samples: ring of rx_buf_size+1 frames with carry-over (rx.cpp:116,147-156,180-189), find_t2sin -> find_preamble
 * (+1) -> copy rx_len samples -> pos += message.size, sentinels -1 / -10 handled as rx.cpp:137-178 does, at
 * most `iterations` (config) loop turns.  The state machine runs on the device (one CTA, stream.cuh), the frames
 * found are gathered and demodulated in one batch.  pr_begin_abs[i] = absolute sample
 * index of frame i's preamble in the capture; bytes[i*usefull_size ...]; *n_found = number of frames. */
int cofdm_rx_stream(cofdm_t *h, const int16_t *capture, size_t n_samples, size_t max_frames,
                    long long *pr_begin_abs, uint8_t *bytes, size_t *n_found);

/* The same loop over a capture cut into `n_shards` contiguous ranges of whole SDR blocks (+ one overlap block each),
 * every range scanned by its own CTA running the state machine of rx.cpp:126-198 on the device (stream.cuh), the
 * per-range chains merged on the host: the earlier range's chain is followed until it meets a preamble position the
 * later range also found (two chains coincide from the first frame both detect); a frame belongs to the range
 * that contains its preamble.  n_shards = 1 is exactly cofdm_rx_stream.  space: COFDM_HOST = `capture` and `bytes` are host
 * memory (the capture is uploaded once); COFDM_DEVICE_IN = `capture` on the device, `bytes` on the host; COFDM_DEVICE =
 * `capture` AND `bytes` on the device (the payloads never cross PCIe; 4-byte aligned, max_frames * usefull_size bytes).
 * pr_begin_abs and the counters are always host memory.
 * *n_unmerged = boundaries whose chains did not meet inside the overlap (0 in practice). */
int cofdm_rx_stream_sharded(cofdm_t *h, const int16_t *capture, size_t n_samples, int space, int n_shards,
                            size_t max_frames, long long *pr_begin_abs, uint8_t *bytes, size_t *n_found,
                            size_t *n_unmerged);

/* FRAME_FORM::form_int16_to_double  OFDM/Frame.hpp:472-481 (fp32 on the device) */
int cofdm_i16_to_cf32(cofdm_t *h, const int16_t *in, float *out, size_t n_samples, int space);

/* FRAME_FORM::form_int16_to_double on the receiver's ring  OFDM/Frame.hpp:472-481 (rx.cpp:89,110 call it once per
 * SDR block): uploads from_sdr_int16_buf (host, n_samples int16 I,Q pairs) into a device buffer owned by the handle and
 * returns its device address.  The searches and the receive chain then read it in place with fmt = COFDM_CI16,
 * space = COFDM_DEVICE_IN (`ring_dev + 2 * sample_index`), so a frame costs three small result copies instead of
 * three uploads of the whole ring. */
int cofdm_ring_load(cofdm_t *h, const int16_t *ring_host, size_t n_samples, const int16_t **ring_dev);

/* the one collective of a multi-GPU job -- SUM of n_sum integer counters (bit errors, bits, frames,
 * samples ...) and MAX of n_max values (elapsed times) over the ranks of an NCCL communicator, in place, host arrays.
 * nccl_comm is the caller's ncclComm_t (passed as void*) whose rank uses this handle's device; the library resolves
 * ncclAllReduce from the libnccl.so.2 already in the process (dlopen), it does not link NCCL. */
int cofdm_allreduce_counters(cofdm_t *h, void *nccl_comm, unsigned long long *sum_counters, size_t n_sum,
                             double *max_values, size_t n_max);

/* Device time of the kernels of the last COFDM_DEVICE call on this handle (and of the single-stream COFDM_HOST calls: the
 * searches), measured with CUDA events on the handle's stream (milliseconds; < 0 if timing is disabled or the last call
 * was a chunked COFDM_HOST tx/rx pipeline, which runs on several streams). */
int cofdm_enable_timing(cofdm_t *h, int on);
float cofdm_last_kernel_ms(const cofdm_t *h);
/* Per-stage times (milliseconds) of the work this handle did since cofdm_enable_timing(h, 1) / since the start of the last
 * cofdm_rx_stream* call -- the measured counterpart of the per-stage trace rx.cpp:32-36,128-235 prints.  CUDA events on
 * the launching stream, except MERGE (host wall clock).  Stage -> rx.cpp trace keys: SCAN = T2SIN + the untraced preamble
 * search (rx.cpp:133,161); ACQUIRE = PILOT_SINH + FREQ_PHASE_SINH + the channel fit (rx.cpp:202-211); DEMOD = PFC + the
 * demap half of MAC (rx.cpp:212-220); UPLOAD = CONVERT (form_int16_to_double); GATHER = the 5760-sample copy (rx.cpp:192-196). */
#define COFDM_STAGE_UPLOAD 0
#define COFDM_STAGE_SCAN 1
#define COFDM_STAGE_MERGE 2
#define COFDM_STAGE_GATHER 3
#define COFDM_STAGE_ACQUIRE 4
#define COFDM_STAGE_DEMOD 5
#define COFDM_STAGE_D2H 6
#define COFDM_STAGE_COUNT 7
int cofdm_last_stage_ms(cofdm_t *h, float *out, int n);
/* number of kernels this library has launched on the handle since creation */
unsigned long long cofdm_launch_count(const cofdm_t *h);

#ifdef __cplusplus
}
#endif
#endif
