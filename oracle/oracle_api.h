/* oracle_api.h -- C ABI shared by the two CPU checkers of this repo.  TEST INFRASTRUCTURE ONLY.
 *
 *   oracle/_ref/libcofdm_ref.so   the UNMODIFIED reference sources (/root/reference/OFDM/{Frame,modulation}.cpp,
 *                                 config/parser.cpp) compiled where they lie, wrapped by
 *                                 oracle/ref_shim.cpp, linked against the stand-in FFT
 *                                 (oracle/standin).  Only buildable where /root/reference exists.
 *   oracle/libcofdm_oracle.so     oracle/cofdm_oracle.c: a plain-C double-precision restatement
 *                                 of the same algorithms (each function cites the reference
 *                                 file:line it follows).  Builds anywhere gcc is.
 *
 * Both export exactly the functions below, so every parity test can run against either.
 * Nothing under c-ofdm_b200/ (the product) may include, link or load any of this; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Conventions: complex arrays are interleaved double (re, im); "frame" = output_size complex
 * samples laid out [T2SIN | preamble | message] as FRAME_FORM::buf (reference OFDM/Frame.cpp:219-231).
 */
#ifndef COFDM_ORACLE_API_H
#define COFDM_ORACLE_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oc_handle oc_handle;

typedef struct {
    int fft_size, num_data_subc, num_pilot_subc, cp_size, num_symb, num_pr_symb;
    int pr_sin_len, pr_seed, t2sin_size, t2_f1, t2_f2, smooth, mod_type;
    int ofdm_len;        /* fft_size + cp_size                         Frame.cpp:168 */
    int preamble_size;   /* ofdm_len * num_pr_symb                     Frame.cpp:169 */
    int message_size;    /* ofdm_len * num_symb                                      */
    int output_size;     /* t2sin_size + preamble_size + message_size  Frame.cpp:219,224 */
    int usefull_size;    /* payload bytes per frame                    Frame.cpp:223 */
    int constell_size;   /* num_data_subc * num_symb                   Frame.cpp:170 */
    int mult, rx_buf_size, iterations;
    int cor_size;        /* 2*t2sin_size + pr_sin_len lags             Frame.cpp:266 */
    double t2_level, pr_level, pilot_ampl;
} oc_sizes;

/* which implementation answers: "reference" (libcofdm_ref) or "port" (libcofdm_oracle) */
const char *oc_kind(void);

oc_handle *oc_create(const char *config_path);           /* NULL + message via oc_last_error() */
void oc_destroy(oc_handle *h);
const char *oc_last_error(void);
void oc_get_sizes(const oc_handle *h, oc_sizes *out);

/* frame-invariant constants of the config; any pointer may be NULL.
 *  t2sin_tone[2*t2sin_size], t2_mask[t2sin_size], preamble_bytes[num_data_subc*num_pr_symb/8],
 *  ofdm_preamble[2*preamble_size], mod_preamble[2*num_data_subc*num_pr_symb],
 *  matched[2*pr_sin_len], constell[2<<mod_type]                                              */
void oc_get_constants(oc_handle *h, double *t2sin_tone, double *t2_mask, uint8_t *preamble_bytes,
                      double *ofdm_preamble, double *mod_preamble, double *matched, double *constell);

/* Modulation (reference OFDM/modulation.cpp).  mod_type: 1,2,4,6,8. */
int oc_bit_stream_converter(int out_bits, int in_bits, const uint8_t *in, int n_in, uint8_t *out);
int oc_mod(int mod_type, const uint8_t *bytes, int n_bytes, double *points);           /* -> n points */
int oc_demod(int mod_type, double *points_inout, int n_points, uint8_t *bytes);        /* -> n bytes; clamps in place */

/* TX: FRAME_FORM::write + get (+ get_int16 when frame_i16 != NULL).  bytes[usefull_size]. */
void oc_tx(oc_handle *h, const uint8_t *bytes, double *frame, int16_t *frame_i16);

/* Sync (T2SIN_FORM::corr / find_t2sin, PREAMBLE_FORM::find_corr / find_preamble). */
int oc_t2sin_corr(oc_handle *h, const double *sig, long n, double *out);   /* -> n/t2sin_size values */
int oc_find_t2sin(oc_handle *h, const double *sig, long n, int start);
void oc_find_corr(oc_handle *h, const double *sig, long n, int start, double *cor /*cor_size*/);
int oc_find_preamble(oc_handle *h, const double *sig, long n, int start);

/* Full aligned rx chain in the call order of reference main.cpp:60-80 / rx.cpp:200-220.
 * rx_samples: preamble_size+message_size complex samples starting at the preamble (what the
 * apps copy to buf + t2sin.size).  Outputs (any may be NULL):
 *   scal[0]=coarse shift (pilot_freq_sinh), scal[1]=a, scal[2]=b of chan_char_lq (recovered
 *           from chan_est), scal[3]=pr_phase angle is not observable -> 0
 *   synced[2*(preamble+message)]  samples after freq_shift+cp_freq_sinh+pr_phase_sinh
 *   grid[2*num_symb*fft_size]     message FFT_buf after FFT_FORM::read (normalised bins)
 *   chan[2*num_data_subc]         chan_char_lq()
 *   constell[2*constell_size]     equalised points BEFORE demod's in-place clamp
 *   bytes[usefull_size]           Modulation::demod                                          */
void oc_rx_aligned(oc_handle *h, const double *rx_samples, double *scal, double *synced,
                   double *grid, double *chan, double *constell, uint8_t *bytes);

/* FRAME_FORM::read(void*): sync-less demodulation of a whole frame (Frame.cpp:239-242).
 * restored (may be NULL) receives FFT_FORM::read()'s restored_buf before demod clamps it. */
void oc_read(oc_handle *h, const double *frame, double *restored, uint8_t *bytes);

/* PREAMBLE_FORM::chan_char on the preamble currently in `rx_samples` (no sync applied). */
void oc_chan_char(oc_handle *h, const double *rx_samples, double *chan);

/* Streaming receiver: the acquisition state machine of reference rx.cpp:101-235 replayed over an
 * in-memory int16 capture (interleaved I,Q) that stands in for consecutive SDR::recv blocks of
 * output_size*rx_buf_size samples.  Stops when the capture is exhausted, `iterations` loop
 * turns were made, or max_frames frames were decoded.  pr_begin_abs[i] = absolute sample index
 * (in the capture) of frame i's preamble start; bytes[i*usefull_size...].  Returns frame count. */
int oc_rx_stream(oc_handle *h, const int16_t *capture, long n_samples, int max_frames,
                 long *pr_begin_abs, uint8_t *bytes);

/* CPU-baseline loop (bench.py): for each of n_frames payloads do FRAME_FORM::write -> get_int16 ->
 * (int16 -> double, Frame.hpp:472-481) -> the aligned rx chain of main.cpp:60-80 -> bytes.
 * Everything runs inside the library so a caller thread holds no interpreter lock.  Returns the number
 * of payload bytes that did not come back. */
long oc_txrx_loop(oc_handle *h, const uint8_t *payloads, int n_frames, uint8_t *bytes_out);

#ifdef __cplusplus
}
#endif
#endif
