// params.h -- plain-old-data description of one modem configuration, shared by host and device.
// Everything here is derived from the reference's config file (config/config.txt keys, consumed
// at OFDM/Frame.cpp:99-105,157-176,213-226,259-267) by cofdm_host.cu; the kernels never parse anything.
#pragma once
#include <stdint.h>

namespace cofdmk {

constexpr int kMaxPilots = 128;
constexpr int kGenMaxSym = 32;      // frame symbols (preamble + message) the any-size path keeps per-symbol scalars for (GenFrame)
constexpr int kMaxFusedSymb = 15;   // message symbols per frame the fused 512 kernels accept (one warp each)

struct Params {
    // ---- config keys (names as in config/config.txt) ----
    int fft_size, num_data_subc, num_pilot_subc, cp_size, num_symb, num_pr_symb;
    int pr_sin_len, t2sin_size, mod_type;
    float mult;            // int16 scale of FRAME_FORM::get_int16 (Frame.cpp:252)
    float pilot_ampl;      // pilot_ampl/1000 (Frame.cpp:172)
    float t2_level;        // T2_sin_level/1000 (Frame.cpp:105)
    float pr_level;        // pr_level/1000 (Frame.cpp:261)
    // ---- derived sizes (Frame.cpp:9-10,168-170,219-225,266) ----
    int ofdm_len;          // fft_size + cp_size
    int n_sym_rx;          // num_pr_symb + num_symb (message_with_preamble)
    int rx_len;            // ofdm_len * n_sym_rx   samples the rx chain consumes per frame
    int frame_len;         // t2sin_size + rx_len   (output_size)
    int seg_size, seg_step;
    int bytes_per_frame;   // usefull_size
    int pts_per_sym;       // num_data_subc
    int cor_size;          // 2*t2sin_size + pr_sin_len lags of find_preamble
    // ---- coarse-CFO windows of pilot_freq_sinh on the preamble (Frame.hpp:311-334) ----
    int pf_size;           // preamble size = ofdm_len*num_pr_symb
    int pf_pilot_w;        // window width in bins
    int pf_border0;        // first border (clamped at 0)
    // shift = (sum_argmax / num_pilot_subc - pf_size/2) / pf_size  ==  pf_num / pf_den cycles/sample,
    // pf_num = sum_argmax - num_pilot_subc*(pf_size/2), pf_den = num_pilot_subc*pf_size
    int pf_den;
    float pf_bins512;      // 512 / pf_den: whole fft-512 bins per unit of the coarse-shift numerator
    float pf_binsN;        // fft_size / pf_den: the same for the configuration's own transform size (big.cuh)
    float inv_pilot_norm;  // 1 / (num_symb * num_pilot_subc * pilot_ampl): pilot amplitude normaliser (Frame.cpp:76-80)
    // ---- radix schedules of the generic (any-size) path: products equal fft_size resp. pf_size ----
    int fft_nr, fft_radix[8];
    int pf_nr, pf_radix[8];
    // ---- device tables (all fp32 roundings of host fp64 values) ----
    const float2 *tw_fft;        // [fft_size]  exp(-j*2*pi*k/fft_size)
    const float2 *tw_pf;         // [pf_size]   exp(-j*2*pi*k/pf_size)
    const float2 *tw_t2;         // [t2sin_size]
    const float *t2_mask;        // [t2sin_size] detect_mask (Frame.cpp:120-133)
    const float2 *t2_tone;       // [t2sin_size] sync tone (Frame.cpp:139-154)
    const float2 *preamble_td;   // [pf_size]   ofdm_preamble (Frame.cpp:282)
    const float2 *matched;       // [pr_sin_len] conjected_sinh_part (Frame.cpp:285-293)
    const float2 *mod_preamble;  // [num_data_subc*num_pr_symb] (Frame.cpp:283)
    const float2 *constell;      // [1<<mod_type] (modulation.cpp:23-36)
    const int16_t *bin_map;      // [fft_size] -1 null, -2 pilot, else data index within the symbol
    const int16_t *data_bin;     // [num_data_subc] bin of data index i
    const int16_t *pilot_bin;    // [num_pilot_subc]
    int big_tmask;               // big.cuh: bit t set <=> some bin j + 256 t is used (data or pilot)
    int big_phmask;              // big.cuh: bit t set <=> some bin j + 256 t carries a data index below num_data_subc / 2 (chan_char_lq's bins)
    int big_dstep;               // big.cuh: data-index step between consecutive rows of one thread (0: none)
    int big_lay;                 // big.cuh: 1 = the sub-carrier map has the row layout the specialised demod instance is compiled for (kBigLay*)
    const uint4 *big_eq;         // [256][3] big.cuh, big_lay only: per thread .x .. .w of entries 0 / 1 = its data rows 0..3 / 12..15: [15:0] byte the
                                 //      demapped symbol goes to (data index, or a dump slot behind the data), [31:16] byte offset of the segment
                                 //      coefficient; entry 2 .x: the (extrapolated) channel-line abscissa i' at rows 0 [15:0] and 12 [31:16], signed
    const uint4 *big_txd;        // [256][4] big.cuh, big_tx_kernel: per thread one word per row (bins j + 256 u): [15:0] byte of the symbol's payload where
                                 //      the sub-carrier's bits start, [18:16] bit offset in that byte, [25:24] 1 = pilot, 2 = null
    const uint4 *big_roles;      // [256][2] big.cuh: the 16 roles of thread j (bins j + 256 t, t = 0..15) as int16, packed (fft 4096 only)
    const int16_t *bin_role;     // [fft_size] >= 0 data index within the symbol, -1 null, -2 - p pilot number p (big.cuh)
    // ---- one-warp-per-symbol receive kernels of the fft-512 geometry (rx512n.cuh) ----
    // After warp_fft512 lane 2 k1 + g holds mn[i] = X[k1 + 16 i + (g ? 384 : 0)] and ot[i] = X[k1 + 16 i + (g ? 128 : 256)].
    // A COMBINATION is a set of data bins that share (array, g, i, segment): inside it the data index i' (minus 256 in the
    // negative half: the channel line's abscissa, Frame.hpp:425-430) is k1 + off.
    const uint4 *lane_desc;      // [32] per lane 8 x 16 bits for mn[0..7]: [15:7] data index in the symbol (256: no data), [6:2] combination
    const uint2 *lane_aux;       // [32] .x: combination number `lane`: [2:0] segment, [31:16] off (signed); and the LANE's routing: [9:8] 1 = it holds
                                 //      pilot [7:4] (in mn[], or in its used ot[] register), 2 = it holds straggler bin [7:4]; .y (lanes 0..6): the straggler
                                 //      data bin number `lane` (bins 128..131, 381..383, held in ot[]): descriptor as above | (its k1) << 16
    const uint2 *acq_desc;       // [32] acquire kernel: per lane 4 x 16 bits (even lane: mn[0..3]; odd lane: mn[4..7] of lane - 1): [7:0] index of
                                 //      the phase the slot produces (0..127), [15] take the product of straggler bin 128 + [9:8] instead
    const uint4 *tx_desc;        // [32][2] tx512w.cuh: per lane 10 x 16 bits, the grid rows n1 = 0..4, 11..15 of bins lane + 32 n1: where the symbol's
                                 //      bits sit in the payload ([7:0] byte, [11:8] shift), 0x4000 = null, 0x8000 = pilot (rows 5..10 are never used)
    const float2 *grid_lane;     // [32][10] conj(tx grid of the preamble) / sqrt(fft_size) at the bins of mn[0..7] and of the used ot[] register
    int n_combos;
};

// Optional debug/parity taps of the fused rx kernel (device pointers, any may be null).
struct RxTaps {
    float *scal;       // [n][48]: shift, a, b, theta, g, kc, ..; [16+s] m_s; [32+s] Arg(C_s) (turns)
    float2 *grid;      // [n][num_symb*fft_size]   normalised message bins (FFT_FORM::read's FFT_buf)
    float2 *chan;      // [n][num_data_subc]       chan_char_lq
    float2 *constell;  // [n][num_data_subc*num_symb] equalised points before the demap clamp
    float2 *synced;    // [n][rx_len]              samples after the three time-domain corrections
};

enum SampleFormat { kCF32 = 0, kCI16 = 1 };

}  // namespace cofdmk
