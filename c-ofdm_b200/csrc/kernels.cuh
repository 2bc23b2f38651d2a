// kernels.cuh -- the sm_100a kernels of the C-OFDM hot path.
//
//   rx_acquire512w_kernel + rx_demod512_kernel (rx512n.cuh)   aligned frame -> payload bytes, every sample read once:
//                        coarse CFO (pilot_freq_sinh), fine CFO (cp_freq_sinh), preamble phase lock
//                        (pr_phase_sinh), 9 FFT-512, pilot normalisation + segment correction
//                        (FFT_FORM::read), linear-phase channel fit (chan_char_lq), equalise, hard demap.
//   tx512w_kernel (tx512w.cuh)   payload bytes -> QAM map -> pilot insertion -> 8 IFFT-512 -> CP -> frame
//   t2sin_metric_kernel  sync-tone block detector metric (T2SIN_FORM::corr / find_t2sin)
//   preamble_corr_kernel tiled sliding dot product + sliding energy (find_corr / find_preamble)
//   mod_kernel / demod_kernel, int16 <-> cf32 converters
//
// The "512" kernels are specialised for fft_size 512, cp 128, 8 pilots, 256 data sub-carriers
// (segment = one warp of 32 bins) and one preamble symbol: the reference's shipped config.txt.
// One CTA per frame, one warp per OFDM symbol.  See DESIGN.md for the data layout and rooflines.
#pragma once
#include "compat.cuh"
#include "params.h"
#include "fft.cuh"
#include "modem.cuh"
#include "rx512n.cuh"
#include "generic.cuh"

namespace cofdmk {

// wide: the frame buffer is 16-byte aligned (one 16- / 8-byte store for the pair); otherwise one store per sample
template <int FMT>
COFDM_DEV void store_sample_pair(void *frame_out, int idx /*even sample index*/, float2 a, float2 b, float mult, bool wide) {
    if (FMT == kCI16) {
        // Frame.cpp:252: int16(trunc(re*mult)), int16(trunc(im*mult))
        short2 sa = make_short2((short)__float2int_rz(a.x * mult), (short)__float2int_rz(a.y * mult));
        short2 sb = make_short2((short)__float2int_rz(b.x * mult), (short)__float2int_rz(b.y * mult));
        uint2 pk;
        pk.x = ((unsigned)(unsigned short)sa.x) | ((unsigned)(unsigned short)sa.y << 16);
        pk.y = ((unsigned)(unsigned short)sb.x) | ((unsigned)(unsigned short)sb.y << 16);
        if (wide) reinterpret_cast<uint2 *>(frame_out)[idx >> 1] = pk;
        else { reinterpret_cast<unsigned *>(frame_out)[idx] = pk.x; reinterpret_cast<unsigned *>(frame_out)[idx + 1] = pk.y; }
    } else {
        if (wide) reinterpret_cast<float4 *>(frame_out)[idx >> 1] = make_float4(a.x, a.y, b.x, b.y);
        else { reinterpret_cast<float2 *>(frame_out)[idx] = a; reinterpret_cast<float2 *>(frame_out)[idx + 1] = b; }
    }
}

// rel = sum(mask*|X|^2) / sum(|X|^2) of the 256 samples a warp has staged in A at slot i (unpadded; B = scratch, both
// kT2Slots long); every lane returns it.
// Blocks with zero or NaN energy report 0 (Frame.hpp:132-138 `continue`).
COFDM_DEV float t2sin_block_rel(const Params &P, float2 *A, float2 *B, int lane) {
    // total spectral energy by Parseval: sum_k |X_k|^2 = 256 * sum_n |x_n|^2 (saves evaluating unmasked bins)
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) tot += cnorm2(A[lane + 32 * i]);
    tot *= 256.0f;
    __syncwarp();
    stockham_pass<8, false, false, 0, 4>(A, B, 256, 1, P.tw_t2, lane, 32);
    __syncwarp();
    stockham_pass<8, false, true, 4, 64>(B, A, 256, 8, P.tw_t2, lane, 32);  // twiddle index <= 7*7*4 < 256
    __syncwarp();
    // last pass (radix 4, ns = 64): butterfly j yields bins j, j+64, j+128, j+192; only butterflies that feed a
    // masked bin are evaluated (the shipped mask covers bins 12..22 and 46..56: 22 of 64 butterflies)
    float sine = 0.f;
#pragma unroll
    for (int jj = 0; jj < 2; jj++) {
        const int j = lane + 32 * jj;
        float mk[4];
        bool any = false;
#pragma unroll
        for (int q = 0; q < 4; q++) { mk[q] = __ldg(&P.t2_mask[j + 64 * q]); any |= mk[q] != 0.f; }
        if (any) {
            float2 v[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                v[q] = A[pad_slot<64>(j + 64 * q)];
                if (q > 0) v[q] = cmul(v[q], __ldg(&P.tw_t2[q * j]));           // q*j <= 3*63 < 256
            }
            dft4<false>(v);
#pragma unroll
            for (int q = 0; q < 4; q++) sine += mk[q] * cnorm2(v[q]);
        }
    }
    tot = warp_sum(tot);
    sine = warp_sum(sine);
    float rel = sine / tot;
    if (tot == 0.f || rel != rel) rel = 0.f;
    return rel;
}

// ------------------------------------------------------------------------------------------------
// t2sin_metric_kernel: T2SIN_FORM::corr / find_t2sin block metric (Frame.hpp:112-143).
// One warp per non-overlapping 256-sample block: FFT-256 (8x8x4 Stockham in the warp's private
// shared memory), rel = sum(mask*|X|^2) / sum(|X|^2); blocks with zero or NaN energy report 0.
// ------------------------------------------------------------------------------------------------
constexpr int kT2WarpsPerCta = 8;
constexpr int kT2Slots = 288;                  // 256 samples + padding (pad_slot<4>: 271, pad_slot<64>: 280): conflict-free Stockham passes
template <int FMT>
__global__ void __launch_bounds__(32 * kT2WarpsPerCta)
t2sin_metric_kernel(const Params P, const void *__restrict__ samples, long long start, long long n_blocks,
                    float *__restrict__ rel_out) {
    __shared__ float2 buf[kT2WarpsPerCta][2][kT2Slots];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long blk = (long long)blockIdx.x * kT2WarpsPerCta + warp;
    if (blk >= n_blocks) return;
    float2 *A = buf[warp][0], *B = buf[warp][1];
    const long long s0 = start + blk * 256;
    if (FMT == kCI16) {
        const unsigned *src = reinterpret_cast<const unsigned *>(samples) + s0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const unsigned w = __ldg(src + lane + 32 * i);
            A[lane + 32 * i] = make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
        }
    } else {
        const float2 *src = reinterpret_cast<const float2 *>(samples) + s0;
#pragma unroll
        for (int i = 0; i < 8; i++) A[lane + 32 * i] = __ldg(src + lane + 32 * i);
    }
    const float rel = t2sin_block_rel(P, A, B, lane);
    if (lane == 0) rel_out[blk] = rel;
}

// t2sin_metric_any_kernel: the same metric for any power-of-two T2sin_size from 16 to 1024 (the reference takes the size from
// the configuration, Frame.cpp:99-136; 256 has the two tuned kernels here).  One warp per block, radix-8 / 4 / 2 Stockham passes
// between two buffers in the warp's shared memory, every bin evaluated.
constexpr int kT2AnyWarps = 4;
COFDM_HD size_t t2sin_any_smem_bytes(int size) { return (size_t)kT2AnyWarps * 2 * (size_t)size * sizeof(float2); }
template <int FMT>
__global__ void __launch_bounds__(32 * kT2AnyWarps)
t2sin_metric_any_kernel(const Params P, const void *__restrict__ samples, long long start, long long n_blocks, float *__restrict__ rel_out) {
    COFDM_DYN_SMEM(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = P.t2sin_size;
    const long long blk = (long long)blockIdx.x * kT2AnyWarps + warp;
    if (blk >= n_blocks) return;
    float2 *A = reinterpret_cast<float2 *>(smem_raw) + (size_t)warp * 2 * n, *B = A + n;
    const long long s0 = start + blk * n;
    float tot = 0.f;
    for (int i = lane; i < n; i += 32) {
        float2 x;
        if (FMT == kCI16) {
            const unsigned w = __ldg(reinterpret_cast<const unsigned *>(samples) + s0 + i);
            x = make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
        } else {
            x = __ldg(reinterpret_cast<const float2 *>(samples) + s0 + i);
        }
        A[i] = x;
        tot += cnorm2(x);                                       // Parseval: sum_k |X_k|^2 = n sum_i |x_i|^2
    }
    tot *= (float)n;
    __syncwarp();
    int ns = 1;
    while (ns < n) {
        const int rem = n / ns;
        if (rem % 8 == 0) { stockham_pass<8, false>(A, B, n, ns, P.tw_t2, lane, 32); ns *= 8; }
        else if (rem % 4 == 0) { stockham_pass<4, false>(A, B, n, ns, P.tw_t2, lane, 32); ns *= 4; }
        else { stockham_pass<2, false>(A, B, n, ns, P.tw_t2, lane, 32); ns *= 2; }
        __syncwarp();
        float2 *t = A; A = B; B = t;
    }
    float sine = 0.f;
    for (int k = lane; k < n; k += 32) sine += __ldg(&P.t2_mask[k]) * cnorm2(A[k]);
    tot = warp_sum(tot);
    sine = warp_sum(sine);
    float rel = sine / tot;
    if (tot == 0.f || rel != rel) rel = 0.f;                    // Frame.hpp:132-138 `continue`
    if (lane == 0) rel_out[blk] = rel;
}

// t2sin_metric2_kernel: the same metric with TWO consecutive blocks per warp, packed f32x2 (block 2b in the low,
// block 2b+1 in the high half of every register pair): half the FFT instructions per block.
constexpr int kT2PairWarps = 4;
template <int FMT>
__global__ void __launch_bounds__(32 * kT2PairWarps)
t2sin_metric2_kernel(const Params P, const void *__restrict__ samples, long long start, long long n_blocks,
                     float *__restrict__ rel_out) {
    __shared__ float2 buf[kT2PairWarps][4][288];                // per warp: re / im planes of two ping-pong buffers, padded (pad_slot<4> / <64>)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long blk = 2 * ((long long)blockIdx.x * kT2PairWarps + warp);
    if (blk >= n_blocks) return;
    const bool has1 = blk + 1 < n_blocks;
    float2 *Are = buf[warp][0], *Aim = buf[warp][1], *Bre = buf[warp][2], *Bim = buf[warp][3];
    const long long s0 = start + blk * 256;
    float2 tot = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float2 x0, x1 = make_float2(0.f, 0.f);
        if (FMT == kCI16) {
            const unsigned *src = reinterpret_cast<const unsigned *>(samples) + s0 + lane + 32 * i;
            const unsigned w0 = __ldg(src);
            x0 = make_float2((float)(short)(w0 & 0xffffu), (float)(short)(w0 >> 16));
            if (has1) { const unsigned w1 = __ldg(src + 256); x1 = make_float2((float)(short)(w1 & 0xffffu), (float)(short)(w1 >> 16)); }
        } else {
            const float2 *src = reinterpret_cast<const float2 *>(samples) + s0 + lane + 32 * i;
            x0 = __ldg(src);
            if (has1) x1 = __ldg(src + 256);
        }
        Are[lane + 32 * i] = make_float2(x0.x, x1.x);
        Aim[lane + 32 * i] = make_float2(x0.y, x1.y);
        tot.x += cnorm2(x0); tot.y += cnorm2(x1);              // Parseval: sum_k |X_k|^2 = 256 sum_n |x_n|^2
    }
    __syncwarp();
    stockham_pass_pc<8, false, false, 0, 4>(Are, Aim, Bre, Bim, 256, 1, P.tw_t2, lane, 32);
    __syncwarp();
    stockham_pass_pc<8, false, true, 4, 64>(Bre, Bim, Are, Aim, 256, 8, P.tw_t2, lane, 32);   // twiddle index <= 7*7*4 < 256
    __syncwarp();
    // last pass (radix 4, ns = 64), only butterflies that feed a masked bin (see t2sin_block_rel)
    float2 sine = make_float2(0.f, 0.f);
#pragma unroll
    for (int jj = 0; jj < 2; jj++) {
        const int j = lane + 32 * jj;
        float mk[4];
        bool any = false;
#pragma unroll
        for (int q = 0; q < 4; q++) { mk[q] = __ldg(&P.t2_mask[j + 64 * q]); any |= mk[q] != 0.f; }
        if (any) {
            pc v[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                v[q].re = Are[pad_slot<64>(j + 64 * q)]; v[q].im = Aim[pad_slot<64>(j + 64 * q)];
                if (q > 0) v[q] = cmul(v[q], __ldg(&P.tw_t2[q * j]));
            }
            dft4<false>(v);
#pragma unroll
            for (int q = 0; q < 4; q++) sine = p_fma(p_bcast(mk[q]), p_fma(v[q].re, v[q].re, p_mul(v[q].im, v[q].im)), sine);
        }
    }
    tot = warp_sum(tot);
    sine = warp_sum(sine);
    if (lane == 0) {
        const float t0 = tot.x * 256.0f, t1 = tot.y * 256.0f;
        float r0 = sine.x / t0, r1 = sine.y / t1;
        if (t0 == 0.f || r0 != r0) r0 = 0.f;                   // Frame.hpp:132-138 `continue`
        if (t1 == 0.f || r1 != r1) r1 = 0.f;
        rel_out[blk] = r0;
        if (has1) rel_out[blk + 1] = r1;
    }
}

// ------------------------------------------------------------------------------------------------
// preamble_corr_kernel: PREAMBLE_FORM::find_corr / find_preamble (Frame.cpp:297-378).
// One CTA per candidate start.  The window (cor_size + pr_sin_len samples) and the matched filter
// are staged in shared memory; each thread owns lags t, t+128, ... and slides the 128-tap complex
// dot product and the window energy along them.  c_i = |dot| / sqrt(E_i) when E_i > 1.
// first_idx[c] = start + (first lag with c_i > pr_level), or -10 (Frame.cpp:377).
// ------------------------------------------------------------------------------------------------
constexpr int kPcThreads = 128;
template <int FMT>
__global__ void __launch_bounds__(kPcThreads)
preamble_corr_kernel(const Params P, const void *__restrict__ samples, long long n_samples,
                     const long long *__restrict__ starts, int n_starts,
                     float *__restrict__ cor_out /* [n_starts][cor_size] or null */,
                     long long *__restrict__ first_idx /* [n_starts] or null */) {
    COFDM_DYN_SMEM(smem_raw);
    const int c = blockIdx.x;
    if (c >= n_starts) return;
    const int tid = threadIdx.x;
    const int L = P.pr_sin_len, NC = P.cor_size, WN = NC + L;
    float2 *win = reinterpret_cast<float2 *>(smem_raw);
    float2 *hf = win + WN;
    __shared__ int first;
    if (tid == 0) first = 0x7fffffff;
    const long long st = starts[c];
    for (int i = tid; i < WN; i += kPcThreads) {
        const long long g = st + i;
        float2 x = make_float2(0.f, 0.f);
        if (g >= 0 && g < n_samples) {
            if (FMT == kCI16) {
                const unsigned w = __ldg(reinterpret_cast<const unsigned *>(samples) + g);
                x = make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
            } else {
                x = __ldg(reinterpret_cast<const float2 *>(samples) + g);
            }
        }
        win[i] = x;
    }
    for (int i = tid; i < L; i += kPcThreads) hf[i] = __ldg(&P.matched[i]);
    __syncthreads();
    const float lvl2 = P.pr_level * P.pr_level;
    for (int i0 = tid; i0 < NC; i0 += 4 * kPcThreads) {
        float2 acc[4];
        float en[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { acc[q] = make_float2(0.f, 0.f); en[q] = 0.f; }
        for (int j = 0; j < L; j++) {
            const float2 hj = hf[j];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = i0 + q * kPcThreads;
                if (i < NC) {
                    const float2 x = win[i + j];
                    cmac(acc[q], x, hj);                  // Frame.cpp:320-321 (h is already conjugated)
                    en[q] += cnorm2(x);                   // Frame.cpp:305-309,327-333 sliding energy
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = i0 + q * kPcThreads;
            if (i >= NC) continue;
            float cval = 0.f;
            if (en[q] > 1.0f) {                           // Frame.cpp:319
                cval = sqrtf(cnorm2(acc[q]) / en[q]);
                if (cnorm2(acc[q]) > lvl2 * en[q]) atomicMin(&first, i);   // Frame.cpp:364
            }
            if (cor_out != nullptr) cor_out[(size_t)c * NC + i] = cval;
        }
    }
    __syncthreads();
    if (first_idx != nullptr && tid == 0) first_idx[c] = first == 0x7fffffff ? -10 : st + first;
}

// ------------------------------------------------------------------------------------------------
// The matched-filter core with 4 CONSECUTIVE lags per thread (lags 4t .. 4t+3): the window is held transposed in 4
// planes, sample idx at win4[(idx & 3) * plane + (idx >> 2)], so that the one new sample a thread needs per filter
// tap is at consecutive addresses across threads (conflict-free), and each tap costs 2 shared-memory loads for
// 4 lags instead of 5.  L (filter length) must be a multiple of 4; plane >= (4 t_max + L + 3) / 4 + 1.
// ------------------------------------------------------------------------------------------------
// hf4[j] = (h.x, h.y, -h.y, h.x) of filter tap j (h already conjugated, Frame.cpp:285-293): a complex multiply-accumulate
// a += x h is then TWO packed FFMA2 with a scalar-broadcast operand, a = fma(hf4.lo, x.x, a); a = fma(hf4.hi, x.y, a).
// The window energy E_i = sum_j |x[i + j]|^2 (Frame.cpp:305-309) is accumulated for the thread's first lag only and slid to
// its other three lags by adding the sample that enters and removing the one that leaves -- the reference's own recurrence
// (Frame.cpp:327-333).
COFDM_DEV void corr4_lags(const float2 *win4, int plane, const float4 *hf4, int L, int t, float2 (&a)[4], float (&e)[4]) {
    const float2 *w0 = win4, *w1 = win4 + plane, *w2 = win4 + 2 * plane, *w3 = win4 + 3 * plane;
    float2 x0 = w0[t], x1 = w1[t], x2 = w2[t], x3 = w3[t];
    const float f0 = cnorm2(x0), f1 = cnorm2(x1), f2 = cnorm2(x2);          // the samples that leave the window for lags 1, 2, 3
    float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    float e0 = 0.f;
#define COFDM_CMAC4(A0, X0, A1, X1, A2, X2, A3, X3, H)                                             \
    do {                                                                                           \
        const float2 hl_ = make_float2((H).x, (H).y), hh_ = make_float2((H).z, (H).w);             \
        A0 = p_fma(hl_, p_bcast((X0).x), A0); A1 = p_fma(hl_, p_bcast((X1).x), A1);                \
        A2 = p_fma(hl_, p_bcast((X2).x), A2); A3 = p_fma(hl_, p_bcast((X3).x), A3);                \
        A0 = p_fma(hh_, p_bcast((X0).y), A0); A1 = p_fma(hh_, p_bcast((X1).y), A1);                \
        A2 = p_fma(hh_, p_bcast((X2).y), A2); A3 = p_fma(hh_, p_bcast((X3).y), A3);                \
    } while (0)
    for (int j = 0; j < L; j += 4) {
        const int k = t + (j >> 2) + 1;
        float4 h = hf4[j];
        COFDM_CMAC4(a0, x0, a1, x1, a2, x2, a3, x3, h);
        e0 += cnorm2(x0);
        x0 = w0[k];
        h = hf4[j + 1];
        COFDM_CMAC4(a0, x1, a1, x2, a2, x3, a3, x0, h);
        e0 += cnorm2(x1);
        x1 = w1[k];
        h = hf4[j + 2];
        COFDM_CMAC4(a0, x2, a1, x3, a2, x0, a3, x1, h);
        e0 += cnorm2(x2);
        x2 = w2[k];
        h = hf4[j + 3];
        COFDM_CMAC4(a0, x3, a1, x0, a2, x1, a3, x2, h);
        e0 += cnorm2(x3);
        x3 = w3[k];
    }
#undef COFDM_CMAC4
    // after the loop x0, x1, x2 are the samples 4 t + L, + L + 1, + L + 2: the ones that enter the window for lags 1, 2, 3
    a[0] = a0; a[1] = a1; a[2] = a2; a[3] = a3;
    e[0] = e0;
    e[1] = e0 + (cnorm2(x0) - f0);
    e[2] = e[1] + (cnorm2(x1) - f1);
    e[3] = e[2] + (cnorm2(x2) - f2);
}

// preamble_corr4_kernel: the same search as preamble_corr_kernel on the 4-lags-per-thread core (needs pr_sin_len and
// the lag count to be multiples of 4).  One CTA per candidate start.
constexpr int kPc4Threads = 160;
COFDM_HD size_t preamble_corr4_smem_bytes(int cor_size, int pr_sin_len) {
    return (4 * ((size_t)(cor_size + pr_sin_len) / 4 + 4) + 2 * (size_t)pr_sin_len) * sizeof(float2);
}
template <int FMT>
__global__ void __launch_bounds__(kPc4Threads)
preamble_corr4_kernel(const Params P, const void *__restrict__ samples, long long n_samples,
                      const long long *__restrict__ starts, int n_starts,
                      float *__restrict__ cor_out /* [n_starts][cor_size] or null */,
                      long long *__restrict__ first_idx /* [n_starts] or null */) {
    COFDM_DYN_SMEM(smem_raw);
    const int c = blockIdx.x;
    if (c >= n_starts) return;
    const int tid = threadIdx.x;
    const int L = P.pr_sin_len, NC = P.cor_size, WN = NC + L, plane = WN / 4 + 4;
    float2 *win4 = reinterpret_cast<float2 *>(smem_raw);
    float4 *hf = reinterpret_cast<float4 *>(win4 + 4 * (size_t)plane);       // (plane is a multiple of 2 slots: 16-byte aligned)
    __shared__ int first;
    if (tid == 0) first = 0x7fffffff;
    const long long st = starts[c];
    for (int i = tid; i < 4 * plane; i += kPc4Threads) {
        const int q = i / plane, k = i - q * plane, idx = 4 * k + q;
        const long long g = st + idx;
        float2 x = make_float2(0.f, 0.f);
        if (idx < WN && g >= 0 && g < n_samples) {
            if (FMT == kCI16) {
                const unsigned w = __ldg(reinterpret_cast<const unsigned *>(samples) + g);
                x = make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
            } else {
                x = __ldg(reinterpret_cast<const float2 *>(samples) + g);
            }
        }
        win4[i] = x;
    }
    for (int i = tid; i < L; i += kPc4Threads) { const float2 hm = __ldg(&P.matched[i]); hf[i] = make_float4(hm.x, hm.y, -hm.y, hm.x); }
    __syncthreads();
    const float lvl2 = P.pr_level * P.pr_level;
    for (int t = tid; 4 * t < NC; t += kPc4Threads) {
        float2 a[4];
        float e[4];
        corr4_lags(win4, plane, hf, L, t, a, e);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = 4 * t + q;
            float cval = 0.f;
            if (e[q] > 1.0f) {                                 // Frame.cpp:319
                cval = sqrtf(cnorm2(a[q]) / e[q]);
                if (cnorm2(a[q]) > lvl2 * e[q]) atomicMin(&first, i);   // Frame.cpp:364
            }
            if (cor_out != nullptr) cor_out[(size_t)c * NC + i] = cval;
        }
    }
    __syncthreads();
    if (first_idx != nullptr && tid == 0) first_idx[c] = first == 0x7fffffff ? -10 : st + first;
}

// ------------------------------------------------------------------------------------------------
// Modulation::mod / demod as stand-alone batched kernels (modulation.cpp:39-87)
// ------------------------------------------------------------------------------------------------
__global__ void mod_kernel(const float2 *__restrict__ table, int mod, const uint8_t *__restrict__ bytes, long long n_bytes,
                           float2 *__restrict__ points, long long n_points) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points) return;
    const long long bitpos = i * mod;
    const long long b0 = bitpos >> 3;
    const int off = (int)(bitpos & 7);
    unsigned w = 0;
    if (b0 < n_bytes) w = (unsigned)bytes[b0] << 8;
    if (b0 + 1 < n_bytes) w |= (unsigned)bytes[b0 + 1];
    const int sym = (int)((w >> (16 - mod - off)) & ((1u << mod) - 1u));
    points[i] = __ldg(&table[sym]);
}

// one thread demaps 8 points -> `mod` bytes; n_points is padded with zero symbols to a multiple of 8
// exactly like bit_stream_converter pads its last output word.
__global__ void demod_kernel(int mod, const float2 *__restrict__ points, long long n_points,
                             uint8_t *__restrict__ bytes, long long n_bytes, unsigned long long *__restrict__ ambiguous) {
    const long long grp = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int n_amb = 0;
    if (grp * 8 < n_points) {
        unsigned long long bits = 0;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const long long i = grp * 8 + e;
            int sym = 0;
            if (i < n_points) {
                bool amb;
                sym = demap_point(points[i], mod, amb);
                n_amb += amb ? 1 : 0;
            }
            bits = (bits << mod) | (unsigned long long)sym;
        }
        for (int bq = 0; bq < mod; bq++) {
            const long long o = grp * mod + bq;
            if (o < n_bytes) bytes[o] = (uint8_t)(bits >> (8 * (mod - 1 - bq)));
        }
    }
    if (ambiguous != nullptr && n_amb) atomicAdd(ambiguous, (unsigned long long)n_amb);
}

// first index i with rel[i] > level -> *pos = start + i*block, else -1 (Frame.hpp:190-196); *pos must
// be pre-set to a large value; a second launch of fix_not_found_kernel maps "large" to -1.
__global__ void first_above_kernel(const float *__restrict__ rel, long long n, float level, long long start,
                                   int block, unsigned long long *__restrict__ pos) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && rel[i] > level) {
        const unsigned long long v = (unsigned long long)(start + i * block);
        // 64-bit atomicMin via CAS (keeps the emulator simple and is contention-free in practice)
        unsigned long long old = *pos;
        while (v < old) {
#ifdef COFDM_EMU
            if (__atomic_compare_exchange_n(pos, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) break;
#else
            const unsigned long long prev = atomicCAS(pos, old, v);
            if (prev == old) break;
            old = prev;
#endif
        }
    }
}

// FRAME_FORM::form_int16_to_double analogue (Frame.hpp:472-481), fp32 on the device
__global__ void i16_to_cf32_kernel(const unsigned *__restrict__ in, float2 *__restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned w = __ldg(in + i);
    out[i] = make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
}

}  // namespace cofdmk

#include "tx512w.cuh"   // uses store_sample_pair
