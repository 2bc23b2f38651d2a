// kernels.cuh -- the sm_100a kernels of the C-OFDM hot path.
//
//   rx_fused512_kernel   aligned frame -> payload bytes in ONE pass over the samples:
//                        coarse CFO (pilot_freq_sinh), fine CFO (cp_freq_sinh), preamble phase lock
//                        (pr_phase_sinh), 9 FFT-512, pilot normalisation + segment correction
//                        (FFT_FORM::read), linear-phase channel fit (chan_char_lq), equalise, hard demap.
//   tx512_kernel         payload bytes -> QAM map -> pilot insertion -> 8 IFFT-512 -> CP -> frame
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

namespace cofdmk {

// ------------------------------------------------------------------------------------------------
// shared-memory plan of rx_fused512_kernel
// ------------------------------------------------------------------------------------------------
constexpr int kRxMaxSym = 16;
struct RxMisc {
    uint64_t mbar[kRxMaxSym];
    float2 cpcorr[kRxMaxSym];        // raw CP correlation per symbol
    float2 pilots[kRxMaxSym][8];     // un-normalised pilot bins per symbol
    float pabs[kRxMaxSym];           // sum |pilot| per symbol
    int amax[8];                     // arg-max per coarse-CFO window
    double a, b;                     // chan_char_lq line
    float2 rot_theta;                // exp(-j*theta) of pr_phase_sinh
    float theta;
};
constexpr int kRxScratch = 640;      // float2 per FFT-640 scratch buffer
COFDM_HD size_t rx_fused512_smem_bytes(int nsym) {
    return (size_t)nsym * kFft512Slots * sizeof(float2)   // X
           + 2 * kRxScratch * sizeof(float2)              // SA, SB
           + 256 * sizeof(float2)                         // conj(H)
           + (size_t)nsym * 256                           // demapped symbols
           + sizeof(RxMisc);
}

// Load one symbol's 640 samples into shared memory without TMA (int16 wire format, or cf32 whose
// frame stride is not 16-byte aligned).
template <int FMT>
COFDM_DEV void load_symbol_direct(float2 *dst, const void *src_frame, int sym, int lane) {
    if (FMT == kCI16) {
        const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const short2 *>(src_frame) + (size_t)sym * 640);
#pragma unroll
        for (int i = 0; i < 5; i++) {
            const int idx = lane + 32 * i;
            const uint4 raw = __ldg(src + idx);
            const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int e = 0; e < 4; e++)
                dst[4 * idx + e] = make_float2((float)(short)(w[e] & 0xffffu), (float)(short)(w[e] >> 16));
        }
    } else {
        const float4 *src = reinterpret_cast<const float4 *>(reinterpret_cast<const float2 *>(src_frame) + (size_t)sym * 640);
        float4 *d4 = reinterpret_cast<float4 *>(dst);
#pragma unroll
        for (int i = 0; i < 10; i++) d4[lane + 32 * i] = __ldg(src + lane + 32 * i);
    }
}

// FMT: sample format of `samples`; USE_TMA: stage cf32 frames with cp.async.bulk + mbarrier.
// MAXSYM bounds the symbols (= warps) per frame so the register budget can target 3 resident
// CTAs per SM for the shipped 9-symbol frame.
template <int FMT, bool USE_TMA, int MAXSYM>
__global__ void __launch_bounds__(32 * MAXSYM, MAXSYM <= 9 ? 3 : 1)
rx_fused512_kernel(const Params P, const void *__restrict__ samples, long long frame_stride /*samples*/,
                   int n_frames, uint8_t *__restrict__ out_bytes, unsigned long long *__restrict__ ambiguous,
                   const RxTaps taps) {
    COFDM_DYN_SMEM(smem_raw);
    const int nsym = P.n_sym_rx;              // 1 preamble + num_symb message symbols
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthr = blockDim.x;
    const int frame = blockIdx.x;
    if (frame >= n_frames) return;

    float2 *X = reinterpret_cast<float2 *>(smem_raw);
    float2 *SA = X + (size_t)nsym * kFft512Slots;
    float2 *SB = SA + kRxScratch;
    float2 *Hc = SB + kRxScratch;
    uint8_t *symbuf = reinterpret_cast<uint8_t *>(Hc + 256);
    RxMisc *M = reinterpret_cast<RxMisc *>(symbuf + (size_t)nsym * 256);
    float2 *Xw = X + (size_t)warp * kFft512Slots;     // this warp's symbol / FFT work region

    const size_t sample_bytes = (FMT == kCI16) ? 4 : 8;
    const char *frame_src = reinterpret_cast<const char *>(samples) + (size_t)frame * (size_t)frame_stride * sample_bytes;

    // ---- phase 0: stage the frame -------------------------------------------------------------
    if (USE_TMA) {
        if (tid == 0) {
            for (int s = 0; s < nsym; s++) mbar_init(&M->mbar[s], 1);
            mbar_fence_init();
            for (int s = 0; s < nsym; s++) {
                mbar_arrive_expect_tx(&M->mbar[s], 640 * 8);
                tma_load_1d(X + (size_t)s * kFft512Slots, frame_src + (size_t)s * 640 * 8, 640 * 8, &M->mbar[s]);
            }
        }
        __syncthreads();
        mbar_wait(&M->mbar[warp], 0);
    } else {
        load_symbol_direct<FMT>(Xw, frame_src, warp, lane);
        __syncwarp();
    }

    // ---- phase A: raw cyclic-prefix correlation of this symbol (Frame.hpp:251-253) --------------
    {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; i++) cmac_conj(acc, Xw[lane + 32 * i], Xw[lane + 32 * i + 512]);
        acc = warp_sum(acc);
        if (lane == 0) M->cpcorr[warp] = acc;
    }

    // ---- phase B: 640-point spectrum of the received preamble, CP included (Frame.hpp:286-309) ---
    if (USE_TMA) mbar_wait(&M->mbar[0], 0); else __syncthreads();
    stockham_pass<5, false>(X, SA, 640, 1, P.tw_pf, tid, nthr);
    __syncthreads();
    stockham_pass<8, false>(SA, SB, 640, 5, P.tw_pf, tid, nthr);
    __syncthreads();
    {   // last pass (radix 16, ns = 40): only |X|^2 is kept, in SA viewed as float[640]
        float *mag = reinterpret_cast<float *>(SA);
        for (int j = tid; j < 40; j += nthr) {
            float2 v[16];
#pragma unroll
            for (int q = 0; q < 16; q++) {
                v[q] = SB[j + 40 * q];
                if (q > 0) v[q] = cmul(v[q], __ldg(&P.tw_pf[(q * j) % 640]));
            }
            dft16<false>(v);
#pragma unroll
            for (int q = 0; q < 16; q++) mag[j + 40 * q] = cnorm2(v[q]);
        }
    }
    __syncthreads();

    // ---- phase C: arg-max of |spectrum| in the pilot windows (Frame.hpp:311-331) -----------------
    {
        const float *mag = reinterpret_cast<const float *>(SA);
        const int np = P.num_pilot_subc, half = P.pf_size / 2;
        for (int wi = warp; wi < np; wi += (nthr >> 5)) {
            const int win = wi < np / 2 ? wi : wi + 1;                 // window np/2 (DC) is skipped
            int lo = P.pf_border0 + win * P.pf_pilot_w, hi = lo + P.pf_pilot_w;
            if (win == 0 && lo < 0) lo = 0;
            float best = -1.0f;
            int besti = 0x7fffffff;
            for (int ks = lo + lane; ks < hi; ks += 32) {              // ks = fft-shifted index
                const int k = ks < half ? ks + half : ks - half;
                const float m = mag[k];
                if (m > best) { best = m; besti = ks; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {                         // first maximum wins ties
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
            }
            if (lane == 0) M->amax[wi] = besti;
        }
    }
    __syncthreads();

    // ---- phase D: per-symbol frequency = coarse shift + CP-correlation angle ---------------------
    // shift = pf_num/pf_den cycles/sample exactly (Frame.hpp:332-334); the CP correlation is taken
    // after freq_shift in the reference, which only rotates it by exp(-j*2*pi*shift*fft_size).
    int pf_num = 0;
    for (int i = 0; i < P.num_pilot_subc; i++) pf_num += M->amax[i];
    pf_num -= P.num_pilot_subc * (P.pf_size / 2);
    const double fc = (double)pf_num / (double)P.pf_den;
    float phi_l = 0.f;                                    // lane s holds phi_s
    if (lane < nsym) {
        const float2 c = cmul(M->cpcorr[lane], cis_neg_turns(fc * 512.0));
        phi_l = atan2f(c.y, c.x);                         // Frame.hpp:254 std::arg(phase)
    }
    const float phi_w = __shfl_sync(0xffffffffu, phi_l, warp);
    const float phi_0 = __shfl_sync(0xffffffffu, phi_l, 0);
    const double inv2pi = 0.15915494309189533577;
    const double nu = fc + (double)phi_w * inv2pi / 512.0;   // turns per sample inside this symbol
    // constant phase (turns) carried into symbol `warp` by freq_shift's global index and by
    // cp_freq_sinh's accumulated `shift` (Frame.hpp:248,261); only needed for symbol 1 and the taps
    double psi_turns = 0.0;
    {
        double acc = 0.0;
        for (int s = 0; s < nsym - 1; s++) {
            acc += (double)__shfl_sync(0xffffffffu, phi_l, s) * inv2pi * (640.0 / 512.0);
            if (s + 1 == warp) psi_turns = acc;
        }
        psi_turns += fc * 640.0 * (double)warp;
    }

    // ---- phase E: rotate while loading, FFT-512 of this symbol -----------------------------------
    float2 v[2][8];
    float2 Ph[2], Q[8];
    {
        float2 ql = make_float2(1.f, 0.f);
        if (lane < 8) ql = cis_neg_turns(nu * 64.0 * (double)lane);
#pragma unroll
        for (int r = 0; r < 8; r++) {
            Q[r].x = __shfl_sync(0xffffffffu, ql.x, r);
            Q[r].y = __shfl_sync(0xffffffffu, ql.y, r);
        }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int t = lane + 32 * h;
        Ph[h] = cis_neg_turns(nu * (double)(128 + t));
#pragma unroll
        for (int r = 0; r < 8; r++) v[h][r] = cmul(Xw[128 + t + 64 * r], cmul(Ph[h], Q[r]));
    }
    if (taps.synced != nullptr) {
        // debug tap: all rx_len samples after the three corrections need theta, which is only
        // known after warp 0 finishes; the host wrapper applies exp(-j*theta) from taps.scal.
        float2 *dst = taps.synced + (size_t)frame * P.rx_len + (size_t)warp * 640;
        const float2 cph = cis_neg_turns(psi_turns);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int t = lane + 32 * h;
#pragma unroll
            for (int r = 0; r < 8; r++) dst[128 + t + 64 * r] = cmul(v[h][r], cph);
            dst[t] = cmul(cmul(Xw[t], cmul(Ph[h], cconj(Q[2]))), cph);
            dst[t + 64] = cmul(cmul(Xw[t + 64], cmul(Ph[h], cconj(Q[1]))), cph);
        }
    }
    if (warp == 0) {
        // pr_phase_sinh (Frame.hpp:265-274): theta = arg sum_{i<640} conj(ref[i]) * y[i], y = rotated preamble
        float2 z = make_float2(0.f, 0.f);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int t = lane + 32 * h;
#pragma unroll
            for (int r = 0; r < 8; r++) cmac_conj(z, __ldg(&P.preamble_td[128 + t + 64 * r]), v[h][r]);
            const float2 y0 = cmul(Xw[t], cmul(Ph[h], cconj(Q[2])));          // j = t       : nu*(t)    = nu*(128+t) - nu*128
            const float2 y1 = cmul(Xw[t + 64], cmul(Ph[h], cconj(Q[1])));     // j = t + 64  : nu*(t+64) = nu*(128+t) - nu*64
            cmac_conj(z, __ldg(&P.preamble_td[t]), y0);
            cmac_conj(z, __ldg(&P.preamble_td[t + 64]), y1);
        }
        z = warp_sum(z);
        const float inv = rsqrtf(fmaxf(cnorm2(z), 1e-30f));
        if (lane == 0) {
            M->rot_theta = make_float2(z.x * inv, -z.y * inv);
            M->theta = atan2f(z.y, z.x);
        }
    }
    __syncwarp();                                          // every lane has read Xw before it is overwritten
    warp_fft512_head<false>(v, P.tw_p1, lane);
    warp_fft512_tail<false>(v, Xw, P.tw_p2, lane);

    // ---- phase F: pilots; channel line on the preamble (warp 0) ----------------------------------
    {
        float pa = 0.f;
        if (lane < 8) {
            const float2 pv = Xw[spec_slot(__ldg(&P.pilot_bin[lane]))];
            M->pilots[warp][lane] = pv;
            pa = sqrtf(cnorm2(pv));
        }
        pa = warp_sum(pa);
        if (lane == 0) M->pabs[warp] = pa;
    }
    if (warp == 0) {
        // chan_char_lq (Frame.hpp:389-434).  phase[i] = arg(pr[i]/mod_preamble[i]); the division by the
        // positive pilot-amplitude normaliser inside FFT_FORM::read cannot change an argument.
        const float2 rot = M->rot_theta;                   // written by lane 0 above, same warp
        float ph[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int i = 4 * lane + e;
            const float2 y = cmul(Xw[spec_slot(__ldg(&P.data_bin[i]))], rot);
            const float2 d = cmulc(y, __ldg(&P.mod_preamble[i]));
            ph[e] = atan2f(d.y, d.x);
        }
        // one-step unwrap (Frame.hpp:407-414) is a 3-state chain: state c in {-1,0,+1} = multiple of 2*pi
        // added to the previous element.  Each lane builds the transition map of its 4 elements for
        // every incoming state, the maps are composed across lanes by a warp scan, then replayed.
        const float PI_F = 3.14159265358979323846f, TWO_PI_F = 6.28318530717958647692f;
        float prev_raw = __shfl_up_sync(0xffffffffu, ph[3], 1);   // phase[4*lane-1] before unwrapping
        unsigned map = 0;                                          // 2 bits per incoming state (c+1)
#pragma unroll
        for (int cin = 0; cin < 3; cin++) {
            int c = cin - 1;
            float pv = prev_raw;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (lane == 0 && e == 0) { c = 0; pv = ph[0]; continue; }   // i = 0 is never adjusted
                const float dlt = ph[e] - (pv + (float)c * TWO_PI_F);
                c = dlt > PI_F ? -1 : (dlt < -PI_F ? 1 : 0);
                pv = ph[e];
            }
            map |= (unsigned)(c + 1) << (2 * cin);
        }
        // inclusive scan of map composition: after the scan, map_l(s) = state after lane l's elements
        // given state s before lane 0's.
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, map, o);
            if (lane >= o) {
                unsigned comp = 0;
#pragma unroll
                for (int cin = 0; cin < 3; cin++) {
                    const unsigned mid = (up >> (2 * cin)) & 3u;
                    comp |= ((map >> (2 * mid)) & 3u) << (2 * cin);
                }
                map = comp;
            }
        }
        unsigned before = __shfl_up_sync(0xffffffffu, map, 1);
        int c = lane == 0 ? 0 : (int)((before >> 2) & 3u) - 1;     // start state 0 (index 1)
        float pv = prev_raw;
        double sy = 0.0, sxy = 0.0;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int i = 4 * lane + e;
            float val = ph[e];
            if (!(lane == 0 && e == 0)) {
                const float dlt = ph[e] - (pv + (float)c * TWO_PI_F);
                c = dlt > PI_F ? -1 : (dlt < -PI_F ? 1 : 0);
                val = ph[e] + (float)c * TWO_PI_F;
            } else {
                c = 0;
            }
            pv = ph[e];
            sy += (double)val;
            sxy += (double)val * (double)i;
        }
        sy = warp_sum(sy);
        sxy = warp_sum(sxy);
        if (lane == 0) {
            const double n = 128.0;
            const double sx = n * (n - 1.0) / 2.0, sx2 = (n - 1.0) * n * (2.0 * n - 1.0) / 6.0;
            const double b = (sxy - sx * sy) / (sx2 - sx * sx);    // Frame.hpp:422 (sums, not means)
            M->b = b;
            M->a = sy - b * sx;                                    // Frame.hpp:423
        }
    }
    __syncthreads();

    // ---- phase G: H = exp(j(b*i'+a)), i' = i (i<128) or i-256 (Frame.hpp:425-430); 1/H = conj(H) ---
    const double la = M->a, lb = M->b;
    for (int i = tid; i < 256; i += nthr) {
        const int ip = i < 128 ? i : i - 256;
        const float2 hval = cis_turns((lb * (double)ip + la) * inv2pi);
        Hc[i] = cconj(hval);
        if (taps.chan != nullptr) taps.chan[(size_t)frame * 256 + i] = hval;
    }
    // pilot amplitude normaliser over all message symbols (Frame.cpp:76-80)
    float g = 0.f;
    for (int s = 1; s < nsym; s++) g += M->pabs[s];
    g /= (float)((nsym - 1) * 8) * P.pilot_ampl;
    const float inv_g = 1.0f / g;
    // constant rotation of message symbol 0: accumulated CFO phase + theta
    const float2 rot_theta = M->rot_theta;
    float2 wc = make_float2(0.f, 0.f);
    if (warp >= 1 && lane < 8) {
        // Frame.cpp:89-92 + rx.cpp:214-216:  out = (X/g) / ((P[s,p]/g)/(P[1,p]/g)) / H
        // the per-symbol constant rotation cancels between X[s,.] and P[s,p]; P[1,p] keeps its own.
        double psi1 = (double)phi_0 * inv2pi * (640.0 / 512.0) + fc * 640.0;
        const float2 rot1 = cmul(cis_neg_turns(psi1), rot_theta);
        const float2 p1 = cmul(M->pilots[1][lane], rot1);
        const float2 ps = M->pilots[warp][lane];
        const float den = 1.0f / (cnorm2(ps) * g);
        wc = cscale(cmulc(p1, ps), den);
    }
    if (taps.scal != nullptr && tid == 0) {
        float *sc = taps.scal + (size_t)frame * 8;
        sc[0] = (float)fc; sc[1] = (float)la; sc[2] = (float)lb; sc[3] = M->theta;
        sc[4] = g; sc[5] = (float)pf_num; sc[6] = 0.f; sc[7] = 0.f;
    }
    if (taps.grid != nullptr && warp >= 1) {
        // FFT_buf after FFT_FORM::read's normalisation (all 512 bins of this symbol, fully rotated)
        const float2 rs = cscale(cmul(cis_neg_turns(psi_turns), rot_theta), inv_g);
        float2 *dst = taps.grid + ((size_t)frame * (nsym - 1) + (warp - 1)) * 512;
        for (int k = lane; k < 512; k += 32) dst[k] = cmul(Xw[spec_slot(k)], rs);
    }
    __syncthreads();

    // ---- phase H: equalise + hard demap + pack (modulation.cpp:53-87) ----------------------------
    if (warp >= 1) {
        const int mod = P.mod_type;
        uint8_t *sb = symbuf + (size_t)warp * 256;
        int n_amb = 0;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int i = lane + 32 * e;
            float2 we;
            we.x = __shfl_sync(0xffffffffu, wc.x, e);
            we.y = __shfl_sync(0xffffffffu, wc.y, e);
            const float2 z = cmul(cmul(Xw[spec_slot(__ldg(&P.data_bin[i]))], we), Hc[i]);
            if (taps.constell != nullptr) taps.constell[((size_t)frame * (nsym - 1) + (warp - 1)) * 256 + i] = z;
            bool amb;
            sb[i] = (uint8_t)demap_point(z, mod, amb);
            n_amb += amb ? 1 : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_amb += __shfl_xor_sync(0xffffffffu, n_amb, o);
        if (ambiguous != nullptr && lane == 0 && n_amb) atomicAdd(ambiguous, (unsigned long long)n_amb);
        __syncwarp();
        // 8 consecutive symbols of `mod` bits = `mod` whole bytes, MSB first (modulation.cpp:90-125)
        unsigned long long bits = 0;
#pragma unroll
        for (int e = 0; e < 8; e++) bits = (bits << mod) | (unsigned long long)sb[8 * lane + e];
        uint8_t *dst = out_bytes + (size_t)frame * P.bytes_per_frame + (size_t)(warp - 1) * 32 * mod + (size_t)lane * mod;
        for (int bq = 0; bq < mod; bq++) dst[bq] = (uint8_t)(bits >> (8 * (mod - 1 - bq)));
    }
}

// ------------------------------------------------------------------------------------------------
// tx512_kernel: FRAME_FORM::write + get / get_int16 (Frame.cpp:185-198, 54-70, 244-256)
// one CTA per frame; warp 0 copies the frame-invariant sync tone + preamble, warps 1..num_symb
// build one OFDM symbol each: map bits, insert pilots, IFFT-512, /sqrt(512), prepend CP.
// ------------------------------------------------------------------------------------------------
COFDM_HD size_t tx512_smem_bytes(int num_symb, int bytes_per_frame) {
    return (size_t)num_symb * kFft512Slots * sizeof(float2) + (size_t)((bytes_per_frame + 15) & ~15);
}

template <int FMT>
COFDM_DEV void store_sample_pair(void *frame_out, int idx /*even sample index*/, float2 a, float2 b, float mult) {
    if (FMT == kCI16) {
        // Frame.cpp:252: int16(trunc(re*mult)), int16(trunc(im*mult))
        short2 sa = make_short2((short)__float2int_rz(a.x * mult), (short)__float2int_rz(a.y * mult));
        short2 sb = make_short2((short)__float2int_rz(b.x * mult), (short)__float2int_rz(b.y * mult));
        uint2 pk;
        pk.x = ((unsigned)(unsigned short)sa.x) | ((unsigned)(unsigned short)sa.y << 16);
        pk.y = ((unsigned)(unsigned short)sb.x) | ((unsigned)(unsigned short)sb.y << 16);
        reinterpret_cast<uint2 *>(frame_out)[idx >> 1] = pk;
    } else {
        reinterpret_cast<float4 *>(frame_out)[idx >> 1] = make_float4(a.x, a.y, b.x, b.y);
    }
}

template <int FMT>
__global__ void __launch_bounds__(32 * (kRxMaxSym + 1))
tx512_kernel(const Params P, const uint8_t *__restrict__ payload, int n_frames, void *__restrict__ frames) {
    COFDM_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.x;
    if (frame >= n_frames) return;
    const int ns = P.num_symb;
    float2 *W = reinterpret_cast<float2 *>(smem_raw);
    uint8_t *pl = reinterpret_cast<uint8_t *>(W + (size_t)ns * kFft512Slots);
    const size_t sample_bytes = (FMT == kCI16) ? 4 : 8;
    char *fout = reinterpret_cast<char *>(frames) + (size_t)frame * P.frame_len * sample_bytes;

    for (int i = tid; i < P.bytes_per_frame; i += blockDim.x) pl[i] = payload[(size_t)frame * P.bytes_per_frame + i];
    __syncthreads();

    if (warp == 0) {
        // T2SIN tone + preamble are constants of the configuration (Frame.cpp:228-229)
        const int n_const = P.t2sin_size + P.pf_size;
        for (int i = 2 * lane; i < n_const; i += 64) {
            const float2 a = i < P.t2sin_size ? __ldg(&P.t2_tone[i]) : __ldg(&P.preamble_td[i - P.t2sin_size]);
            const float2 b = i + 1 < P.t2sin_size ? __ldg(&P.t2_tone[i + 1]) : __ldg(&P.preamble_td[i + 1 - P.t2sin_size]);
            store_sample_pair<FMT>(fout, i, a, b, P.mult);
        }
        return;
    }
    const int s = warp - 1;
    float2 *Ww = W + (size_t)s * kFft512Slots;
    const int mod = P.mod_type;
    float2 v[2][8];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int t = lane + 32 * h;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int m = __ldg(&P.bin_map[t + 64 * r]);
            float2 val = make_float2(0.f, 0.f);                                     // Frame.cpp:55
            if (m == -2) val = make_float2(P.pilot_ampl, 0.f);                       // Frame.cpp:56-57
            else if (m >= 0) {                                                       // Frame.cpp:59-62 + modulation.cpp:39-50
                const int sym = extract_bits(pl, P.bytes_per_frame, (s * P.num_data_subc + m) * mod, mod);
                val = __ldg(&P.constell[sym]);
            }
            v[h][r] = val;
        }
    }
    warp_fft512_head<true>(v, P.tw_p1, lane);                                        // Frame.cpp:64 (backward, unnormalised)
    warp_fft512_tail<true>(v, Ww, P.tw_p2, lane);
    const float sc = 0.04419417382415922028f;                                        // 1/sqrt(512), Frame.cpp:66-68
    const int base = P.t2sin_size + P.pf_size + s * 640;
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int n = 2 * lane + 64 * it;
        const float2 a = cscale(Ww[spec_slot(n)], sc), b = cscale(Ww[spec_slot(n + 1)], sc);
        store_sample_pair<FMT>(fout, base + 128 + n, a, b, P.mult);                   // Frame.cpp:191-192
        if (n >= 384) store_sample_pair<FMT>(fout, base + n - 384, a, b, P.mult);     // Frame.cpp:196-197 cyclic prefix
    }
}

// ------------------------------------------------------------------------------------------------
// t2sin_metric_kernel: T2SIN_FORM::corr / find_t2sin block metric (Frame.hpp:112-143).
// One warp per non-overlapping 256-sample block: FFT-256 (8x8x4 Stockham in the warp's private
// shared memory), rel = sum(mask*|X|^2) / sum(|X|^2); blocks with zero or NaN energy report 0.
// ------------------------------------------------------------------------------------------------
constexpr int kT2WarpsPerCta = 8;
template <int FMT>
__global__ void __launch_bounds__(32 * kT2WarpsPerCta)
t2sin_metric_kernel(const Params P, const void *__restrict__ samples, long long start, long long n_blocks,
                    float *__restrict__ rel_out) {
    __shared__ float2 buf[kT2WarpsPerCta][2][256];
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
    __syncwarp();
    stockham_pass<8, false>(A, B, 256, 1, P.tw_t2, lane, 32);
    __syncwarp();
    stockham_pass<8, false>(B, A, 256, 8, P.tw_t2, lane, 32);
    __syncwarp();
    stockham_pass<4, false>(A, B, 256, 64, P.tw_t2, lane, 32);
    __syncwarp();
    float tot = 0.f, sine = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float e = cnorm2(B[lane + 32 * i]);
        tot += e;
        sine += __ldg(&P.t2_mask[lane + 32 * i]) * e;
    }
    tot = warp_sum(tot);
    sine = warp_sum(sine);
    if (lane == 0) {
        float rel = sine / tot;
        if (tot == 0.f || rel != rel) rel = 0.f;          // Frame.hpp:132-138 `continue`
        rel_out[blk] = rel;
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
