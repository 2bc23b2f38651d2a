// rx512n.cuh -- the receive chain for the fft-512 geometry, second formulation: ONE WARP PER OFDM SYMBOL.
//
// Reference chain (main.cpp:60-80 == rx.cpp:200-220): pilot_freq_sinh (Frame.hpp:285-337), freq_shift (:340-348),
// cp_freq_sinh (:238-263), pr_phase_sinh (:265-274), chan_char_lq (:389-434), message.fft (:276-282 + Frame.cpp:73-96),
// the equaliser loop (rx.cpp:214-216) and Modulation::demod (modulation.cpp:53-87).
//
// rx_demod512_kernel -- the message symbols of one frame per CTA, one warp each:
//   TMA bulk copy of the symbol's 640 samples -> CP correlation -> theta_s -> (the coarse shift kc of the acquire kernel
//   is already known, so the WHOLE-bin part m_s of the rotation is applied in the time domain too: after the transform
//   every sub-carrier sits in a fixed lane and register) -> rotation merged into the first pass and its twiddles ->
//   warp FFT-512 (fft512w.cuh) -> the 8 pilots and sum|pilot| to shared memory -> ONE block barrier ->
//   equalise + hard-demap the lane's data bins STRAIGHT FROM REGISTERS -> one byte per symbol to shared memory ->
//   MSB-first bit packing, one coalesced store per lane.
//   The spectrum never goes to shared memory; the only cross-warp traffic is 8 pilots + 1 float per symbol.
//
// Algebra: symbol s is rotated by exp(-j 2 pi beta_s j / 512), beta_s = theta_s + m_s, j = sample index in
// the symbol (CP included), theta_s = Arg(CP correlation) in turns, m_s = the integer that makes theta_s - 512 shift + m_s
// fall into (-0.5, 0.5] (Frame.hpp:254).  Per-symbol constant phases cancel between a data bin and its segment pilot
// (Frame.cpp:89-92) except for message symbol 0 (the reference of every segment): c_1 = exp(-j 2 pi 1.25 (theta_0 + m_0)) exp(-j theta).
// Equalised point of data index i (bin k, segment e):
//   z = X_s[k] * P_1[e] c_1 / (P_s[e] g) * exp(-j (b i' + a)),   i' = i (i < 128) or i - 256
// and i' = (k mod 64) - 1 + off(e, k div 64), so exp(-j b i') splits into a per-LANE factor exp(-j b ((k mod 64) - 1)) and a
// per-(segment, k div 64) factor that is folded into the segment coefficient: 12 coefficients per symbol, 2 phasors per lane.
#pragma once
#include "compat.cuh"
#include "params.h"
#include "fft512w.cuh"
#include "modem.cuh"
#include "rx512.cuh"     // FrameScal, sym_turns, staged_sample

namespace cofdmk {

// position of bin k after warp_fft512: lane, slot (0 = a, 1 = b), register k3
COFDM_HD constexpr int f512_lane(int k) { return (((k & 63) & ~1) >> 3) + 8 * ((((k & 63) & ~1) & 7) >> 1); }
COFDM_HD constexpr int f512_slot(int k) { return k & 1; }
COFDM_HD constexpr int f512_k3(int k) { return k >> 6; }

// The fft-512 / 256 data / 8 pilot sub-carrier map (Frame.cpp:31-44) is fixed by the geometry; build_tables() checks
// these lists against the tables it derives from the config.
#define COFDM_F512_PILOTS(X) X(0, 33) X(1, 66) X(2, 99) X(3, 132) X(4, 380) X(5, 413) X(6, 446) X(7, 479)
// data bins whose register (k3 = 2 or 5) holds almost no other used bin: handled apart, seven lanes in one go
#define COFDM_F512_STRAG(X) X(0, 128) X(1, 129) X(2, 130) X(3, 131) X(4, 381) X(5, 382) X(6, 383)
constexpr int kF512Strag = 7;
constexpr int kF512Combos = 12;

constexpr int kDemodTabOff = kFft512wBytes;    // phasor table of the rotation (21 float2), beside the exchange planes
constexpr int kDemodRegion = 5376;             // bytes of a warp's region: staging (5120) / exchanges (5152) + phasor table (168); 42 x 128
// after the transform the region holds: [0, 1024) registers k3 = 2, 5 of every lane as four planes (k3 = 2 slot a, b;
// k3 = 5 slot a, b); [1024, 1152) the 12 segment coefficients; [1280, 1552) one byte per demapped symbol (+ dummy slots
// for bins that carry no data)
constexpr int kDemodWtOff = 1024, kDemodSymOff = 1280;

struct alignas(16) DemodShared {
    uint64_t mbar[kRxMaxSym];
    float2 pil[kRxMaxSym][8];        // pilot bins per message symbol (index s - 1)
    float pabs[kRxMaxSym];           // sum |pilot| per message symbol, zero beyond the last one
    float4 lcl[32];                  // per pass-3 lane: exp(-j b (c0 - 1)), exp(-j b c0)
    float2 ftab[16];                 // per (segment, k3) combination: c_1 exp(-j (b off + a))
    float theta_t[kRxMaxSym + 1];    // Arg(C_s) in turns, by frame symbol index (taps)
    int mshift[kRxMaxSym + 1];       // m_s
};

COFDM_HD size_t rx_demod512_smem_bytes(int num_symb) { return (size_t)num_symb * kDemodRegion + sizeof(DemodShared); }

// 640 samples of one symbol -> the warp's region, by the warp itself (sources that are not 16-byte aligned)
template <int FMT>
COFDM_DEV void warp_stage_symbol(void *dst, const char *src, int lane) {
    if (FMT == kCI16) {
        const unsigned *s = reinterpret_cast<const unsigned *>(src);
        unsigned *d = reinterpret_cast<unsigned *>(dst);
        for (int i = lane; i < 640; i += 32) d[i] = __ldg(s + i);
    } else {
        const float2 *s = reinterpret_cast<const float2 *>(src);
        float2 *d = reinterpret_cast<float2 *>(dst);
        for (int i = lane; i < 640; i += 32) d[i] = __ldg(s + i);
    }
}

// two adjacent samples (index 2u, 2u + 1) of a symbol staged in shared memory, or (GLOBAL) straight from the capture
template <int FMT, bool GLOBAL = false>
COFDM_DEV void staged_pair(const void *region, int u, float2 &a, float2 &b) {
    if (FMT == kCI16) {
        const uint2 w = GLOBAL ? __ldg(reinterpret_cast<const uint2 *>(region) + u) : reinterpret_cast<const uint2 *>(region)[u];
        a = make_float2((float)(short)(w.x & 0xffffu), (float)(short)(w.x >> 16));
        b = make_float2((float)(short)(w.y & 0xffffu), (float)(short)(w.y >> 16));
    } else {
        const float4 q = GLOBAL ? __ldg(reinterpret_cast<const float4 *>(region) + u) : reinterpret_cast<const float4 *>(region)[u];
        a = make_float2(q.x, q.y);
        b = make_float2(q.z, q.w);
    }
}

// exp(-j 2 pi (theta + m) J / 512) for an integer sample index J: the whole-bin part is reduced exactly in integers
COFDM_DEV float2 rot_phasor(float theta, int m, int J) {
    return fast_cis_turns(-(theta * ((float)J * (1.0f / 512.0f)) + (float)((m * J) & 511) * (1.0f / 512.0f)));
}

// hard decision of one equalised point, natural layout (modulation.cpp:53-87): clamp to [-1,1], (v + 1) half + 0.5, truncate
// == truncate v half + (half + 0.5) with the level clamped to [0, 2 half]; the float -> unsigned conversion saturates at 0.
// MOD > 0: modulation order known at compile time; MOD == 0: taken from dk.
template <int MOD>
COFDM_DEV unsigned demap_n(float2 z, const DemapK &dk) {
    const int mod = MOD ? MOD : dk.mod;
    if (mod == 1) return z.x + z.y > 0.0f ? 1u : 0u;
    const float half = MOD ? 0.5f * (float)((1 << (MOD >> 1)) - 1) : dk.half;
    const unsigned lmax = MOD ? (unsigned)((1 << (MOD >> 1)) - 1) : (unsigned)dk.lmax;
    const float2 u = p_fma(z, make_float2(half, half), make_float2(half + 0.5f, half + 0.5f));
    const unsigned li = min(__float2uint_rz(u.x), lmax), lq = min(__float2uint_rz(u.y), lmax);
    return MOD ? lq * (unsigned)(1 << (MOD >> 1)) + li : (li | (lq << dk.qshift));
}

// MSB-first packing of 8 demapped symbols (one per byte of raw) into `mod` bytes (modulation.cpp:90-125)
template <int MOD>
COFDM_DEV void pack8(uint2 raw, uint8_t *dst, int mod_rt) {
    const int mod = MOD ? MOD : mod_rt;
    if (mod == 4) {
        // 16-QAM: wire byte k = (symbol 2k << 4) | symbol 2k+1
        const unsigned ux = ((raw.x << 4) & 0x00f000f0u) | ((raw.x >> 8) & 0x000f000fu);
        const unsigned uy = ((raw.y << 4) & 0x00f000f0u) | ((raw.y >> 8) & 0x000f000fu);
        const unsigned lo = (ux & 0xffu) | ((ux >> 8) & 0xff00u), hi = (uy & 0xffu) | ((uy >> 8) & 0xff00u);
        *reinterpret_cast<unsigned *>(dst) = lo | (hi << 16);
    } else if (mod == 2) {
        const unsigned b0 = ((raw.x & 3u) << 6) | ((raw.x >> 4) & 0x30u) | ((raw.x >> 14) & 0xcu) | (raw.x >> 24);
        const unsigned b1 = ((raw.y & 3u) << 6) | ((raw.y >> 4) & 0x30u) | ((raw.y >> 14) & 0xcu) | (raw.y >> 24);
        *reinterpret_cast<unsigned short *>(dst) = (unsigned short)(b0 | (b1 << 8));
    } else {
        unsigned long long bits = 0;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const unsigned sy = ((e < 4 ? raw.x : raw.y) >> (8 * (e & 3))) & 0xffu;
            bits = (bits << mod) | (unsigned long long)sy;
        }
        for (int bq = 0; bq < mod; bq++) dst[bq] = (uint8_t)(bits >> (8 * (mod - 1 - bq)));
    }
}

// MOD: modulation order the instance is specialised for (2, 4), or 0 = any (read from the configuration)
#ifndef COFDM_DEMOD_MINB
#define COFDM_DEMOD_MINB 4
#endif
template <int FMT, bool USE_TMA, bool TAPS, int MAXW, int MOD>
__global__ void __launch_bounds__(32 * MAXW, MAXW <= 8 ? COFDM_DEMOD_MINB : 1)
rx_demod512_kernel(const Params P, const void *__restrict__ samples, long long frame_stride /*samples*/, int n_frames,
                   uint8_t *__restrict__ out_bytes, unsigned long long *__restrict__ ambiguous, const RxTaps taps,
                   const int sync_less, const FrameScal *__restrict__ fscal) {
    COFDM_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform by construction: lets the compiler keep the warp's bases in uniform registers
    const int frame = blockIdx.x;
    if (frame >= n_frames) return;
    const int nw = P.num_symb;                     // one warp per message symbol
    const int s = warp + 1;                        // frame symbol index (0 = preamble)
    char *region = reinterpret_cast<char *>(smem_raw) + (size_t)warp * kDemodRegion;
    DemodShared *M = reinterpret_cast<DemodShared *>(smem_raw + (size_t)nw * kDemodRegion);
    const size_t sample_bytes = (FMT == kCI16) ? 4 : 8;
    const char *src = reinterpret_cast<const char *>(samples) + ((size_t)frame * (size_t)frame_stride + (size_t)s * 640) * sample_bytes;

    // ---- stage the symbol: one TMA bulk copy issued by the warp that consumes it ----
#ifdef COFDM_DEMOD_DIRECT
    if (USE_TMA) {
    } else
#endif
    if (USE_TMA) {
        if (lane == 0) {
            mbar_init(&M->mbar[warp], 1);
            mbar_fence_init();
            mbar_arrive_expect_tx(&M->mbar[warp], 640 * (unsigned)sample_bytes);
            tma_load_1d(region, src, 640 * (unsigned)sample_bytes, &M->mbar[warp]);
        }
    } else {
        warp_stage_symbol<FMT>(region, src, lane);
    }
    // ---- while the copy is in flight: the acquire kernel's scalars and the frame-wide tables ----
    // (FrameScal is 40 bytes: lanes 0..4 fetch 8 bytes each -- {kc, m0} {th0, theta} {rot_theta} {a} {b} -- and the fields
    //  travel by shuffle to the few lanes that need them; every lane needs kc only)
    uint2 fsw = make_uint2(0u, 0u);
    if (!sync_less && lane < 5) fsw = __ldg(reinterpret_cast<const uint2 *>(fscal + frame) + lane);
    if (sync_less && lane == 2) fsw.x = 0x3f800000u;        // sync-less: kc = m0 = 0, th0 = theta = 0, rot_theta = 1, a = b = 0
    const int kc = (int)__shfl_sync(0xffffffffu, fsw.x, 0);
    if (tid >= nw && tid < kRxMaxSym) M->pabs[tid] = 0.f;   // unused entries (the others are written by their warps); ordered by the block barrier
    if (warp == 0) {
        // per pass-3 lane: exp(-j b (c0 - 1)), exp(-j b c0)
        const double fb = __hiloint2double((int)__shfl_sync(0xffffffffu, fsw.y, 4), (int)__shfl_sync(0xffffffffu, fsw.x, 4));
        const float bt = (float)fb * 0.15915494309189533577f;           // channel-line slope in turns per data index
        const int c0 = fft512w_c0(lane);
        const float2 la = cis_neg_turns_f(bt * (float)(c0 - 1)), lb = cis_neg_turns_f(bt * (float)c0);
        M->lcl[lane] = make_float4(la.x, la.y, lb.x, lb.y);
    }
    float2 rot_theta = make_float2(1.f, 0.f);
    if (warp == nw - 1 || TAPS) {
        const int m0 = (int)__shfl_sync(0xffffffffu, fsw.y, 0);
        const float th0 = __uint_as_float(__shfl_sync(0xffffffffu, fsw.x, 1));
        rot_theta = make_float2(__uint_as_float(__shfl_sync(0xffffffffu, fsw.x, 2)), __uint_as_float(__shfl_sync(0xffffffffu, fsw.y, 2)));
        const double fa = __hiloint2double((int)__shfl_sync(0xffffffffu, fsw.y, 3), (int)__shfl_sync(0xffffffffu, fsw.x, 3));
        const double fb = __hiloint2double((int)__shfl_sync(0xffffffffu, fsw.y, 4), (int)__shfl_sync(0xffffffffu, fsw.x, 4));
        if (warp == nw - 1 && lane < kF512Combos) {
            // c_1 exp(-j (b off + a)): the constant phase of message symbol 0 (Psi_1 = 1.25 (theta_0 + m_0) mod 1), theta, the channel line
            float acc = th0 * (640.0f / 512.0f);
            acc -= rintf(acc);
            const float psi1 = acc + (float)((5 * m0) & 3) * 0.25f;
            const float2 c1 = nmul(cis_neg_turns_f(psi1), rot_theta);
            const float2 ee = cis_neg_turns_f((float)((fb * (double)P.combo_off[lane] + fa) * 0.15915494309189533577));
            M->ftab[lane] = nmul(c1, ee);
        }
        if (TAPS && tid == 0) { M->theta_t[0] = th0; M->mshift[0] = m0; }
    }
    __syncwarp();
#ifndef COFDM_DEMOD_DIRECT
    if (USE_TMA) mbar_wait(&M->mbar[warp], 0);
#endif

    // ---- the lane's 16 body samples (pass-1 layout: t = 2 lane, 2 lane + 1; index 128 + t + 64 r) and 4 CP samples ----
    float2 va[8], vb[8], cpa[2], cpb[2];
#ifdef COFDM_DEMOD_DIRECT
    if (USE_TMA) {
#pragma unroll
        for (int r = 0; r < 8; r++) staged_pair<FMT, true>(src, 64 + lane + 32 * r, va[r], vb[r]);
#pragma unroll
        for (int c = 0; c < 2; c++) staged_pair<FMT, true>(src, lane + 32 * c, cpa[c], cpb[c]);
    } else
#endif
    {
#pragma unroll
    for (int r = 0; r < 8; r++) staged_pair<FMT>(region, 64 + lane + 32 * r, va[r], vb[r]);
#pragma unroll
    for (int c = 0; c < 2; c++) staged_pair<FMT>(region, lane + 32 * c, cpa[c], cpb[c]);
    }
    __syncwarp();                                  // the region may now be reused (phasor table, exchanges)

    // ---- CP correlation (Frame.hpp:251-253): CP sample j pairs with body sample j + 512, i.e. r = 6, 7 ----
    float theta = 0.f;
    int m = 0;
    if (!sync_less) {
        float2 c = nmac_conj(nmac_conj(make_float2(0.f, 0.f), cpa[0], va[6]), cpb[0], vb[6]);
        c = nmac_conj(nmac_conj(c, cpa[1], va[7]), cpb[1], vb[7]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c = nadd(c, make_float2(__shfl_xor_sync(0xffffffffu, c.x, o), __shfl_xor_sync(0xffffffffu, c.y, o)));
        theta = fast_atan2_turns(c.y, c.x);
        // m_s and the reference's phi_s (Frame.hpp:254): phi = theta - 512 shift + m in (-0.5, 0.5]
        m = (int)ceilf(-(theta - (float)kc * P.pf_bins512) - 0.5f);
    }
    if (TAPS && lane == 0) { M->theta_t[s] = theta; M->mshift[s] = m; }

    // ---- rotation phasors, ONE sincos per lane: table entry `lane` = exp(-j 2 pi beta J / 512) with
    //      J = 64 lane (lane < 8: Q^r), 1 (lane 8: D), 16 (lane - 9) (9..12: U^u), 128 + 2 (lane - 13) (13..20: V^v);
    //      P(t = 2 lane) = exp(-j 2 pi beta (128 + 2 lane) / 512) = V^(lane & 7) U^(lane >> 3) ----
    float2 *qt = reinterpret_cast<float2 *>(region + kDemodTabOff);
    {
        const int J = lane < 8 ? 64 * lane : (lane == 8 ? 1 : (lane < 13 ? 16 * (lane - 9) : 128 + 2 * (lane - 13)));
        const float2 ph = rot_phasor(theta, m, J);
        if (lane < 21) qt[lane] = ph;
    }
    __syncwarp();
    const float2 pa = nmul(qt[13 + (lane & 7)], qt[9 + (lane >> 3)]);
    const float2 pb = nmul(pa, qt[8]);
    if (TAPS && taps.synced != nullptr) {
        // debug tap, completed by rx_synced_fixup2_kernel (per-symbol constant phase and theta)
        float2 *d = taps.synced + (size_t)frame * P.rx_len + (size_t)s * 640;
        const int t = 2 * lane;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            d[128 + t + 64 * r] = nmul(nmul(va[r], qt[r]), pa);
            d[129 + t + 64 * r] = nmul(nmul(vb[r], qt[r]), pb);
        }
        // CP samples j = t + 64 c: exp(-j 2 pi beta j / 512) = P(t) conj(Q^(2 - c))
        d[t] = nmul(nmulc(cpa[0], qt[2]), pa);      d[t + 1] = nmul(nmulc(cpb[0], qt[2]), pb);
        d[t + 64] = nmul(nmulc(cpa[1], qt[1]), pa); d[t + 65] = nmul(nmulc(cpb[1], qt[1]), pb);
    }
    {
        const float4 *q4 = reinterpret_cast<const float4 *>(qt);
#pragma unroll
        for (int rr = 0; rr < 4; rr++) {
            const float4 q = q4[rr];
            if (rr > 0) { va[2 * rr] = nmul(va[2 * rr], make_float2(q.x, q.y)); vb[2 * rr] = nmul(vb[2 * rr], make_float2(q.x, q.y)); }
            va[2 * rr + 1] = nmul(va[2 * rr + 1], make_float2(q.z, q.w));
            vb[2 * rr + 1] = nmul(vb[2 * rr + 1], make_float2(q.z, q.w));
        }
    }
    warp_fft512(va, vb, pa, pb, reinterpret_cast<float2 *>(region), P.tw_fft, lane);
    // now va[k3] = X[c0 + 64 k3], vb[k3] = X[c0 + 1 + 64 k3], c0 = 2 (lane >> 3) + 8 (lane & 7); the region is free again

    // ---- pilots and sum |pilot| (Frame.cpp:76-80) to the CTA's shared memory; registers k3 = 2, 5 (the straggler data bins
    //      128..131 and 381..383 live there) to the warp's region ----
    float2 *scratch = reinterpret_cast<float2 *>(region);
    {
        float2 *pil = M->pil[warp];
#define COFDM_X(p, bin) if (lane == f512_lane(bin)) pil[p] = (f512_slot(bin) ? vb : va)[f512_k3(bin)];
        COFDM_F512_PILOTS(COFDM_X)
#undef COFDM_X
        scratch[lane] = va[2]; scratch[32 + lane] = vb[2]; scratch[64 + lane] = va[5]; scratch[96 + lane] = vb[5];
        __syncwarp();
        float pm = 0.f;
        if (lane < 8) pm = sqrtf(cnorm2(pil[lane]));
        pm += __shfl_xor_sync(0xffffffffu, pm, 4);
        pm += __shfl_xor_sync(0xffffffffu, pm, 2);
        pm += __shfl_xor_sync(0xffffffffu, pm, 1);
        if (lane == 0) M->pabs[warp] = pm;
    }
    __syncthreads();                               // pilots of every symbol, pabs, lcl, ftab are ready

    float g;                                       // pilot amplitude normaliser over all message symbols (Frame.cpp:76-80)
    {
        float pv = M->pabs[lane & (kRxMaxSym - 1)];
#pragma unroll
        for (int o = kRxMaxSym / 2; o > 0; o >>= 1) pv += __shfl_xor_sync(0xffffffffu, pv, o);
        g = pv * P.inv_pilot_norm;
    }
    const float4 lcv = M->lcl[lane];
    const float2 lca = make_float2(lcv.x, lcv.y), lcb = make_float2(lcv.z, lcv.w);

    // ---- the 12 segment coefficients of this symbol (Frame.cpp:89-92 + rx.cpp:214-216):
    //      W[q] = P_1[e] conj(P_s[e]) / (|P_s[e]|^2 g) * ftab[q],  e = segment of combination q ----
    char *wt = region + kDemodWtOff;
    if (lane < kF512Combos) {
        const int e = (int)((P.combo_seg_packed >> (4 * lane)) & 7ull);
        const float2 p1 = M->pil[0][e], ps = M->pil[warp][e];
        const float2 w = nscale(nmulc(p1, ps), __fdividef(1.0f, cnorm2(ps) * g));
        reinterpret_cast<float2 *>(wt)[lane] = nmul(w, M->ftab[lane]);
    }
    __syncwarp();

    if (TAPS) {
        if (taps.scal != nullptr && lane == 0) {
            float *sc = taps.scal + (size_t)frame * 48;
            if (warp == 0) {
                sc[4] = g;
                if (sync_less) { sc[0] = 0.f; sc[1] = 0.f; sc[2] = 0.f; sc[3] = 0.f; sc[5] = 0.f; sc[6] = 0.f; sc[7] = 0.f; sc[16] = 0.f; sc[32] = 0.f; }
            }
            sc[16 + s] = (float)m;
            sc[32 + s] = theta;
        }
        if (taps.grid != nullptr) {
            // FFT_buf after FFT_FORM::read's normalisation: every bin, with the symbol's constant phase and theta
            const float psi = sym_turns(M->theta_t, M->mshift, s);
            const float2 rs = nscale(nmul(cis_neg_turns_f(psi), rot_theta), 1.0f / g);
            float2 *dst = taps.grid + ((size_t)frame * nw + (s - 1)) * 512;
            const int c0 = fft512w_c0(lane);
#pragma unroll
            for (int k3 = 0; k3 < 8; k3++) { dst[c0 + 64 * k3] = nmul(va[k3], rs); dst[c0 + 1 + 64 * k3] = nmul(vb[k3], rs); }
        }
    }

    // ---- equalise + hard demap (modulation.cpp:53-87) straight from the registers.  Per slot 16 descriptor bits:
    //      [15:7] data index (>= 256: a dummy slot, the bin carries no data), [6:3] combination.  No branches. ----
    const DemapK dk = make_demapk(P.mod_type);
    uint8_t *sb = reinterpret_cast<uint8_t *>(region + kDemodSymOff);
    const uint4 desc = __ldg(&P.lane_desc[lane]);
    float2 *ctap = (TAPS && taps.constell != nullptr) ? taps.constell + ((size_t)frame * nw + (s - 1)) * 256 : nullptr;
    const float2 xa0 = nmul(va[0], lca), xa1 = nmul(va[1], lca), xa6 = nmul(va[6], lca), xa7 = nmul(va[7], lca);
    const float2 xb0 = nmul(vb[0], lcb), xb1 = nmul(vb[1], lcb), xb6 = nmul(vb[6], lcb), xb7 = nmul(vb[7], lcb);
    float2 xs = make_float2(0.f, 0.f);             // the straggler this lane equalises (lanes 0..6)
    unsigned ds = 0x8000u;                         // dummy slot 256
    if (lane < kF512Strag) {
        const unsigned d32 = P.strag_desc[lane];   // descriptor | scratch slot << 16 | (origin lane * 2 + slot) << 24
        ds = d32 & 0xffffu;
        xs = nmul(scratch[(d32 >> 16) & 0xffu], reinterpret_cast<const float2 *>(M->lcl)[d32 >> 24]);
    }
#define COFDM_EQ(X, IDX, WOFF)                                                                   \
    do {                                                                                         \
        const float2 z_ = nmul((X), *reinterpret_cast<const float2 *>(wt + (WOFF)));             \
        if (TAPS && ctap != nullptr && (IDX) < 256u) ctap[(IDX)] = z_;                           \
        sb[(IDX)] = (uint8_t)demap_n<MOD>(z_, dk);                                               \
    } while (0)
#define COFDM_EQ_ALL(F)                                                                          \
    F(xa0, (desc.x >> 7) & 0x1ffu, desc.x & 0x78u); F(xb0, desc.x >> 23, (desc.x >> 16) & 0x78u); \
    F(xa1, (desc.y >> 7) & 0x1ffu, desc.y & 0x78u); F(xb1, desc.y >> 23, (desc.y >> 16) & 0x78u); \
    F(xa6, (desc.z >> 7) & 0x1ffu, desc.z & 0x78u); F(xb6, desc.z >> 23, (desc.z >> 16) & 0x78u); \
    F(xa7, (desc.w >> 7) & 0x1ffu, desc.w & 0x78u); F(xb7, desc.w >> 23, (desc.w >> 16) & 0x78u); \
    F(xs, (ds >> 7) & 0x1ffu, ds & 0x78u)
    COFDM_EQ_ALL(COFDM_EQ);
#undef COFDM_EQ
    if (ambiguous != nullptr) {
        // optional count of boundary-ambiguous decisions (margin kAmbigMargin): the points are recomputed, off the fast path
        int n_amb = 0;
#define COFDM_AMB(X, IDX, WOFF) \
        if ((IDX) < 256u) n_amb += demap_ambiguous(nmul((X), *reinterpret_cast<const float2 *>(wt + (WOFF))), dk) ? 1 : 0
        COFDM_EQ_ALL(COFDM_AMB);
#undef COFDM_AMB
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_amb += __shfl_xor_sync(0xffffffffu, n_amb, o);
        if (lane == 0 && n_amb) atomicAdd(ambiguous, (unsigned long long)n_amb);
    }
#undef COFDM_EQ_ALL
    __syncwarp();
    // ---- pack: 8 consecutive symbols of `mod` bits = `mod` whole bytes, MSB first (modulation.cpp:90-125) ----
    {
        const int mod = MOD ? MOD : P.mod_type;
        const uint2 raw = *reinterpret_cast<const uint2 *>(sb + 8 * lane);
        pack8<MOD>(raw, out_bytes + (size_t)frame * P.bytes_per_frame + (size_t)(s - 1) * 32 * mod + (size_t)lane * mod, mod);
    }
}

// ================================================================================================================
// rx_acquire512w_kernel -- the preamble of one frame per WARP (four frames per CTA, nothing shared between them):
//   coarse CFO   pilot_freq_sinh (Frame.hpp:285-337): 640-point spectrum (10 x 8 x 8 Stockham, natural layout), |X|^2,
//                arg-max in the pilot windows -> kc
//   fine CFO     cp_freq_sinh (:238-263) on the preamble: CP correlation -> theta_0, m_0; rotation, warp FFT-512
//   phase lock   pr_phase_sinh (:265-274): theta = arg sum conj(ref) y, body part by Parseval on the used bins
//   channel fit  chan_char_lq (:389-434): 128 phases, the reference's one-step unwrap, the (bug-compatible) line
// and hands 40 bytes of scalars (FrameScal) to the demod kernel.  The lane's 20 raw samples x[2 lane (+1) + 64 q] are the
// inputs of BOTH transforms (radix-10 first pass of the 640-point one, CP + radix-8 first pass of the 512-point one):
// they are read from the staged copy once and stay in registers.
// Shared memory per warp: S (5120 B: staged samples -> second coarse plane -> phasor table, phases) and A (5376 B:
// first coarse plane, padded -> |X|^2 -> FFT-512 exchanges).
// ================================================================================================================
constexpr int kAcqwWarps = 4;
constexpr int kAcqwS = 5120, kAcqwA = 5376;
constexpr int kAcqwRegion = kAcqwS + kAcqwA;
COFDM_HD constexpr size_t rx_acquire512w_smem_bytes() { return (size_t)kAcqwWarps * kAcqwRegion + kAcqwWarps * sizeof(uint64_t); }

template <int FMT, bool USE_TMA, bool TAPS>
__global__ void __launch_bounds__(32 * kAcqwWarps, 5)
rx_acquire512w_kernel(const Params P, const void *__restrict__ samples, long long frame_stride /*samples*/, int n_frames,
                      const RxTaps taps, FrameScal *__restrict__ fscal, const int sync_less) {
    COFDM_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int frame = blockIdx.x * kAcqwWarps + warp;
    if (frame >= n_frames) return;                     // whole warp; warps never meet at a block barrier
    char *S = reinterpret_cast<char *>(smem_raw) + (size_t)warp * kAcqwRegion;
    float2 *A = reinterpret_cast<float2 *>(S + kAcqwS);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kAcqwWarps * kAcqwRegion) + warp;
    const size_t sample_bytes = (FMT == kCI16) ? 4 : 8;
    const char *src = reinterpret_cast<const char *>(samples) + (size_t)frame * (size_t)frame_stride * sample_bytes;
    if (USE_TMA) {
        if (lane == 0) {
            mbar_init(mbar, 1);
            mbar_fence_init();
            mbar_arrive_expect_tx(mbar, 640 * (unsigned)sample_bytes);
            tma_load_1d(S, src, 640 * (unsigned)sample_bytes, mbar);
        }
    } else {
        warp_stage_symbol<FMT>(S, src, lane);
    }
    __syncwarp();
    if (USE_TMA) mbar_wait(mbar, 0);
    // ---- the lane's 20 raw samples: ra[q] = x[2 lane + 64 q], rb[q] = x[2 lane + 1 + 64 q] ----
    float2 ra[10], rb[10];
#pragma unroll
    for (int q = 0; q < 10; q++) staged_pair<FMT>(S, lane + 32 * q, ra[q], rb[q]);
    __syncwarp();                                      // S may be overwritten from here on

    int kc = 0;
    if (!sync_less) {
        // ================= coarse CFO: 640-point spectrum of the received preamble, CP included =================
        // pass 1: radix 10, butterflies j = 2 lane, 2 lane + 1; output index I = 10 j + q lives at slot I + I / 20
        {
            float2 v[10];
#pragma unroll
            for (int q = 0; q < 10; q++) v[q] = ra[q];
            ndft10(v);
#pragma unroll
            for (int q = 0; q < 10; q++) A[21 * lane + q] = v[q];
#pragma unroll
            for (int q = 0; q < 10; q++) v[q] = rb[q];
            ndft10(v);
#pragma unroll
            for (int q = 0; q < 10; q++) A[21 * lane + 10 + q] = v[q];
        }
        __syncwarp();
        float2 *B = reinterpret_cast<float2 *>(S);
        // pass 2: radix 8, ns = 10: butterfly j (80 of them), k = j mod 10; inputs I = j + 80 q at slot I + I / 20 = (j + j / 20) + 84 q
#pragma unroll 1
        for (int j = lane; j < 80; j += 32) {
            const int g = j / 10, k = j - 10 * g;
            float2 v[8], w[8];
            const float2 *in = A + j + j / 20;
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = in[84 * q];
            npowers7(__ldg(&P.tw_pf[8 * k]), w);                          // W640^{8 k q}
#pragma unroll
            for (int q = 1; q < 8; q++) v[q] = nmul(v[q], w[q]);
            ndft8(v);
            float2 *out = B + 80 * g + k;
#pragma unroll
            for (int q = 0; q < 8; q++) out[10 * q] = v[q];
        }
        __syncwarp();
        // pass 3: radix 8, ns = 80: only |X|^2 is kept, as float[640] in A
        float *mag = reinterpret_cast<float *>(A);
#pragma unroll 1
        for (int j = lane; j < 80; j += 32) {
            float2 v[8], w[8];
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = B[j + 80 * q];
            npowers7(__ldg(&P.tw_pf[j]), w);                              // W640^{j q}
#pragma unroll
            for (int q = 1; q < 8; q++) v[q] = nmul(v[q], w[q]);
            ndft8(v);
#pragma unroll
            for (int q = 0; q < 8; q++) { const float2 sq = p_mul(v[q], v[q]); mag[j + 80 * q] = sq.x + sq.y; }
        }
        __syncwarp();
        // arg-max of |spectrum| in the pilot windows, first maximum wins (Frame.hpp:311-331); warp arg-max by two hardware
        // reductions: the magnitudes are non-negative floats, so their bit patterns order like unsigned integers
        const int np = P.num_pilot_subc, half = P.pf_size / 2;
        int ksum = 0;
        for (int wi = 0; wi < np; wi++) {
            const int win = wi < np / 2 ? wi : wi + 1;                     // window np/2 (DC) is skipped
            int lo = P.pf_border0 + win * P.pf_pilot_w;
            const int hi = lo + P.pf_pilot_w;
            if (win == 0 && lo < 0) lo = 0;
            float best = -1.0f;
            int bi = 0x7fffffff;
            for (int ks = lo + lane; ks < hi; ks += 32) {                  // ks = fft-shifted index
                const float mv = mag[ks < half ? ks + half : ks - half];
                if (mv > best) { best = mv; bi = ks; }
            }
            const unsigned bb = best < 0.f ? 0u : __float_as_uint(best);
            const unsigned mx = __reduce_max_sync(0xffffffffu, bb);
            ksum += __reduce_min_sync(0xffffffffu, bb == mx ? bi : 0x7fffffff);
        }
        kc = ksum - np * half;                                             // shift = kc / pf_den (Frame.hpp:332-334)
        __syncwarp();                                                      // the planes are free again
    }

    // ================= fine CFO of the preamble (cp_freq_sinh): CP sample j pairs with body sample j + 512 (q = 8, 9) =================
    float theta0 = 0.f;
    int m0 = 0;
    if (!sync_less) {
        float2 c = nmac_conj(nmac_conj(make_float2(0.f, 0.f), ra[0], ra[8]), rb[0], rb[8]);
        c = nmac_conj(nmac_conj(c, ra[1], ra[9]), rb[1], rb[9]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c = nadd(c, make_float2(__shfl_xor_sync(0xffffffffu, c.x, o), __shfl_xor_sync(0xffffffffu, c.y, o)));
        theta0 = fast_atan2_turns(c.y, c.x);
        m0 = (int)ceilf(-(theta0 - (float)kc * P.pf_bins512) - 0.5f);
    }
    // rotation phasors (see rx_demod512_kernel): table entry `lane` = exp(-j 2 pi beta J / 512)
    float2 *qt = reinterpret_cast<float2 *>(S);
    {
        const int J = lane < 8 ? 64 * lane : (lane == 8 ? 1 : (lane < 13 ? 16 * (lane - 9) : 128 + 2 * (lane - 13)));
        const float2 ph = rot_phasor(theta0, m0, J);
        if (lane < 21) qt[lane] = ph;
    }
    __syncwarp();
    const float2 pa = nmul(qt[13 + (lane & 7)], qt[9 + (lane >> 3)]);
    const float2 pb = nmul(pa, qt[8]);
    // CP samples j = t + 64 c rotated: exp(-j 2 pi beta j / 512) = P(t) conj(Q^(2 - c))
    const float2 y0a = nmul(nmulc(ra[0], qt[2]), pa), y0b = nmul(nmulc(rb[0], qt[2]), pb);
    const float2 y1a = nmul(nmulc(ra[1], qt[1]), pa), y1b = nmul(nmulc(rb[1], qt[1]), pb);
    float2 va[8], vb[8];
    va[0] = ra[2]; vb[0] = rb[2];
#pragma unroll
    for (int r = 1; r < 8; r++) { va[r] = nmul(ra[r + 2], qt[r]); vb[r] = nmul(rb[r + 2], qt[r]); }
    if (TAPS && taps.synced != nullptr) {
        // debug tap, completed by rx_synced_fixup2_kernel (theta; the preamble has no other constant phase)
        float2 *d = taps.synced + (size_t)frame * P.rx_len;
        const int t = 2 * lane;
#pragma unroll
        for (int r = 0; r < 8; r++) { d[128 + t + 64 * r] = nmul(va[r], pa); d[129 + t + 64 * r] = nmul(vb[r], pb); }
        d[t] = y0a; d[t + 1] = y0b; d[t + 64] = y1a; d[t + 65] = y1b;
    }
    __syncwarp();                                                          // everybody has read the phasor table: A / S are free
    warp_fft512(va, vb, pa, pb, A, P.tw_fft, lane);
    // now va[k3] = Y[c0 + 64 k3], vb[k3] = Y[c0 + 1 + 64 k3], the true spectrum of the preamble (symbol 0 has no constant phase)

    const int c0 = fft512w_c0(lane);
    if (sync_less) {
        // PREAMBLE_FORM::chan_char (Frame.hpp:375-385) on the preamble as it stands: pr = preamble.fft() (own pilot
        // normalisation, Frame.cpp:76-84; coef == 1), chan_est[i] = pr[i] / mod_preamble[i]
        if (TAPS && taps.chan != nullptr) {
            float pm = 0.f;
#define COFDM_X(p, bin) if (lane == f512_lane(bin)) pm = sqrtf(cnorm2((f512_slot(bin) ? vb : va)[f512_k3(bin)]));
            COFDM_F512_PILOTS(COFDM_X)
#undef COFDM_X
            pm = warp_sum(pm);
            const float igp = (8.0f * P.pilot_ampl) / pm;
#pragma unroll
            for (int k3 = 0; k3 < 8; k3++) {
                if (k3 == 3 || k3 == 4) continue;
#pragma unroll
                for (int sl = 0; sl < 2; sl++) {
                    const int k = c0 + sl + 64 * k3, i = __ldg(&P.bin_map[k]);
                    if (i >= 0) {
                        const float2 mp = __ldg(&P.mod_preamble[i]);
                        const float2 y = cscale(sl ? vb[k3] : va[k3], igp);
                        taps.chan[(size_t)frame * 256 + i] = cscale(cmulc(y, mp), 1.0f / cnorm2(mp));
                    }
                }
            }
        }
        return;
    }

    // ================= pr_phase_sinh: z = sum_{i<640} conj(ref[i]) y[i]; body by Parseval: (1/sqrt 512) sum_k conj(G[k]) Y[k],
    //                   G = tx grid of the preamble (P.grid_conj = conj(G) / sqrt 512, zero on unused bins) =================
    float2 prod[4];                                    // Y conj(G) of the slots k3 = 0, 1 (x slot a, b): the first 128 data sub-carriers live there
    float2 z;
    {
        const float4 *g4 = reinterpret_cast<const float4 *>(P.grid_conj) + (c0 >> 1);
        const float4 g0 = __ldg(g4), g1 = __ldg(g4 + 32), g2 = __ldg(g4 + 64), g5 = __ldg(g4 + 160), g6 = __ldg(g4 + 192), g7 = __ldg(g4 + 224);
        prod[0] = nmul(va[0], make_float2(g0.x, g0.y)); prod[1] = nmul(vb[0], make_float2(g0.z, g0.w));
        prod[2] = nmul(va[1], make_float2(g1.x, g1.y)); prod[3] = nmul(vb[1], make_float2(g1.z, g1.w));
        const float2 s2a = nmul(va[2], make_float2(g2.x, g2.y)), s2b = nmul(vb[2], make_float2(g2.z, g2.w));
        z = nadd(nadd(prod[0], prod[1]), nadd(prod[2], prod[3]));
        z = nadd(z, nadd(s2a, s2b));
        z = nmac(nmac(z, va[5], make_float2(g5.x, g5.y)), vb[5], make_float2(g5.z, g5.w));
        z = nmac(nmac(z, va[6], make_float2(g6.x, g6.y)), vb[6], make_float2(g6.z, g6.w));
        z = nmac(nmac(z, va[7], make_float2(g7.x, g7.y)), vb[7], make_float2(g7.z, g7.w));
        // CP part: conj(ref[j]) y[j], j = 2 lane (+1), 64 + 2 lane (+1)
        const float4 *r4 = reinterpret_cast<const float4 *>(P.preamble_td) + lane;
        const float4 r0 = __ldg(r4), r1 = __ldg(r4 + 32);
        z = nmac_conj(nmac_conj(z, make_float2(r0.x, r0.y), y0a), make_float2(r0.z, r0.w), y0b);
        z = nmac_conj(nmac_conj(z, make_float2(r1.x, r1.y), y1a), make_float2(r1.z, r1.w), y1b);
        // the data bins 128..131 (k3 = 2 of lanes 0 and 8) belong to the first 128 sub-carriers too: their products travel
        // through shared memory to the four lanes whose slot holds no data (bin 0 and the pilots 33, 66, 99)
        float2 *sx = qt + 24;
        if (lane == f512_lane(128)) { sx[0] = s2a; sx[1] = s2b; }
        if (lane == f512_lane(130)) { sx[2] = s2a; sx[3] = s2b; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) z = nadd(z, make_float2(__shfl_xor_sync(0xffffffffu, z.x, o), __shfl_xor_sync(0xffffffffu, z.y, o)));
    const float inv = rsqrtf(fmaxf(cnorm2(z), 1e-30f));
    const float2 rot = make_float2(z.x * inv, -z.y * inv);               // exp(-j theta)
    const float theta = TAPS ? atan2f(z.y, z.x) : 0.f;

    // ================= chan_char_lq: phase[i] = arg(pr[i] / mod_preamble[i]), i < 128 (Frame.hpp:403-405) =================
    const float TWO_PI_F = 6.28318530717958647692f, PI_F = 3.14159265358979323846f;
    float *phs = reinterpret_cast<float *>(qt + 32);                      // 128 phases
    {
        const uint2 ad = __ldg(&P.acq_desc[lane]);   // 4 x 16 bits (k3 = 0 a, b; k3 = 1 a, b): [7:0] phase index, [15] straggler, [9:8] which
        const float2 *sx = qt + 24;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned d = ((u < 2 ? ad.x : ad.y) >> (16 * (u & 1))) & 0xffffu;
            float2 pr = prod[u];
            if (d & 0x8000u) pr = sx[(d >> 8) & 3u];
            const float2 dr = nmul(pr, rot);
            phs[d & 0xffu] = fast_atan2_turns(dr.y, dr.x) * TWO_PI_F;
        }
    }
    __syncwarp();
    // one-step unwrap (Frame.hpp:407-414) and the sums of Frame.hpp:416-421: lane l owns phases 4l .. 4l + 3
    float tsy, tsxy;
    {
        const float4 p4v = reinterpret_cast<const float4 *>(phs)[lane];
        float p4[4] = {p4v.x, p4v.y, p4v.z, p4v.w};
        const float prev_raw = __shfl_up_sync(0xffffffffu, p4[3], 1);
        bool jump = (lane > 0 && fabsf(p4[0] - prev_raw) > PI_F) || fabsf(p4[1] - p4[0]) > PI_F || fabsf(p4[2] - p4[1]) > PI_F || fabsf(p4[3] - p4[2]) > PI_F;
        const bool any = __ballot_sync(0xffffffffu, jump) != 0u;
        float ssy = (p4[0] + p4[1]) + (p4[2] + p4[3]);
        float ssxy = fmaf(p4[3], 3.0f, fmaf(p4[2], 2.0f, p4[1])) + (float)(4 * lane) * ssy;
        if (any) {
            // slow path: the adjustment is a 3-state chain (state = multiple of 2 pi carried by the previous element); each lane
            // builds the transition map of its 4 elements for every incoming state, the maps are composed across lanes by a
            // warp scan, then replayed
            unsigned map = 0;
#pragma unroll
            for (int cin = 0; cin < 3; cin++) {
                int cc = cin - 1;
                float pv = prev_raw;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    if (lane == 0 && e == 0) { cc = 0; pv = p4[0]; continue; }
                    const float dlt = p4[e] - (pv + (float)cc * TWO_PI_F);
                    cc = dlt > PI_F ? -1 : (dlt < -PI_F ? 1 : 0);
                    pv = p4[e];
                }
                map |= (unsigned)(cc + 1) << (2 * cin);
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned up = __shfl_up_sync(0xffffffffu, map, o);
                if (lane >= o) {
                    unsigned comp = 0;
#pragma unroll
                    for (int cin = 0; cin < 3; cin++) comp |= ((map >> (2 * ((up >> (2 * cin)) & 3u))) & 3u) << (2 * cin);
                    map = comp;
                }
            }
            const unsigned before = __shfl_up_sync(0xffffffffu, map, 1);
            int cc = lane == 0 ? 0 : (int)((before >> 2) & 3u) - 1;
            float pv = prev_raw;
            ssy = 0.f; ssxy = 0.f;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float val = p4[e];
                if (!(lane == 0 && e == 0)) {
                    const float dlt = p4[e] - (pv + (float)cc * TWO_PI_F);
                    cc = dlt > PI_F ? -1 : (dlt < -PI_F ? 1 : 0);
                    val = p4[e] + (float)cc * TWO_PI_F;
                } else {
                    cc = 0;
                }
                pv = p4[e];
                ssy += val;
                ssxy += val * (float)(4 * lane + e);
            }
        }
        tsy = warp_sum(ssy);
        tsxy = warp_sum(ssxy);
    }
    // sums in float (an error in sum(y) reaches `a` scaled by 0.01, one in sum(xy) by 1e-4), the cancelling final step in double
    const double n = 128.0, sx1 = n * (n - 1.0) / 2.0, sx2 = (n - 1.0) * n * (2.0 * n - 1.0) / 6.0;
    const double lb = ((double)tsxy - sx1 * (double)tsy) / (sx2 - sx1 * sx1);   // Frame.hpp:422 (sums, not means)
    const double la = (double)tsy - lb * sx1;                                    // Frame.hpp:423
    {
        // FrameScal as five 8-byte words: {kc, m0} {th0, theta} {rot_theta} {a} {b}
        uint2 wv;
        if (lane == 0) wv = make_uint2((unsigned)kc, (unsigned)m0);
        else if (lane == 1) wv = make_uint2(__float_as_uint(theta0), __float_as_uint(theta));
        else if (lane == 2) wv = make_uint2(__float_as_uint(rot.x), __float_as_uint(rot.y));
        else if (lane == 3) wv = make_uint2((unsigned)__double2loint(la), (unsigned)__double2hiint(la));
        else wv = make_uint2((unsigned)__double2loint(lb), (unsigned)__double2hiint(lb));
        if (lane < 5) reinterpret_cast<uint2 *>(fscal + frame)[lane] = wv;
    }
    if (TAPS) {
        if (taps.scal != nullptr && lane == 0) {
            float *sc = taps.scal + (size_t)frame * 48;
            sc[0] = (float)((double)kc / (double)P.pf_den); sc[1] = (float)la; sc[2] = (float)lb; sc[3] = theta;
            sc[5] = (float)kc; sc[6] = 0.f; sc[7] = 0.f;
            sc[16] = (float)m0; sc[32] = theta0;
        }
        if (taps.chan != nullptr)
            for (int i = lane; i < 256; i += 32)
                taps.chan[(size_t)frame * 256 + i] = cis_turns((lb * (double)(i < 128 ? i : i - 256) + la) * 0.15915494309189533577);
    }
}

// Completes the `synced` debug tap of rx_demod512_kernel / rx_acquire512w_kernel: they store every sample with the symbol's
// full rotation exp(-j 2 pi beta_s j / 512); apply the per-symbol constant phase Psi_s and theta so that the tap equals
// the reference's buffer after freq_shift + cp_freq_sinh + pr_phase_sinh.
// Symbols below `first_full` (the preamble while the paired acquire kernel serves it) carry the fractional-bin part only.
__global__ void rx_synced_fixup2_kernel(const Params P, int n_frames, const RxTaps taps, int first_full) {
    const int frame = blockIdx.x;
    if (frame >= n_frames || taps.synced == nullptr || taps.scal == nullptr) return;
    const float *sc = taps.scal + (size_t)frame * 48;
    const int nsym = P.n_sym_rx;
    float s_th, c_th;
    sincosf(-sc[3], &s_th, &c_th);
    int mi[kRxMaxSym + 1];
    for (int t = 0; t < nsym; t++) mi[t] = (int)sc[16 + t];
    for (int s = 0; s < nsym; s++) {
        const float psi = sym_turns(sc + 32, mi, s);
        const double ms = s < first_full ? (double)mi[s] : 0.0;
        float2 *x = taps.synced + (size_t)frame * P.rx_len + (size_t)s * 640;
        for (int j = threadIdx.x; j < 640; j += blockDim.x) {
            const float2 r = cmul(cis_neg_turns((double)psi + ms * (double)j / 512.0), make_float2(c_th, s_th));
            x[j] = cmul(x[j], r);
        }
    }
}

}  // namespace cofdmk
