// rx512n.cuh -- the receive chain for the fft-512 geometry: ONE WARP PER OFDM SYMBOL, spectrum kept in registers.
//
// Reference chain (main.cpp:60-80 == rx.cpp:200-220): pilot_freq_sinh (Frame.hpp:285-337), freq_shift (:340-348),
// cp_freq_sinh (:238-263), pr_phase_sinh (:265-274), chan_char_lq (:389-434), message.fft (:276-282 + Frame.cpp:73-96),
// the equaliser loop (rx.cpp:214-216) and Modulation::demod (modulation.cpp:53-87).
//
// Two kernels that together read every sample exactly once:
//   rx_acquire512w_kernel  the preamble of one frame per WARP -> 56 bytes of scalars (FrameScal)
//   rx_demod512_kernel     the message symbols of one frame per CTA, one WARP per symbol -> payload bytes
// Both use warp_fft512 (fft512w.cuh): radix 16 x 16 x 2 on natural-layout packed f32x2 complex numbers, one shared-memory
// exchange private to the warp.  After the transform every sub-carrier sits in a FIXED lane and register (the acquire
// kernel's coarse shift is known before the demod kernel starts, so the whole-bin part of the CFO rotation is applied in the
// time domain as well), and the demod kernel equalises and demaps straight from registers: the spectrum never goes to
// shared memory, the only cross-warp traffic of a frame is 8 pilots + 1 float per symbol and one block barrier.
//
// Algebra: symbol s is rotated by exp(-j 2 pi beta_s j / 512), beta_s = theta_s + m_s, j = sample index in
// the symbol (CP included), theta_s = Arg(CP correlation) in turns, m_s = the integer that makes theta_s - 512 shift + m_s
// fall into (-0.5, 0.5] (Frame.hpp:254).  Per-symbol constant phases cancel between a data bin and its segment pilot
// (Frame.cpp:89-92) except for message symbol 0 (the reference of every segment): c_1 = exp(-j 2 pi 1.25 (theta_0 + m_0)) exp(-j theta).
// Equalised point of data index i (bin k, segment e):
//   z = X_s[k] * P_1[e] c_1 / (P_s[e] g) * exp(-j (b i' + a)),   i' = i (i < 128) or i - 256
// A lane's bins are k1 + 16 i + const, so exp(-j b i') splits into a per-LANE factor exp(-j b k1) and a factor per
// COMBINATION (array, lane parity, register, segment) that is folded into the segment coefficient: at most 24 coefficients
// per symbol (one per lane), one phasor per lane.
#pragma once
#include "compat.cuh"
#include "params.h"
#include "fft512w.cuh"
#include "modem.cuh"

namespace cofdmk {

constexpr int kRxMaxSym = 16;        // frame symbols (preamble + message) the kernels are dimensioned for

// what the acquire kernel hands to the demod kernel, per frame (seven 8-byte words)
struct FrameScal {
    int kc, m0;             // coarse shift numerator (shift = kc / pf_den); whole-bin shift of the preamble
    float th0, theta;       // Arg of the preamble's CP correlation (turns); pr_phase_sinh angle (radians, taps only)
    float2 rot_theta;       // exp(-j theta)
    double a, b;            // chan_char_lq line
    float2 eb1, eb8;        // exp(-j b), exp(-j 8 b): the equaliser's per-lane phasor exp(-j b k1) rides on the demod kernel's
                            // pass-1 twiddle recurrence (V^k1 -> (V exp(-j b))^k1), so it costs two products instead of nine
};

// Constant phase (turns, mod 1) carried into symbol s by freq_shift's global sample index (Frame.hpp:341-347)
// and by cp_freq_sinh's accumulated `shift` (Frame.hpp:248,261):
//     Psi_s = shift*640*s + (640/512) * sum_{t<s} phi_t,   phi_t = theta_t - 512*shift + m_t  (turns)
//           = (640/512) * sum_{t<s} (theta_t + m_t)
// -- the coarse shift cancels exactly; the integer part is reduced mod 1 in integers.
COFDM_DEV float sym_turns(const float *theta_t, const int *mshift, int s) {
    float acc = 0.f;
    int msum = 0;
    for (int t = 0; t < s; t++) {
        acc += theta_t[t] * (640.0f / 512.0f);
        acc -= rintf(acc);                       // stay within half a turn: keeps the float spacing at ~3e-8 turns
        msum += mshift[t];
    }
    return acc + (float)((5 * msum) & 3) * 0.25f;
}

// The fft-512 / 256 data / 8 pilot sub-carrier map (Frame.cpp:31-44) is fixed by the geometry; build_tables() checks
// these lists against the tables it derives from the config.
#define COFDM_F512_PILOTS(X) X(0, 33) X(1, 66) X(2, 99) X(3, 132) X(4, 380) X(5, 413) X(6, 446) X(7, 479)
// used bins that live in the ot[] registers (ot[0] of odd lanes, ot[7] of even lanes): seven data bins and two pilots
#define COFDM_F512_STRAG(X) X(0, 128) X(1, 129) X(2, 130) X(3, 131) X(4, 381) X(5, 382) X(6, 383)
constexpr int kF512Strag = 7;
constexpr int kF512MaxCombos = 24;
constexpr int kF512PfSize = 640, kF512PfW = 41, kF512PfBorder0 = 134;   // coarse-CFO windows of the geometry (Frame.hpp:311-321)

// a warp's region: staged samples (5120 B); then the FFT exchange [0, 4672) + the rotation phasors [4672, 4776); after the
// transform: [0, 64) the straggler bins, [64, 256) the segment coefficients, [256, 528) one byte per demapped symbol
constexpr int kDemodRegion = 5120;
constexpr int kDemodTabOff = kFft512wBytes;
constexpr int kDemodWtOff = 64, kDemodSymOff = 256;

struct alignas(16) DemodShared {
    uint64_t mbar[kRxMaxSym];
    float2 pil[kRxMaxSym][8];        // pilot bins per message symbol (index s - 1)
    float pabs[kRxMaxSym];           // sum |pilot| per message symbol, zero beyond the last one
    float2 lcl[16];                  // exp(-j b k1), k1 = lane >> 1
    float2 ftab[kF512MaxCombos];     // per combination: c_1 exp(-j (b off + a))
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

// Where one warp's 640 samples come from.  frame_pos (optional): record f starts at sample frame_pos[f] of `samples` instead of
// f * frame_stride -- frames detected in a capture are demodulated IN PLACE.  int16 records need not be 16-byte aligned for
// the TMA path: the bulk copy starts at the aligned address below the record and the warp reads its samples `off` bytes in
// (the copy is skipped for the rare record whose aligned span would leave [lo, hi), which is staged by plain loads).
struct RxSrc {
    const long long *frame_pos;
    const char *lo, *hi;
};

// sample `idx` of a symbol staged in shared memory: float2, or int16 I,Q wire data widened here
template <int FMT>
COFDM_DEV float2 staged_at(const void *region, int idx) {
    if (FMT == kCI16) {
        const unsigned w = reinterpret_cast<const unsigned *>(region)[idx];
        return make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
    }
    return reinterpret_cast<const float2 *>(region)[idx];
}

// exp(-j 2 pi (theta + m) J / 512) for an integer sample index J: the whole-bin part is reduced exactly in integers
COFDM_DEV float2 rot_phasor(float theta, int m, int J) {
    return fast_cis_turns(-(theta * ((float)J * (1.0f / 512.0f)) + (float)((m * J) & 511) * (1.0f / 512.0f)));
}

// The per-sample rotation of one symbol, x[l + 32 u] *= exp(-j 2 pi beta (l + 32 u) / 512), split as P(l) R^u:
//   table entry `lane` = exp(-j 2 pi beta J / 512), J = 32 * 2^lane (lanes 0..3: R, R^2, R^4, R^8), 8 (lane - 4) (lanes 4..7: U^u),
//   128 + (lane - 8) (lanes 8..15: V^v);   P(l) = exp(-j 2 pi beta (128 + l) / 512) = V^(l & 7) U^(l >> 3)   (the body starts at sample 128)
// One sincos per lane.  The other powers of R are products of at most four of the directly evaluated ones (a power formed by
// repeated squaring of R alone would carry n times the error of R).  Applied to the body samples v[n1] (index 128 + l + 32 n1).
// Returns P(l); rp[1..4] = R^1..R^4 for the callers that also rotate the cyclic prefix.
COFDM_DEV float2 rotate_body(float2 (&v)[16], float theta, int m, float2 *qt, int lane, float2 (&rp)[5]) {
    {
        const int J = lane < 4 ? (32 << lane) : (lane < 8 ? 8 * (lane - 4) : 128 + (lane - 8));
        const float2 ph = rot_phasor(theta, m, J);
        if (lane < 16) qt[lane] = ph;
    }
    __syncwarp();
    const float4 r12 = reinterpret_cast<const float4 *>(qt)[0], r48 = reinterpret_cast<const float4 *>(qt)[1];
    const float2 R1 = make_float2(r12.x, r12.y), R2 = make_float2(r12.z, r12.w), R4 = make_float2(r48.x, r48.y), R8 = make_float2(r48.z, r48.w);
    const float2 P = nmul(qt[8 + (lane & 7)], qt[4 + (lane >> 3)]);
    const float2 R3 = nmul(R2, R1), R5 = nmul(R4, R1), R6 = nmul(R4, R2), R7 = nmul(R4, R3);
    v[1] = nmul(v[1], R1); v[2] = nmul(v[2], R2); v[3] = nmul(v[3], R3); v[4] = nmul(v[4], R4);
    v[5] = nmul(v[5], R5); v[6] = nmul(v[6], R6); v[7] = nmul(v[7], R7); v[8] = nmul(v[8], R8);
    v[9] = nmul(nmul(v[9], R1), R8);   v[10] = nmul(nmul(v[10], R2), R8); v[11] = nmul(nmul(v[11], R3), R8);
    v[12] = nmul(nmul(v[12], R4), R8); v[13] = nmul(nmul(v[13], R5), R8); v[14] = nmul(nmul(v[14], R6), R8);
    v[15] = nmul(nmul(v[15], R7), R8);
    rp[0] = make_float2(1.f, 0.f); rp[1] = R1; rp[2] = R2; rp[3] = R3; rp[4] = R4;
    return P;
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

#ifndef COFDM_DEMOD_MINB
#define COFDM_DEMOD_MINB 4
#endif
// MOD: modulation order the instance is specialised for (2, 4), or 0 = any (read from the configuration)
template <int FMT, bool USE_TMA, bool TAPS, int MAXW, int MOD>
__global__ void __launch_bounds__(32 * MAXW, MAXW <= 8 ? COFDM_DEMOD_MINB : 1)
rx_demod512_kernel(const Params P, const void *__restrict__ samples, long long frame_stride /*samples*/, int n_frames,
                   uint8_t *__restrict__ out_bytes, unsigned long long *__restrict__ ambiguous, const RxTaps taps,
                   const int sync_less, const FrameScal *__restrict__ fscal, const RxSrc rs) {
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
    const size_t rec0 = rs.frame_pos != nullptr ? (size_t)__ldg(rs.frame_pos + frame) : (size_t)frame * (size_t)frame_stride;
    const char *src = reinterpret_cast<const char *>(samples) + (rec0 + (size_t)s * 640) * sample_bytes;

    // ---- stage the symbol: one TMA bulk copy issued by the warp that consumes it ----
    unsigned off = 0;                              // int16 records: bytes between the 16-byte aligned copy start and the first sample
    bool tma = USE_TMA;
    if (USE_TMA && FMT == kCI16) {
        off = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);
        tma = src - off >= rs.lo && src - off + ((640 * 4 + off + 15u) & ~15u) <= rs.hi;
        if (!tma) off = 0;
    }
    if (tma) {
        if (lane == 0) {
            const unsigned nb = (640 * (unsigned)sample_bytes + off + 15u) & ~15u;
            mbar_init(&M->mbar[warp], 1);
            mbar_fence_init();
            mbar_arrive_expect_tx(&M->mbar[warp], nb);
            tma_load_1d(region, src - off, nb, &M->mbar[warp]);
        }
    } else {
        warp_stage_symbol<FMT>(region, src, lane);
    }
    // ---- while the copy is in flight: the acquire kernel's scalars and the frame-wide tables ----
    // (lanes 0..4 fetch the first five 8-byte words of FrameScal -- {kc, m0} {th0, theta} {rot_theta} {a} {b} -- and the fields
    //  travel by shuffle to the few lanes that need them; every lane needs kc only)
    uint2 fsw = make_uint2(0u, 0u);
    if (!sync_less && lane < 5) fsw = __ldg(reinterpret_cast<const uint2 *>(fscal + frame) + lane);
    if (sync_less && lane == 2) fsw.x = 0x3f800000u;        // sync-less: kc = m0 = 0, th0 = theta = 0, rot_theta = 1, a = b = 0
    const int kc = (int)__shfl_sync(0xffffffffu, fsw.x, 0);
    const uint2 aux = __ldg(&P.lane_aux[lane]);             // .x: combination `lane` (segment, offset) + the lane's routing bits; .y: straggler `lane`
    if (tid >= nw && tid < kRxMaxSym) M->pabs[tid] = 0.f;   // unused entries (the others are written by their warps); ordered by the block barrier
    // The equaliser's factor exp(-j b i') splits into exp(-j b k1) per lane (k1 = lane >> 1 after the transform) and a factor per
    // combination.  Production instances fold the per-lane part into the transform's pass-1 twiddles (eb1, eb8 from the acquire
    // kernel); the instances with taps keep the spectrum untouched and multiply afterwards (table lcl).
    float2 eb1 = make_float2(1.f, 0.f), eb8 = eb1;
    if (!TAPS && !sync_less) {
        eb1 = __ldg(&fscal[frame].eb1);
        eb8 = __ldg(&fscal[frame].eb8);
    }
    if (TAPS && warp == 0) {
        // exp(-j b k1), k1 = 0..15
        const double fb = __hiloint2double((int)__shfl_sync(0xffffffffu, fsw.y, 4), (int)__shfl_sync(0xffffffffu, fsw.x, 4));
        const float bt = (float)fb * 0.15915494309189533577f;           // channel-line slope in turns per data index
        if (lane < 16) M->lcl[lane] = cis_neg_turns_f(bt * (float)lane);
    }
    float2 rot_theta = make_float2(1.f, 0.f);
    if (warp == nw - 1 || TAPS) {
        const int m0 = (int)__shfl_sync(0xffffffffu, fsw.y, 0);
        const float th0 = __uint_as_float(__shfl_sync(0xffffffffu, fsw.x, 1));
        rot_theta = make_float2(__uint_as_float(__shfl_sync(0xffffffffu, fsw.x, 2)), __uint_as_float(__shfl_sync(0xffffffffu, fsw.y, 2)));
        const double fa = __hiloint2double((int)__shfl_sync(0xffffffffu, fsw.y, 3), (int)__shfl_sync(0xffffffffu, fsw.x, 3));
        const double fb = __hiloint2double((int)__shfl_sync(0xffffffffu, fsw.y, 4), (int)__shfl_sync(0xffffffffu, fsw.x, 4));
        if (warp == nw - 1 && lane < P.n_combos) {
            // c_1 exp(-j (b off + a)): the constant phase of message symbol 0 (Psi_1 = 1.25 (theta_0 + m_0) mod 1), theta, the channel line
            float acc = th0 * (640.0f / 512.0f);
            acc -= rintf(acc);
            const float psi1 = acc + (float)((5 * m0) & 3) * 0.25f;
            const float2 c1 = nmul(cis_neg_turns_f(psi1), rot_theta);
            const float2 ee = cis_neg_turns_f((float)((fb * (double)((int)aux.x >> 16) + fa) * 0.15915494309189533577));
            M->ftab[lane] = nmul(c1, ee);
        }
        if (TAPS && tid == 0) { M->theta_t[0] = th0; M->mshift[0] = m0; }
    }
    __syncwarp();
    if (tma) mbar_wait(&M->mbar[warp], 0);

    // ---- the lane's 16 body samples v[n1] = x[128 + lane + 32 n1] and 4 CP samples cp[c] = x[lane + 32 c] ----
    float2 v[16], cp[4];
    const char *stg = region + off;
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) v[n1] = staged_at<FMT>(stg, 128 + lane + 32 * n1);
#pragma unroll
    for (int c = 0; c < 4; c++) cp[c] = staged_at<FMT>(stg, lane + 32 * c);
    __syncwarp();                                  // the region may now be reused (phasor table, exchange)

    // ---- CP correlation (Frame.hpp:251-253): CP sample j pairs with body sample j + 512, i.e. n1 = 12 + c ----
    float theta = 0.f;
    int m = 0;
    if (!sync_less) {
        float2 c = nmac_conj(nmac_conj(make_float2(0.f, 0.f), cp[0], v[12]), cp[1], v[13]);
        c = nmac_conj(nmac_conj(c, cp[2], v[14]), cp[3], v[15]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c = nadd(c, make_float2(__shfl_xor_sync(0xffffffffu, c.x, o), __shfl_xor_sync(0xffffffffu, c.y, o)));
        theta = fast_atan2_turns(c.y, c.x);
        // m_s and the reference's phi_s (Frame.hpp:254): phi = theta - 512 shift + m in (-0.5, 0.5]
        m = (int)ceilf(-(theta - (float)kc * P.pf_bins512) - 0.5f);
    }
    if (TAPS && lane == 0) { M->theta_t[s] = theta; M->mshift[s] = m; }

    float2 rp[5];
    const float2 pl = rotate_body(v, theta, m, reinterpret_cast<float2 *>(region + kDemodTabOff), lane, rp);
    if (TAPS && taps.synced != nullptr) {
        // debug tap, completed by rx_synced_fixup_kernel (per-symbol constant phase and theta)
        float2 *d = taps.synced + (size_t)frame * P.rx_len + (size_t)s * 640;
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) d[128 + lane + 32 * n1] = nmul(v[n1], pl);
        // CP sample j = lane + 32 c: exp(-j 2 pi beta j / 512) = P(lane) conj(R^(4 - c))
#pragma unroll
        for (int c = 0; c < 4; c++) d[lane + 32 * c] = nmul(nmulc(cp[c], rp[4 - c]), pl);
    }
    float2 mn[8], ot[8];
    {
        float2 V = __ldg(P.tw_fft + lane), V8 = __ldg(P.tw_fft + 8 * lane);
        if (!TAPS) { V = nmul(V, eb1); V8 = nmul(V8, eb8); }
        warp_fft512(v, pl, reinterpret_cast<float2 *>(region), V, V8, lane, mn, ot);
    }
    // now mn[i] = X[k1 + 16 i + (g ? 384 : 0)], ot[i] = X[k1 + 16 i + (g ? 128 : 256)], lane = 2 k1 + g (production instances:
    // times exp(-j b k1)); the region is free again

    // ---- pilots and sum |pilot| (Frame.cpp:76-80) to the CTA's shared memory; the straggler data bins to the warp's region ----
    float2 *scratch = reinterpret_cast<float2 *>(region);
    {
        float2 *pil = M->pil[warp];
        // Each of the 8 pilots and 7 straggler bins sits in ONE lane, in a register fixed by the geometry: the lane picks its
        // value by selects (no branches) and its destination from the routing bits of lane_aux, then one store serves all 15.
        float2 val = (lane & 1) ? ot[0] : ot[7];                 // the only ot[] registers that hold used bins (pilots 132, 380; stragglers)
#define COFDM_X(p, bin) if (f512_main(bin)) val = lane == f512_lane(bin) ? mn[f512_i(bin)] : val;
        COFDM_F512_PILOTS(COFDM_X)
#undef COFDM_X
        const unsigned route = aux.x >> 3;                       // [4:1] pilot / straggler number, [6:5] 0 none, 1 pilot, 2 straggler
        float2 *dst = ((route >> 5) & 2u) ? scratch + ((route >> 1) & 15u) : pil + ((route >> 1) & 15u);
        if ((route >> 5) & 3u) *dst = val;
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
    const float2 lc = TAPS ? M->lcl[lane >> 1] : make_float2(1.f, 0.f);

    // ---- the segment coefficients of this symbol (Frame.cpp:89-92 + rx.cpp:214-216), one per combination:
    //      W[q] = P_1[e] conj(P_s[e]) / (|P_s[e]|^2 g) * ftab[q],  e = segment of combination q ----
    char *wt = region + kDemodWtOff;
    if (lane < P.n_combos) {
        const int e = (int)(aux.x & 7u);
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
            const int k1 = lane >> 1, gg = lane & 1;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                dst[k1 + 16 * i + (gg ? 384 : 0)] = nmul(mn[i], rs);
                dst[k1 + 16 * i + (gg ? 128 : 256)] = nmul(ot[i], rs);
            }
        }
    }

    // ---- equalise + hard demap (modulation.cpp:53-87) straight from the registers.  Per register 16 descriptor bits:
    //      [15:7] data index (256: a dummy slot, the bin carries no data), [6:2] combination.  No branches. ----
    const DemapK dk = make_demapk(P.mod_type);
    uint8_t *sb = reinterpret_cast<uint8_t *>(region + kDemodSymOff);
    const uint4 desc = __ldg(&P.lane_desc[lane]);
    float2 *ctap = (TAPS && taps.constell != nullptr) ? taps.constell + ((size_t)frame * nw + (s - 1)) * 256 : nullptr;
    float2 xs = make_float2(0.f, 0.f);             // the straggler this lane equalises (lanes 0..6)
    unsigned ds = 256u << 7;                       // dummy slot
    if (lane < kF512Strag) {
        ds = aux.y & 0xffffu;                      // descriptor | (k1 of the lane that held the bin) << 16
        xs = TAPS ? nmul(scratch[lane], M->lcl[aux.y >> 16]) : scratch[lane];
    }
#define COFDM_EQ(X, D16)                                                                         \
    do {                                                                                         \
        const unsigned d_ = (D16);                                                               \
        const unsigned i_ = d_ >> 7;                                                             \
        const float2 z_ = nmul((X), *reinterpret_cast<const float2 *>(wt + ((d_ & 0x7cu) << 1))); \
        if (TAPS && ctap != nullptr && i_ < 256u) ctap[i_] = z_;                                 \
        sb[i_] = (uint8_t)demap_n<MOD>(z_, dk);                                                  \
    } while (0)
#define COFDM_EQ_ALL(F)                                                                          \
    F(COFDM_LC(mn[0]), desc.x & 0xffffu); F(COFDM_LC(mn[1]), desc.x >> 16);                      \
    F(COFDM_LC(mn[2]), desc.y & 0xffffu); F(COFDM_LC(mn[3]), desc.y >> 16);                      \
    F(COFDM_LC(mn[4]), desc.z & 0xffffu); F(COFDM_LC(mn[5]), desc.z >> 16);                      \
    F(COFDM_LC(mn[6]), desc.w & 0xffffu); F(COFDM_LC(mn[7]), desc.w >> 16);                      \
    F(xs, ds)
#define COFDM_LC(X) (TAPS ? nmul((X), lc) : (X))
    COFDM_EQ_ALL(COFDM_EQ);
#undef COFDM_EQ
    if (ambiguous != nullptr) {
        // optional count of boundary-ambiguous decisions (margin kAmbigMargin): the points are recomputed, off the fast path
        int n_amb = 0;
#define COFDM_AMB(X, D16)                                                                        \
        do {                                                                                     \
            const unsigned d_ = (D16);                                                           \
            if ((d_ >> 7) < 256u) n_amb += demap_ambiguous(nmul((X), *reinterpret_cast<const float2 *>(wt + ((d_ & 0x7cu) << 1))), dk) ? 1 : 0; \
        } while (0)
        COFDM_EQ_ALL(COFDM_AMB);
#undef COFDM_AMB
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_amb += __shfl_xor_sync(0xffffffffu, n_amb, o);
        if (lane == 0 && n_amb) atomicAdd(ambiguous, (unsigned long long)n_amb);
    }
#undef COFDM_EQ_ALL
#undef COFDM_LC
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
// and hands 56 bytes of scalars (FrameScal) to the demod kernel.  The lane's 20 raw samples x[lane + 32 u] are the inputs of
// BOTH transforms (radix-10 first pass of the 640-point one: butterflies lane and lane + 32; CP + radix-16 first pass of the
// 512-point one): they are read from the staged copy once and stay in registers.
// Shared memory per warp: S (5120 B: staged samples -> second coarse plane -> phasor table, phases) and A (5632 B:
// first coarse plane, rows padded 10 -> 11 -> |X|^2 -> FFT-512 exchange).
// ================================================================================================================
constexpr int kAcqwWarps = 4;
constexpr int kAcqwS = 5120, kAcqwA = 5632;
constexpr int kAcqwRegion = kAcqwS + kAcqwA;
COFDM_HD constexpr size_t rx_acquire512w_smem_bytes() { return (size_t)kAcqwWarps * kAcqwRegion + kAcqwWarps * sizeof(uint64_t); }

template <int FMT, bool USE_TMA, bool TAPS>
__global__ void __launch_bounds__(32 * kAcqwWarps, 5)
rx_acquire512w_kernel(const Params P, const void *__restrict__ samples, long long frame_stride /*samples*/, int n_frames,
                      const RxTaps taps, FrameScal *__restrict__ fscal, const int sync_less, const RxSrc rs) {
    COFDM_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int frame = blockIdx.x * kAcqwWarps + warp;
    if (frame >= n_frames) return;                     // whole warp; warps never meet at a block barrier
    char *S = reinterpret_cast<char *>(smem_raw) + (size_t)warp * kAcqwRegion;
    float2 *A = reinterpret_cast<float2 *>(S + kAcqwS);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kAcqwWarps * kAcqwRegion) + warp;
    const size_t sample_bytes = (FMT == kCI16) ? 4 : 8;
    const size_t rec0 = rs.frame_pos != nullptr ? (size_t)__ldg(rs.frame_pos + frame) : (size_t)frame * (size_t)frame_stride;
    const char *src = reinterpret_cast<const char *>(samples) + rec0 * sample_bytes;
    unsigned off = 0;                                  // see RxSrc
    bool tma = USE_TMA;
    if (USE_TMA && FMT == kCI16) {
        off = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);
        tma = src - off >= rs.lo && src - off + ((640 * 4 + off + 15u) & ~15u) <= rs.hi;
        if (!tma) off = 0;
    }
    if (tma) {
        if (lane == 0) {
            const unsigned nb = (640 * (unsigned)sample_bytes + off + 15u) & ~15u;
            mbar_init(mbar, 1);
            mbar_fence_init();
            mbar_arrive_expect_tx(mbar, nb);
            tma_load_1d(S, src - off, nb, mbar);
        }
    } else {
        warp_stage_symbol<FMT>(S, src, lane);
    }
    __syncwarp();
    if (tma) mbar_wait(mbar, 0);
    // ---- the lane's 20 raw samples raw[u] = x[lane + 32 u] ----
    float2 raw[20];
#pragma unroll
    for (int u = 0; u < 20; u++) raw[u] = staged_at<FMT>(S + off, lane + 32 * u);
    __syncwarp();                                      // S may be overwritten from here on

    int kc = 0;
    if (!sync_less) {
        // ================= coarse CFO: 640-point spectrum of the received preamble, CP included =================
        // pass 1: radix 10, butterflies j = lane (inputs raw[2 q]) and j = lane + 32 (raw[2 q + 1]); output 10 j + q at slot 11 j + q
        {
            float2 t[10];
#pragma unroll
            for (int q = 0; q < 10; q++) t[q] = raw[2 * q];
            ndft10(t);
#pragma unroll
            for (int q = 0; q < 10; q++) A[11 * lane + q] = t[q];
#pragma unroll
            for (int q = 0; q < 10; q++) t[q] = raw[2 * q + 1];
            ndft10(t);
#pragma unroll
            for (int q = 0; q < 10; q++) A[11 * (lane + 32) + q] = t[q];
        }
        __syncwarp();
        float2 *B = reinterpret_cast<float2 *>(S);
        // pass 2: radix 8, ns = 10: butterfly j (80 of them), k = j mod 10; inputs I = j + 80 q at slot I + I / 10 = (j + j / 10) + 88 q
#pragma unroll 1
        for (int j = lane; j < 80; j += 32) {
            const int gq = j / 10, k = j - 10 * gq;
            float2 t[8], w[8];
            const float2 *in = A + j + gq;
#pragma unroll
            for (int q = 0; q < 8; q++) t[q] = in[88 * q];
            npowers7(__ldg(&P.tw_pf[8 * k]), w);                          // W640^{8 k q}
#pragma unroll
            for (int q = 1; q < 8; q++) t[q] = nmul(t[q], w[q]);
            ndft8(t);
            float2 *out = B + 80 * gq + k;
#pragma unroll
            for (int q = 0; q < 8; q++) out[10 * q] = t[q];
        }
        __syncwarp();
        // pass 3: radix 8, ns = 80: only |X|^2 is kept, as float[640] in A
        float *mag = reinterpret_cast<float *>(A);
#pragma unroll 1
        for (int j = lane; j < 80; j += 32) {
            float2 t[8], w[8];
#pragma unroll
            for (int q = 0; q < 8; q++) t[q] = B[j + 80 * q];
            npowers7(__ldg(&P.tw_pf[j]), w);                              // W640^{j q}
#pragma unroll
            for (int q = 1; q < 8; q++) t[q] = nmul(t[q], w[q]);
            ndft8(t);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (q == 3 || q == 4) continue;        // bins 240..399: no pilot window reaches them (fft-shifted 134..297, 339..502)
                const float2 sq = p_mul(t[q], t[q]);
                mag[j + 80 * q] = sq.x + sq.y;
            }
        }
        __syncwarp();
        // arg-max of |spectrum| in the eight pilot windows, first maximum wins (Frame.hpp:311-331).  The windows are fixed by the
        // geometry (41 bins each from fft-shifted index 134, the DC window skipped; checked by build_tables), none straddles
        // the fft-shift seam.  Warp arg-max by two hardware reductions: the magnitudes are non-negative floats, so their bit
        // patterns order like unsigned integers; among the lanes holding the maximum the smallest index wins.
        constexpr int np = 8, half = kF512PfSize / 2;
        int ksum = 0;
#pragma unroll
        for (int wi = 0; wi < np; wi++) {
            const int win = wi < np / 2 ? wi : wi + 1;                     // window np/2 (DC) is skipped
            const int lo = kF512PfBorder0 + win * kF512PfW;                // fft-shifted index of the window's first bin
            const float *mw = mag + (lo < half ? lo + half : lo - half);
            const float m1 = mw[lane];
            const float m2 = lane < kF512PfW - 32 ? mw[32 + lane] : -1.0f;
            const bool second = m2 > m1;
            const unsigned bb = __float_as_uint(second ? m2 : m1);
            const int bi = lo + lane + (second ? 32 : 0);
            const unsigned mx = __reduce_max_sync(0xffffffffu, bb);
            ksum += __reduce_min_sync(0xffffffffu, bb == mx ? bi : 0x7fffffff);
        }
        kc = ksum - np * half;                                             // shift = kc / pf_den (Frame.hpp:332-334)
        __syncwarp();                                                      // the planes are free again
    }

    // ================= fine CFO of the preamble (cp_freq_sinh): CP sample j pairs with body sample j + 512 (u = 16 + c) =================
    float theta0 = 0.f;
    int m0 = 0;
    if (!sync_less) {
        float2 c = nmac_conj(nmac_conj(make_float2(0.f, 0.f), raw[0], raw[16]), raw[1], raw[17]);
        c = nmac_conj(nmac_conj(c, raw[2], raw[18]), raw[3], raw[19]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c = nadd(c, make_float2(__shfl_xor_sync(0xffffffffu, c.x, o), __shfl_xor_sync(0xffffffffu, c.y, o)));
        theta0 = fast_atan2_turns(c.y, c.x);
        m0 = (int)ceilf(-(theta0 - (float)kc * P.pf_bins512) - 0.5f);
    }
    float2 *qt = reinterpret_cast<float2 *>(S);
    float2 v[16], rp[5];
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) v[n1] = raw[4 + n1];
    const float2 pl = rotate_body(v, theta0, m0, qt, lane, rp);
    // CP sample j = lane + 32 c rotated: exp(-j 2 pi beta j / 512) = P(lane) conj(R^(4 - c))
    float2 ycp[4];
#pragma unroll
    for (int c = 0; c < 4; c++) ycp[c] = nmul(nmulc(raw[c], rp[4 - c]), pl);
    if (TAPS && taps.synced != nullptr) {
        // debug tap, completed by rx_synced_fixup_kernel (theta; the preamble has no other constant phase)
        float2 *d = taps.synced + (size_t)frame * P.rx_len;
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) d[128 + lane + 32 * n1] = nmul(v[n1], pl);
#pragma unroll
        for (int c = 0; c < 4; c++) d[lane + 32 * c] = ycp[c];
    }
    float2 mn[8], ot[8];
    warp_fft512(v, pl, A, P.tw_fft, lane, mn, ot);
    // now mn[] / ot[] hold the true spectrum Y of the preamble (symbol 0 has no constant phase)
    const float2 osel = (lane & 1) ? ot[0] : ot[7];                        // the only ot[] registers that hold used bins

    if (sync_less) {
        // PREAMBLE_FORM::chan_char (Frame.hpp:375-385) on the preamble as it stands: pr = preamble.fft() (own pilot
        // normalisation, Frame.cpp:76-84; coef == 1), chan_est[i] = pr[i] / mod_preamble[i]
        if (TAPS && taps.chan != nullptr) {
            float pm = 0.f;
#define COFDM_X(p, bin) if (lane == f512_lane(bin)) pm = sqrtf(cnorm2(f512_main(bin) ? mn[f512_i(bin)] : osel));
            COFDM_F512_PILOTS(COFDM_X)
#undef COFDM_X
            pm = warp_sum(pm);
            const float igp = (8.0f * P.pilot_ampl) / pm;
            const int k1 = lane >> 1, gg = lane & 1;
#pragma unroll
            for (int i = 0; i < 9; i++) {
                const int k = i < 8 ? k1 + 16 * i + (gg ? 384 : 0) : k1 + (gg ? 128 : 256 + 112);
                const int di = __ldg(&P.bin_map[k]);
                if (di >= 0) {
                    const float2 mp = __ldg(&P.mod_preamble[di]);
                    const float2 y = cscale(i < 8 ? mn[i < 8 ? i : 0] : osel, igp);
                    taps.chan[(size_t)frame * 256 + di] = cscale(cmulc(y, mp), 1.0f / cnorm2(mp));
                }
            }
        }
        return;
    }

    // ================= pr_phase_sinh: z = sum_{i<640} conj(ref[i]) y[i]; body by Parseval: (1/sqrt 512) sum_k conj(G[k]) Y[k],
    //                   G = tx grid of the preamble (P.grid_lane: conj(G) / sqrt 512 in lane order, zero on unused bins) =================
    float2 prod[8];                                    // Y conj(G) of mn[0..7]: the first 128 data sub-carriers are mn[] of the even lanes
    float2 z;
    {
        const float4 *g4 = reinterpret_cast<const float4 *>(P.grid_lane) + 5 * lane;   // mn[0..7], osel, pad
        const float4 ga = __ldg(g4), gb = __ldg(g4 + 1), gc = __ldg(g4 + 2), gd = __ldg(g4 + 3), ge = __ldg(g4 + 4);
        prod[0] = nmul(mn[0], make_float2(ga.x, ga.y)); prod[1] = nmul(mn[1], make_float2(ga.z, ga.w));
        prod[2] = nmul(mn[2], make_float2(gb.x, gb.y)); prod[3] = nmul(mn[3], make_float2(gb.z, gb.w));
        prod[4] = nmul(mn[4], make_float2(gc.x, gc.y)); prod[5] = nmul(mn[5], make_float2(gc.z, gc.w));
        prod[6] = nmul(mn[6], make_float2(gd.x, gd.y)); prod[7] = nmul(mn[7], make_float2(gd.z, gd.w));
        const float2 ps = nmul(osel, make_float2(ge.x, ge.y));
        z = nadd(nadd(nadd(prod[0], prod[1]), nadd(prod[2], prod[3])), nadd(nadd(prod[4], prod[5]), nadd(prod[6], prod[7])));
        z = nadd(z, ps);
        // CP part: conj(ref[j]) y[j], j = lane + 32 c
#pragma unroll
        for (int c = 0; c < 4; c++) z = nmac_conj(z, __ldg(&P.preamble_td[lane + 32 * c]), ycp[c]);
        // the data bins 128..131 (ot[0] of lanes 1, 3, 5, 7) belong to the first 128 sub-carriers too: their products travel
        // through shared memory to the four phase slots that hold no data (bin 0 and the pilots 33, 66, 99)
        float2 *sx = qt + 16;
        if ((lane & 1) && lane < 8) sx[lane >> 1] = ps;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) z = nadd(z, make_float2(__shfl_xor_sync(0xffffffffu, z.x, o), __shfl_xor_sync(0xffffffffu, z.y, o)));
    const float inv = rsqrtf(fmaxf(cnorm2(z), 1e-30f));
    const float2 rot = make_float2(z.x * inv, -z.y * inv);               // exp(-j theta)
    const float theta = TAPS ? atan2f(z.y, z.x) : 0.f;

    // ================= chan_char_lq: phase[i] = arg(pr[i] / mod_preamble[i]), i < 128 (Frame.hpp:403-405).  The 128 products sit
    //                   in mn[] of the 16 even lanes; each hands mn[4..7] to its odd neighbour so that every lane evaluates four =================
    const float TWO_PI_F = 6.28318530717958647692f, PI_F = 3.14159265358979323846f;
    float *phs = reinterpret_cast<float *>(qt + 32);                      // 128 phases
    {
        float2 pp[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const float2 up = make_float2(__shfl_sync(0xffffffffu, prod[4 + u].x, lane & ~1), __shfl_sync(0xffffffffu, prod[4 + u].y, lane & ~1));
            pp[u] = (lane & 1) ? up : prod[u];
        }
        __syncwarp();                                                     // sx[] is visible
        const uint2 ad = __ldg(&P.acq_desc[lane]);   // 4 x 16 bits: [7:0] phase index, [15] take straggler product [9:8] instead
        const float2 *sx = qt + 16;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned d = ((u < 2 ? ad.x : ad.y) >> (16 * (u & 1))) & 0xffffu;
            float2 pr = pp[u];
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
        // FrameScal as seven 8-byte words: {kc, m0} {th0, theta} {rot_theta} {a} {b} {exp(-j b)} {exp(-j 8 b)}
        uint2 wv;
        if (lane == 0) wv = make_uint2((unsigned)kc, (unsigned)m0);
        else if (lane == 1) wv = make_uint2(__float_as_uint(theta0), __float_as_uint(theta));
        else if (lane == 2) wv = make_uint2(__float_as_uint(rot.x), __float_as_uint(rot.y));
        else if (lane == 3) wv = make_uint2((unsigned)__double2loint(la), (unsigned)__double2hiint(la));
        else if (lane == 4) wv = make_uint2((unsigned)__double2loint(lb), (unsigned)__double2hiint(lb));
        else {
            const float2 e = cis_neg_turns(lb * (lane == 5 ? 1.0 : 8.0) * 0.15915494309189533577);
            wv = make_uint2(__float_as_uint(e.x), __float_as_uint(e.y));
        }
        if (lane < 7) reinterpret_cast<uint2 *>(fscal + frame)[lane] = wv;
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
__global__ void rx_synced_fixup_kernel(const Params P, int n_frames, const RxTaps taps) {
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
        const float2 r = cmul(cis_neg_turns((double)psi), make_float2(c_th, s_th));
        float2 *x = taps.synced + (size_t)frame * P.rx_len + (size_t)s * 640;
        for (int j = threadIdx.x; j < 640; j += blockDim.x) x[j] = cmul(x[j], r);
    }
}

}  // namespace cofdmk
