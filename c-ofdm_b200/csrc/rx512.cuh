// rx512.cuh -- the fused aligned-frame receive kernel (fft 512 / cp 128 / 8 pilots / 256 data sub-carriers,
// one preamble symbol): samples -> payload bytes in ONE pass, each sample read from HBM once.
//
// Reference chain being fused (main.cpp:60-80 == rx.cpp:200-220):
//   pilot_freq_sinh (Frame.hpp:285-337)  coarse CFO = arg-max of the 640-point preamble spectrum in 8 windows
//   freq_shift      (Frame.hpp:340-348)  x[i] *= exp(-j 2pi shift i)
//   cp_freq_sinh    (Frame.hpp:238-263)  per symbol: phi = arg sum conj(x[j]) x[j+512]; x[j] *= exp(-j phi j/512)
//   pr_phase_sinh   (Frame.hpp:265-274)  theta = arg sum conj(ref[i]) x[i]; x *= exp(-j theta)
//   chan_char_lq    (Frame.hpp:389-434)  line fit through the preamble's sub-carrier phases
//   message.fft     (Frame.hpp:276-282, Frame.cpp:73-96)  8 x FFT-512, pilot normalisation, segment correction
//   equalise + demod (rx.cpp:214-216, modulation.cpp:53-87)
//
// Work split of one CTA = one frame:
//   * FFT teams: two warps per PAIR of OFDM symbols; all FFT arithmetic is packed f32x2 (FADD2/FMUL2/FFMA2)
//     with symbol A in the low and symbol B in the high half of every register pair.  Each lane owns one
//     radix-8 butterfly per pass (warp h of the team owns butterflies 32h..32h+31); the two warps meet at
//     a named barrier (bar.sync id, 64) around each shared-memory exchange.
//   * two "coarse" warps compute the 640-point spectrum (10x8x8 Stockham) and the 8 arg-maxima
//     concurrently with the FFT warps, on their own TMA copy of the preamble.
//   That concurrency is possible because the per-sample rotation of symbol s,
//       nu_s = shift + phi_s/(2 pi 512),  phi_s = Arg(C_s exp(-j 2pi shift 512)),  C_s = raw CP correlation,
//   equals  (Arg(C_s)/(2 pi) + m_s)/512  for an INTEGER m_s: the fractional-bin part is known from the
//   symbol's own CP correlation, and the coarse estimate only contributes m_s whole FFT bins, which is an
//   index shift (and a factor (-j)^m_s) applied after the transform.
//   Per-symbol constant phases cancel between a data bin and its segment pilot (Frame.cpp:89-92) except
//   for message symbol 0, whose pilots are the reference of every segment, and the preamble.
// Three block-wide barriers per frame; everything else is warp-local.
#pragma once
#include "compat.cuh"
#include "params.h"
#include "fft.cuh"
#include "modem.cuh"

namespace cofdmk {

constexpr int kRxMaxSym = 16;
constexpr int kRxMaxPair = kRxMaxSym / 2;
constexpr int kCoarseWarps = 2;
constexpr int kPairSlots = 2 * kFft512Slots;   // float2 slots of one symbol pair

// The chain runs either as ONE kernel per frame (MODE 0) or as TWO kernels that together still read every
// sample exactly once (MODE 1 + MODE 2):
//   MODE 1 "acquire": the preamble only -- coarse CFO, its FFT, theta, the channel line -> 48 bytes per frame
//   MODE 2 "demod":   the message symbols only -- CP correlation, rotation, FFT, pilots, equalise, demap
// The split removes the two in-CTA waits of the fused form (message warps idling until the coarse estimate
// and the channel fit of the preamble are done) and lets 4 demod CTAs (instead of 3 fused ones) share an SM.
struct FrameScal {
    int kc, m0;             // coarse shift numerator; whole-bin shift of the preamble
    float th0, theta;       // Arg of the preamble's CP correlation (turns); pr_phase_sinh angle (radians)
    float2 rot_theta;       // exp(-j theta)
    double a, b;            // chan_char_lq line
};

struct RxMisc {
    uint64_t mbar[kRxMaxSym + 1];
    float4 qtab[kRxMaxSym][8];       // per FFT warp: Q^r (r<8) -- (reA, reB, imA, imB)
    float4 wtab[kRxMaxSym][8];       // per FFT warp: equaliser coefficient per segment (reA, reB, imA, imB)
    float4 pabsm4[kRxMaxSym / 4];    // sum |pilot| per MESSAGE symbol (index s - 1), zero beyond the last one
    float2 pilots[kRxMaxSym][8];     // pilot bins per symbol (shifted bins, before the (-j)^m factor)
    float pabs0;                     // sum |pilot| of the preamble (chan_char tap of the sync-less form)
    float theta_t[kRxMaxSym];        // Arg(C_s) in turns
    int mshift[kRxMaxSym];           // m_s
    int amax[8];
    int kc;                          // coarse shift numerator: shift = kc / pf_den
    double a, b;                     // chan_char_lq line
    float2 rot_theta;                // exp(-j theta)
    float theta;
    // chan_char_lq is computed by warps 0..3 together (one sub-carrier phase per lane)
    float4 cpart[kRxMaxPair][2];     // CP correlation partial sums of a team's two warps (A.re, A.im, B.re, B.im)
    float2 zpart[4];                 // partial sums of the pr_phase_sinh correlation
    float ph[128];                   // raw phases arg(pr[i]/mod_preamble[i])
    float sypart[4], sxypart[4];
    int jumppart[4];
    // MODE 2: frame-wide factors computed once per CTA (by warp 1) instead of once per warp
    FrameScal fsc;                   // the acquire kernel's hand-over
    float2 ltab[32];                 // exp(-j b lane)
    float2 rot1ee[8];                // rot_1 * exp(-j(b 32 e' + a)) per segment (rot_1 = constant phase of message symbol 0)
};

COFDM_HD int rx512_npair(int nsym) { return (nsym + 1) / 2; }
// symbols handled by a kernel of the given mode, for a frame of nsym_all symbols (preamble included)
COFDM_HD int rx512_mode_nsym(int nsym_all, int mode) { return mode == 1 ? 1 : (mode == 2 ? nsym_all - 1 : nsym_all); }
COFDM_HD int rx512_threads(int nsym_all, int mode = 0) {
    return 32 * (2 * rx512_npair(rx512_mode_nsym(nsym_all, mode)) + (mode == 2 ? 0 : kCoarseWarps));
}
COFDM_HD size_t rx512_smem_bytes(int nsym_all, int mode = 0) {
    const int np = rx512_npair(rx512_mode_nsym(nsym_all, mode));
    return (size_t)np * kPairSlots * sizeof(float2)                  // symbol pairs / FFT work planes
           + (mode == 2 ? 0 : 2 * 640 * sizeof(float2))              // coarse-CFO scratch SA, SB
           + (size_t)np * 512                                        // demapped symbols
           + sizeof(RxMisc);
}
COFDM_HD constexpr int rx512_max_threads(int maxsym, int mode) {
    return 32 * (2 * (((mode == 1 ? 1 : (mode == 2 ? maxsym - 1 : maxsym)) + 1) / 2) + (mode == 2 ? 0 : kCoarseWarps));
}
COFDM_HD constexpr int rx512_min_blocks(int maxsym, int mode) { return maxsym > 9 ? 1 : (mode == 1 ? 8 : (mode == 2 ? 4 : 3)); }

// 640 (or n) samples of one symbol -> shared memory without TMA: int16 wire format, or any source that
// is not 16-byte aligned (a frame cut out of a capture at an arbitrary sample).
template <int FMT>
COFDM_DEV void load_symbol_direct(float2 *dst, const void *src_frame, int sym, int tid, int nthr) {
    if (FMT == kCI16) {
        const unsigned *src = reinterpret_cast<const unsigned *>(src_frame) + (size_t)sym * 640;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            for (int idx = tid; idx < 160; idx += nthr) {
                const uint4 raw = __ldg(s4 + idx);
                const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int e = 0; e < 4; e++)
                    dst[4 * idx + e] = make_float2((float)(short)(w[e] & 0xffffu), (float)(short)(w[e] >> 16));
            }
        } else {
            for (int i = tid; i < 640; i += nthr) {
                const unsigned w = __ldg(src + i);
                dst[i] = make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
            }
        }
    } else {
        const float2 *src = reinterpret_cast<const float2 *>(src_frame) + (size_t)sym * 640;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            const float4 *s4 = reinterpret_cast<const float4 *>(src);
            float4 *d4 = reinterpret_cast<float4 *>(dst);
            for (int idx = tid; idx < 320; idx += nthr) d4[idx] = __ldg(s4 + idx);
        } else {
            for (int i = tid; i < 640; i += nthr) dst[i] = __ldg(src + i);
        }
    }
}

// one sample of a symbol staged in shared memory: float2, or -- RAW16: a bulk copy of int16 wire data -- int16 I,Q
template <bool RAW16>
COFDM_DEV float2 staged_sample(const float2 *region, int idx) {
    if (RAW16) {
        const unsigned w = reinterpret_cast<const unsigned *>(region)[idx];
        return make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
    }
    return region[idx];
}

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

// multiply by (-j)^m
COFDM_DEV float2 mul_negj_pow(float2 v, int m) {
    switch (m & 3) {
        case 0: return v;
        case 1: return make_float2(v.y, -v.x);
        case 2: return make_float2(-v.x, -v.y);
        default: return make_float2(-v.y, v.x);
    }
}

// TAPS: compile the debug/parity taps in (tests) or out (production, benchmark).
template <int FMT, bool USE_TMA, int MAXSYM, bool TAPS, int MODE = 0>
__global__ void __launch_bounds__(rx512_max_threads(MAXSYM, MODE), rx512_min_blocks(MAXSYM, MODE))
rx_fused512_kernel(const Params P, const void *__restrict__ samples, long long frame_stride /*samples*/,
                   int n_frames, uint8_t *__restrict__ out_bytes, unsigned long long *__restrict__ ambiguous,
                   const RxTaps taps, const int sync_less, FrameScal *__restrict__ fscal = nullptr) {
    // sync_less != 0: FRAME_FORM::read / OFDM_FORM::read (Frame.cpp:201-208,239-242): no CFO, phase or channel
    // correction at all -- CP strip, FFT, pilot normalisation, segment correction, demap.
    COFDM_DYN_SMEM(smem_raw);
    constexpr bool RAW16 = FMT == kCI16 && USE_TMA;   // int16 wire data bulk-copied as is, widened when read
    static_assert(!RAW16 || MODE == 2, "raw int16 staging is for the demod kernel (the coarse warps transform their copy in place)");
    constexpr int kMaxTeams = (rx512_max_threads(MAXSYM, MODE) / 32 - (MODE == 2 ? 0 : kCoarseWarps)) / 2;
    const int nsym_all = P.n_sym_rx;             // 1 preamble + num_symb message symbols
    const int sym0 = MODE == 2 ? 1 : 0;          // first frame symbol this kernel handles
    const int nsym = rx512_mode_nsym(nsym_all, MODE);   // number of symbols it handles (local index 0..nsym-1)
    const int npair = rx512_npair(nsym);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.x;
    if (frame >= n_frames) return;

    float2 *X = reinterpret_cast<float2 *>(smem_raw);
    float2 *SA = X + (size_t)npair * kPairSlots;
    float2 *SB = SA + 640;
    uint8_t *symbuf = reinterpret_cast<uint8_t *>(MODE == 2 ? SA : SB + 640);
    RxMisc *M = reinterpret_cast<RxMisc *>(symbuf + (size_t)npair * 512);

    if (tid < kRxMaxSym) reinterpret_cast<float *>(M->pabsm4)[tid] = 0.f;   // ordered before the writes by barrier #2
    const size_t sample_bytes = (FMT == kCI16) ? 4 : 8;
    const char *frame_src = reinterpret_cast<const char *>(samples) + (size_t)frame * (size_t)frame_stride * sample_bytes;
    const bool is_coarse = MODE != 2 && warp >= 2 * npair;
    const float inv2pi = 0.15915494309189533577f;

    // ---- stage the frame: one bulk copy per symbol + a private copy of the preamble for the coarse warps.
    //      Every team arms and issues its own copies (lane 0 of its first warp), so no block-wide barrier
    //      is needed before the first wait: a team-wide named barrier makes the mbarrier init visible. ----
    if (USE_TMA) {
        if (is_coarse) {
            if (tid == 64 * npair) {
                mbar_init(&M->mbar[nsym], 1);
                mbar_fence_init();
                mbar_arrive_expect_tx(&M->mbar[nsym], 640 * (unsigned)sample_bytes);
                tma_load_1d(SA, frame_src, 640 * (unsigned)sample_bytes, &M->mbar[nsym]);
            }
            named_bar_sync(1, 32 * kCoarseWarps);
        } else {
            if ((warp & 1) == 0 && lane == 0) {
                const int s0 = warp, s1 = warp + 1;        // local symbols 2*team and 2*team+1
                mbar_init(&M->mbar[s0], 1);
                if (s1 < nsym) mbar_init(&M->mbar[s1], 1);
                mbar_fence_init();
                mbar_arrive_expect_tx(&M->mbar[s0], 640 * (unsigned)sample_bytes);
                tma_load_1d(X + (size_t)(s0 >> 1) * kPairSlots, frame_src + (size_t)(sym0 + s0) * 640 * sample_bytes, 640 * (unsigned)sample_bytes, &M->mbar[s0]);
                if (s1 < nsym) {
                    mbar_arrive_expect_tx(&M->mbar[s1], 640 * (unsigned)sample_bytes);
                    tma_load_1d(X + (size_t)(s0 >> 1) * kPairSlots + 640, frame_src + (size_t)(sym0 + s1) * 640 * sample_bytes, 640 * (unsigned)sample_bytes, &M->mbar[s1]);
                }
            }
            team_bar_sync<kMaxTeams>(warp >> 1);
        }
    }

    if (MODE == 2 && warp == 1) {
        // while the bulk copies are in flight: the acquire kernel's scalars and the factors every warp would
        // otherwise recompute -- exp(-j b lane) and exp(-j(b 32 e' + a))
        const FrameScal f = fscal[frame];          // written by the acquire kernel (same stream, earlier launch)
        if (lane == 0) M->fsc = f;
        M->ltab[lane] = cis_neg_turns_f((float)(f.b * (double)lane) * inv2pi);
        if (lane < 8) M->rot1ee[lane] = cis_neg_turns_f((float)((f.b * (double)(32 * (lane < 4 ? lane : lane - 8)) + f.a) * 0.15915494309189533577));
    }
    const int team = warp >> 1, h = warp & 1;      // FFT warps: team = symbol pair, h = which half of the butterflies
    const int lA = 2 * team, lB = 2 * team + 1;   // local symbol indices (buffers, mbarriers)
    const bool hasB = lB < nsym;
    const int A = sym0 + lA, B = sym0 + lB;        // frame symbol indices (0 = preamble)
    float2 *Wre = X + (size_t)team * kPairSlots, *Wim = Wre + kFft512Slots;
    float thA = 0.f, thB = 0.f;                    // Arg(C_s) in turns

    if (is_coarse) {
        // ================= coarse CFO: 640-point spectrum of the received preamble, CP included =================
        const int ct = tid - 64 * npair, cn = 32 * kCoarseWarps, cw = warp - 2 * npair;
        if (USE_TMA) mbar_wait(&M->mbar[nsym], 0);
        else { load_symbol_direct<FMT>(SA, frame_src, 0, ct, cn); named_bar_sync(1, cn); }
        if (sync_less) { if (ct == 0) M->kc = 0; goto coarse_done; }
        stockham_pass<10, false>(SA, SB, 640, 1, P.tw_pf, ct, cn);
        named_bar_sync(1, cn);
        stockham_pass<8, false, true>(SB, SA, 640, 10, P.tw_pf, ct, cn);   // twiddle index <= 7*9*8 < 640
        named_bar_sync(1, cn);
        {   // last pass (radix 8, ns = 80): only |X|^2 is kept, as float[640] in SB
            float *mag = reinterpret_cast<float *>(SB);
            for (int j = ct; j < 80; j += cn) {
                float2 v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    v[q] = SA[j + 80 * q];
                    if (q > 0) v[q] = cmul(v[q], __ldg(&P.tw_pf[q * j]));       // q*j <= 7*79 < 640
                }
                dft8<false>(v);
#pragma unroll
                for (int q = 0; q < 8; q++) mag[j + 80 * q] = cnorm2(v[q]);
            }
        }
        named_bar_sync(1, cn);
        {   // arg-max of |spectrum| in the pilot windows, first maximum wins (Frame.hpp:311-331)
            const float *mag = reinterpret_cast<const float *>(SB);
            const int np = P.num_pilot_subc, half = P.pf_size / 2;
            for (int wi = cw; wi < np; wi += kCoarseWarps) {
                const int win = wi < np / 2 ? wi : wi + 1;             // window np/2 (DC) is skipped
                int lo = P.pf_border0 + win * P.pf_pilot_w;
                const int hi = lo + P.pf_pilot_w;
                if (win == 0 && lo < 0) lo = 0;
                float best = -1.0f;
                int besti = 0x7fffffff;
                for (int ks = lo + lane; ks < hi; ks += 32) {          // ks = fft-shifted index
                    const float mv = mag[ks < half ? ks + half : ks - half];
                    if (mv > best) { best = mv; besti = ks; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
                }
                if (lane == 0) M->amax[wi] = besti;
            }
        }
        named_bar_sync(1, cn);
        if (ct == 0) {
            int k = 0;
            for (int i = 0; i < P.num_pilot_subc; i++) k += M->amax[i];
            M->kc = k - P.num_pilot_subc * (P.pf_size / 2);           // shift = kc / pf_den (Frame.hpp:332-334)
        }
    coarse_done:;
    } else {
        // ================= FFT team: symbols A and B, this warp owns butterflies t = lane + 32 h =================
        float2 *xa = Wre, *xb = Wre + 640;
        if (USE_TMA) {
            mbar_wait(&M->mbar[lA], 0);
            if (hasB) mbar_wait(&M->mbar[lB], 0);
        } else {
            load_symbol_direct<FMT>(xa, frame_src, A, lane + 32 * h, 64);
            if (hasB) load_symbol_direct<FMT>(xb, frame_src, B, lane + 32 * h, 64);
            team_bar_sync<kMaxTeams>(team);
        }
        // Every sample is read from shared memory ONCE: this warp's 8 body samples per symbol (pass-1 layout,
        // j = 128 + t + 64 r) and the two CP samples j = t, t + 64, which pair with r = 6, 7 in the CP
        // correlation (Frame.hpp:251-253).  The team's two partial correlations meet in shared memory.
        const int t = lane + 32 * h;
        float2 ra[8], rb[8], cpa[2], cpb[2];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            ra[r] = staged_sample<RAW16>(xa, 128 + t + 64 * r);
            rb[r] = hasB ? staged_sample<RAW16>(xb, 128 + t + 64 * r) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            cpa[c] = staged_sample<RAW16>(xa, t + 64 * c);
            cpb[c] = hasB ? staged_sample<RAW16>(xb, t + 64 * c) : make_float2(0.f, 0.f);
        }
        {
            float2 ca = make_float2(0.f, 0.f), cb = make_float2(0.f, 0.f);
            cmac_conj(ca, cpa[0], ra[6]); cmac_conj(ca, cpa[1], ra[7]);
            cmac_conj(cb, cpb[0], rb[6]); cmac_conj(cb, cpb[1], rb[7]);
            ca = warp_sum(ca);
            cb = warp_sum(cb);
            if (lane == 0) M->cpart[team][h] = make_float4(ca.x, ca.y, cb.x, cb.y);
            team_bar_sync<kMaxTeams>(team);           // also: the whole team has read its inputs, the planes may be reused
            const float4 p0 = M->cpart[team][0], p1 = M->cpart[team][1];
            const float2 sel = (lane & 1) ? make_float2(p0.z + p1.z, p0.w + p1.w) : make_float2(p0.x + p1.x, p0.y + p1.y);
            const float ang = sync_less ? 0.f : fast_atan2_turns(sel.y, sel.x);   // one evaluation serves both symbols
            thA = __shfl_sync(0xffffffffu, ang, 0);
            thB = hasB ? __shfl_sync(0xffffffffu, ang, 1) : 0.f;
            if (lane == 0 && h == 0) { M->theta_t[A] = thA; if (hasB) M->theta_t[B] = thB; }
        }
        const float nuA = thA * (1.0f / 512.0f), nuB = thB * (1.0f / 512.0f);   // fractional-bin rotation, turns/sample
        // phasor table of the warp: Q^r = exp(-j 2pi nu 64 r), r<8, both symbols
        float4 *qt = M->qtab[warp];
        {
            const bool forB = (lane & 8) != 0;
            const float2 ph = cis_neg_turns_f((forB ? nuB : nuA) * (float)(64 * (lane & 7)));
            if (lane < 16) {
                float *dst = reinterpret_cast<float *>(&qt[lane & 7]);
                dst[forB ? 1 : 0] = ph.x;
                dst[forB ? 3 : 2] = ph.y;
            }
        }
        __syncwarp();
        const pc Pt = fast_cis_turns2(-nuA * (float)(128 + t), -nuB * (float)(128 + t));
        // rotate in registers: v[r] = x[128 + t + 64 r] * exp(-j 2pi nu (128 + t + 64 r)), both symbols per instruction
        pc v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const float4 qr = qt[r];
            pc Qr; Qr.re = make_float2(qr.x, qr.y); Qr.im = make_float2(qr.z, qr.w);
            v[r] = cmul(make_pc(ra[r], rb[r]), cmul(Pt, Qr));
        }
        if (TAPS || A == 0) {
            // CP samples j = t and j = t + 64: exp(-j 2pi nu j) = P(t) conj(Q^2) resp. P(t) conj(Q^1)
            const float4 q1 = qt[1], q2 = qt[2];
            pc Q1, Q2;
            Q1.re = make_float2(q1.x, q1.y); Q1.im = make_float2(-q1.z, -q1.w);
            Q2.re = make_float2(q2.x, q2.y); Q2.im = make_float2(-q2.z, -q2.w);
            const pc w0 = cmul(Pt, Q2), w1 = cmul(Pt, Q1);
            if (TAPS && taps.synced != nullptr) {
                // debug tap, completed by rx_synced_fixup_kernel: samples rotated by the fractional-bin part only
                float2 *da = taps.synced + (size_t)frame * P.rx_len + (size_t)A * 640, *db = da + 640;
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    da[128 + t + 64 * r] = pc_a(v[r]);
                    if (hasB) db[128 + t + 64 * r] = pc_b(v[r]);
                }
                da[t] = cmul(cpa[0], pc_a(w0));
                da[t + 64] = cmul(cpa[1], pc_a(w1));
                if (hasB) { db[t] = cmul(cpb[0], pc_b(w0)); db[t + 64] = cmul(cpb[1], pc_b(w1)); }
            }
            if (A == 0) {
                // pr_phase_sinh, CP part: conj(ref[j]) x[j] exp(-j 2pi nu' j); the missing factor
                // exp(-j 2pi m_0 j / 512) is applied once the coarse shift is known.  Parked in shared memory.
                float2 *zs = reinterpret_cast<float2 *>(M->wtab);      // 128 float2 = wtab[0..7]; reused before wtab is
                zs[t] = cmulc(cmul(cpa[0], pc_a(w0)), __ldg(&P.preamble_td[t]));
                zs[t + 64] = cmulc(cmul(cpa[1], pc_a(w1)), __ldg(&P.preamble_td[t + 64]));
            }
        }
        team_fft512p_head<false>(v, P.tw_p1, t);
        team_fft512p_tail<false, kMaxTeams, !TAPS>(v, Wre, Wim, P.tw_p2, lane, h, team);   // the grid tap wants all 512 bins
    }
    __syncthreads();                               // #2: spectra (shifted by the unknown m_s) and kc are ready

    if (MODE == 2 && tid == 0) {                   // symbol 0's entries, for the constant phase of message symbol 0
        const FrameScal fs = M->fsc;
        M->theta_t[0] = fs.th0; M->mshift[0] = fs.m0;
        M->a = fs.a; M->b = fs.b; M->rot_theta = fs.rot_theta; M->theta = fs.theta;
    }
    const int kc = MODE == 2 ? M->fsc.kc : M->kc;
    int mA = 0, mB = 0;
    if (!is_coarse) {
        // m_s and the reference's phi_s (Frame.hpp:254): phi_t = theta_t - 512 shift + m_s in (-0.5, 0.5]
        const float sh512 = (float)kc * P.pf_bins512;
        mA = (int)ceilf(-(thA - sh512) - 0.5f);
        mB = (int)ceilf(-(thB - sh512) - 0.5f);
        if (MODE == 2 && warp == 1 && lane < 8) {
            // constant phase of message symbol 0 (team 0's symbol A; its pilots are the reference of every segment):
            // Psi_1 = sym_turns(.., 1) = 1.25 (theta_0 + m_0) mod 1
            float acc = M->fsc.th0 * (640.0f / 512.0f);
            acc -= rintf(acc);
            const float psi1 = acc + (float)((5 * M->fsc.m0) & 3) * 0.25f;
            const float2 rot1 = cmul(mul_negj_pow(cis_neg_turns_f(psi1), mA), M->fsc.rot_theta);
            M->rot1ee[lane] = cmul(rot1, M->rot1ee[lane]);
        }
        if (h == 0) {
            if (lane == 0) { M->mshift[A] = mA; if (hasB) M->mshift[B] = mB; }
            // pilot bins of both symbols
            float pa = 0.f, pb = 0.f;
            if (lane < 8) {
                const int pbin = __ldg(&P.pilot_bin[lane]);
                const int sa = spec_slot((pbin + mA) & 511), sb = spec_slot((pbin + mB) & 511);
                const float2 va = make_float2(Wre[sa].x, Wim[sa].x), vb = make_float2(Wre[sb].y, Wim[sb].y);
                M->pilots[A][lane] = va;
                pa = sqrtf(cnorm2(va));
                if (hasB) { M->pilots[B][lane] = vb; pb = sqrtf(cnorm2(vb)); }
            }
            pa = warp_sum(pa);
            pb = warp_sum(pb);
            if (lane == 0) {
                float *pm = reinterpret_cast<float *>(M->pabsm4);
                if (A >= 1) pm[A - 1] = pa; else M->pabs0 = pa;
                if (hasB) pm[B - 1] = pb;
            }
        }
    }
    if (MODE == 2) {
        // nothing: theta and the channel line come from the acquire kernel
    } else if (sync_less) {
        if (tid == 0) { M->a = 0.0; M->b = 0.0; M->rot_theta = make_float2(1.f, 0.f); M->theta = 0.f; }
    } else if (warp < 4) {
        // ---- pr_phase_sinh (Frame.hpp:265-274) and chan_char_lq (Frame.hpp:389-434) on warps 0..3 together:
        //      thread gi = 32*warp + lane owns data sub-carriers gi and 128+gi of the preamble. ----
        // theta = arg sum_{i<640} conj(ref[i]) y[i]; body part by Parseval:
        //   sum_n conj(r[n]) y[n] = (1/sqrt 512) sum_k conj(R[k]) Y[k], R = tx grid of the preamble,
        //   Y[k] = (-j)^m0 X'[k + m0] the true spectrum of the preamble (symbol 0 has no other constant phase)
        const int m0 = (int)ceilf(-(M->theta_t[0] - (float)kc * P.pf_bins512) - 0.5f);
        const float2 *P0re = X, *P0im = X + kFft512Slots;             // planes of team 0, symbol A = low half
        const int gi = 32 * warp + lane;
        float2 d0, z;
        {
            const int s0 = spec_slot((__ldg(&P.data_bin[gi]) + m0) & 511), s1 = spec_slot((__ldg(&P.data_bin[128 + gi]) + m0) & 511);
            d0 = cmulc(mul_negj_pow(make_float2(P0re[s0].x, P0im[s0].x), m0), __ldg(&P.mod_preamble[gi]));
            const float2 d1 = cmulc(mul_negj_pow(make_float2(P0re[s1].x, P0im[s1].x), m0), __ldg(&P.mod_preamble[128 + gi]));
            z = cadd(d0, d1);
            if (gi < 8) {
                const int sp = spec_slot((__ldg(&P.pilot_bin[gi]) + m0) & 511);
                z = cadd(z, cscale(mul_negj_pow(make_float2(P0re[sp].x, P0im[sp].x), m0), P.pilot_ampl));
            }
            z = cscale(z, 0.04419417382415922028f);                   // 1/sqrt(512)
            const float2 *zs = reinterpret_cast<const float2 *>(M->wtab);
            z = cadd(z, cmul(zs[gi], __ldg(&P.tw_fft[(m0 * gi) & 511])));   // CP part, j = gi < 128
        }
        z = warp_sum(z);
        if (lane == 0) M->zpart[warp] = z;
        named_bar_sync(2, 128);
        z = cadd(cadd(M->zpart[0], M->zpart[1]), cadd(M->zpart[2], M->zpart[3]));
        const float inv = rsqrtf(fmaxf(cnorm2(z), 1e-30f));
        const float2 rot = make_float2(z.x * inv, -z.y * inv);
        if (tid == 0) { M->rot_theta = rot; if (TAPS) M->theta = atan2f(z.y, z.x); }
        // phase[gi] = arg(pr[gi]/mod_preamble[gi])  (Frame.hpp:403-405)
        const float2 dr = cmul(d0, rot);
        const float ph = fast_atan2_turns(dr.y, dr.x) * 6.28318530717958647692f;
        M->ph[gi] = ph;
        // one-step unwrap (Frame.hpp:407-414): nothing moves unless some raw step exceeds pi
        const float PI_F = 3.14159265358979323846f, TWO_PI_F = 6.28318530717958647692f;
        const float prev = __shfl_up_sync(0xffffffffu, ph, 1);
        const bool jump = lane > 0 && fabsf(ph - prev) > PI_F;
        const unsigned jm = __ballot_sync(0xffffffffu, jump);
        const float sy = warp_sum(ph), sxy = warp_sum(ph * (float)gi);
        if (lane == 0) { M->sypart[warp] = sy; M->sxypart[warp] = sxy; M->jumppart[warp] = jm != 0u; }
        named_bar_sync(2, 128);
        if (warp == 0) {
            // steps across the three warp boundaries
            bool bj = false;
            if (lane >= 1 && lane < 4) bj = fabsf(M->ph[32 * lane] - M->ph[32 * lane - 1]) > PI_F;
            const bool any = (__ballot_sync(0xffffffffu, bj) != 0u) || M->jumppart[0] || M->jumppart[1] || M->jumppart[2] || M->jumppart[3];
            float tsy = (M->sypart[0] + M->sypart[1]) + (M->sypart[2] + M->sypart[3]);
            float tsxy = (M->sxypart[0] + M->sxypart[1]) + (M->sxypart[2] + M->sxypart[3]);
            if (any) {
                // slow path: the adjustment is a 3-state chain (state = multiple of 2pi carried by the previous
                // element); each lane builds the transition map of its 4 elements for every incoming state, the
                // maps are composed across lanes by a warp scan, then replayed.
                float p4[4];
#pragma unroll
                for (int e = 0; e < 4; e++) p4[e] = M->ph[4 * lane + e];
                const float prev_raw = __shfl_up_sync(0xffffffffu, p4[3], 1);
                unsigned map = 0;
#pragma unroll
                for (int cin = 0; cin < 3; cin++) {
                    int c = cin - 1;
                    float pv = prev_raw;
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        if (lane == 0 && e == 0) { c = 0; pv = p4[0]; continue; }
                        const float dlt = p4[e] - (pv + (float)c * TWO_PI_F);
                        c = dlt > PI_F ? -1 : (dlt < -PI_F ? 1 : 0);
                        pv = p4[e];
                    }
                    map |= (unsigned)(c + 1) << (2 * cin);
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
                int c = lane == 0 ? 0 : (int)((before >> 2) & 3u) - 1;
                float pv = prev_raw, ssy = 0.f, ssxy = 0.f;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    float val = p4[e];
                    if (!(lane == 0 && e == 0)) {
                        const float dlt = p4[e] - (pv + (float)c * TWO_PI_F);
                        c = dlt > PI_F ? -1 : (dlt < -PI_F ? 1 : 0);
                        val = p4[e] + (float)c * TWO_PI_F;
                    } else {
                        c = 0;
                    }
                    pv = p4[e];
                    ssy += val;
                    ssxy += val * (float)(4 * lane + e);
                }
                tsy = warp_sum(ssy);
                tsxy = warp_sum(ssxy);
            }
            // sums of Frame.hpp:416-421; float partial sums are enough (an error in sum(y) reaches `a` scaled by
            // 0.01, one in sum(xy) by 1e-4), the cancelling final step is done in double
            if (lane == 0) {
                const double n = 128.0, sx = n * (n - 1.0) / 2.0, sx2 = (n - 1.0) * n * (2.0 * n - 1.0) / 6.0;
                const double b = ((double)tsxy - sx * (double)tsy) / (sx2 - sx * sx);   // Frame.hpp:422 (sums, not means)
                M->b = b;
                M->a = (double)tsy - b * sx;                                            // Frame.hpp:423
            }
        }
    }
    __syncthreads();                               // #3: pilots, a, b, theta are ready

    const double la = M->a, lb = M->b;
    float g;                                       // pilot amplitude normaliser over all message symbols (Frame.cpp:76-80)
    {
        const float4 p0 = M->pabsm4[0], p1 = M->pabsm4[1], p2 = M->pabsm4[2], p3 = M->pabsm4[3];   // zero beyond the last symbol
        g = (((p0.x + p0.y) + (p0.z + p0.w)) + ((p1.x + p1.y) + (p1.z + p1.w))) + (((p2.x + p2.y) + (p2.z + p2.w)) + ((p3.x + p3.y) + (p3.z + p3.w)));
        g *= P.inv_pilot_norm;
    }
    const float2 rot_theta = M->rot_theta;

    if (TAPS) {
        if (taps.scal != nullptr && tid == 0) {
            float *sc = taps.scal + (size_t)frame * 48;
            if (MODE != 2) {
                sc[0] = (float)((double)kc / (double)P.pf_den); sc[1] = (float)la; sc[2] = (float)lb; sc[3] = M->theta;
                sc[5] = (float)kc; sc[6] = 0.f; sc[7] = 0.f;
            }
            if (MODE != 1) sc[4] = g;
            for (int s = sym0; s < sym0 + nsym; s++) { sc[16 + s] = (float)M->mshift[s]; sc[32 + s] = M->theta_t[s]; }
        }
        if (MODE == 2) {
            // the channel taps belong to the acquire kernel
        } else if (taps.chan != nullptr && !sync_less) {
            for (int i = tid; i < 256; i += blockDim.x)
                taps.chan[(size_t)frame * 256 + i] = cis_turns((lb * (double)(i < 128 ? i : i - 256) + la) * 0.15915494309189533577);
        }
        if (MODE != 2 && taps.chan != nullptr && sync_less) {
            // PREAMBLE_FORM::chan_char (Frame.hpp:375-385) on the preamble as it stands: pr = preamble.fft()
            // (own pilot normalisation, Frame.cpp:76-84; coef == 1), chan_est[i] = pr[i] / mod_preamble[i]
            const float gp = M->pabs0 / (8.0f * P.pilot_ampl);
            for (int i = tid; i < 256; i += blockDim.x) {
                const int sl = spec_slot(__ldg(&P.data_bin[i]));
                const float2 y = cscale(make_float2(X[sl].x, X[kFft512Slots + sl].x), 1.0f / gp);
                const float2 mp = __ldg(&P.mod_preamble[i]);
                taps.chan[(size_t)frame * 256 + i] = cscale(cmulc(y, mp), 1.0f / cnorm2(mp));
            }
        }
    }
    if (MODE == 1) {
        if (tid == 0) {
            FrameScal o;
            o.kc = kc; o.m0 = sync_less ? 0 : M->mshift[0]; o.th0 = M->theta_t[0]; o.theta = M->theta;
            o.rot_theta = rot_theta; o.a = la; o.b = lb;
            fscal[frame] = o;
        }
        return;
    }
    if (is_coarse) return;

    if (TAPS && taps.grid != nullptr) {
        // FFT_buf after FFT_FORM::read's normalisation: every bin of every message symbol, fully rotated
        const int s = h ? B : A;
        if (s >= 1 && s < nsym_all && (h == 0 || hasB)) {
            const float psi = sym_turns(M->theta_t, M->mshift, s);
            const int ms = h ? mB : mA;
            const float2 rs = cscale(cmul(mul_negj_pow(cis_neg_turns_f(psi), ms), rot_theta), 1.0f / g);
            float2 *dst = taps.grid + ((size_t)frame * (nsym_all - 1) + (s - 1)) * 512;
            for (int k = lane; k < 512; k += 32) {
                const int sl = spec_slot((k + ms) & 511);
                dst[k] = cmul(h ? make_float2(Wre[sl].y, Wim[sl].y) : make_float2(Wre[sl].x, Wim[sl].x), rs);
            }
        }
    }

    // ---- equaliser coefficients per segment (Frame.cpp:89-92 + rx.cpp:214-216) ----
    //   out = (X/g) / ((P[s,p]/g)/(P[1,p]/g)) / H_i = X * [P[1,p] rot1 / (P[s,p] g)] * conj(H_i),
    //   conj(H_i) = exp(-j(b i' + a)), i' = lane + 32 e', e' = e (e<4) or e-8: split into a per-segment
    //   factor folded into the coefficient and a per-lane factor exp(-j b lane).
    float4 *wt = M->wtab[warp];
    if (lane < 8) {
        float2 p1e;                                // P[1,p] * rot_1 * exp(-j(b 32 e' + a))
        if (MODE == 2) {
            p1e = cmul(M->pilots[1][lane], M->rot1ee[lane]);
        } else {
            // constant phase of message symbol 0 (its pilots are the reference of every segment)
            const float psi1 = sym_turns(M->theta_t, M->mshift, 1);
            const float2 rot1 = cmul(mul_negj_pow(cis_neg_turns_f(psi1), M->mshift[1]), rot_theta);
            const int ep = lane < 4 ? lane : lane - 8;
            const float2 ee = cis_neg_turns_f((float)((lb * (double)(32 * ep) + la) * 0.15915494309189533577));
            p1e = cmul(cmul(M->pilots[1][lane], rot1), ee);
        }
        const float2 psa = M->pilots[A][lane], psb = M->pilots[hasB ? B : A][lane];
        const float2 wa = cscale(cmulc(p1e, psa), __fdividef(1.0f, cnorm2(psa) * g));
        const float2 wb = cscale(cmulc(p1e, psb), __fdividef(1.0f, cnorm2(psb) * g));
        wt[lane] = make_float4(wa.x, wb.x, wa.y, wb.y);
    }
    const float2 Ll = MODE == 2 ? M->ltab[lane] : cis_neg_turns_f((float)(lb * (double)lane) * inv2pi);
    __syncwarp();

    // ---- equalise + hard demap (modulation.cpp:53-87); warp h handles segments 4h..4h+3 of both symbols;
    //      symbol A of team 0 is the preamble ----
    const int mod = P.mod_type;
    const DemapK dk = make_demapk(mod);
    uint8_t *sbA = symbuf + (size_t)team * 512, *sbB = sbA + 256;
    const bool doA = A >= 1, count_amb = ambiguous != nullptr;
    int n_amb = 0;
#pragma unroll
    for (int ee = 0; ee < 4; ee++) {
        const int e = 4 * h + ee;
        const int seg0 = __ldg(&P.data_bin[32 * e]);                  // segment e is 32 consecutive bins
        int idx = lane;                                               // data index inside the segment
        float2 Lv = Ll;
        pc x;
        if (MODE == 2 && mA == mB) {
            // the usual case, one 64-bit load per plane.  The lane takes the bin congruent to itself mod 32 (the
            // segment holds each residue once), so a half-warp reads 16 aligned consecutive bins: no bank conflict
            // whatever the shift m is.  exp(-j b idx) then comes from the CTA's table.
            const int b0 = seg0 + mA;
            idx = (lane - b0) & 31;
            const int sl = spec_slot((b0 + idx) & 511);
            x.re = Wre[sl]; x.im = Wim[sl];
            Lv = M->ltab[idx];
        } else {
            const int sa = spec_slot((seg0 + lane + mA) & 511), sb = spec_slot((seg0 + lane + mB) & 511);
            if (mA == mB) { x.re = Wre[sa]; x.im = Wim[sa]; }
            else { x.re = make_float2(Wre[sa].x, Wre[sb].y); x.im = make_float2(Wim[sa].x, Wim[sb].y); }
        }
        const int i = idx + 32 * e;
        const float4 w4 = wt[e];
        pc w; w.re = make_float2(w4.x, w4.y); w.im = make_float2(w4.z, w4.w);
        const pc z = cmul(cmul(x, w), Lv);
        if (doA) {
            if (TAPS && taps.constell != nullptr) taps.constell[((size_t)frame * (nsym_all - 1) + (A - 1)) * 256 + i] = pc_a(z);
            sbA[i] = (uint8_t)demap_fast(pc_a(z), dk);
        }
        if (hasB) {
            if (TAPS && taps.constell != nullptr) taps.constell[((size_t)frame * (nsym_all - 1) + (B - 1)) * 256 + i] = pc_b(z);
            sbB[i] = (uint8_t)demap_fast(pc_b(z), dk);
        }
        if (count_amb) n_amb += (doA && demap_ambiguous(pc_a(z), dk) ? 1 : 0) + (hasB && demap_ambiguous(pc_b(z), dk) ? 1 : 0);
    }
    if (count_amb) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_amb += __shfl_xor_sync(0xffffffffu, n_amb, o);
        if (lane == 0 && n_amb) atomicAdd(ambiguous, (unsigned long long)n_amb);
    }
    team_bar_sync<kMaxTeams>(team);
    // ---- pack: 8 consecutive symbols of `mod` bits = `mod` whole bytes, MSB first (modulation.cpp:90-125);
    //      warp h packs symbol (h ? B : A) ----
    {
        const int s = h ? B : A;
        if (s >= 1 && (h == 0 || hasB)) {
            const uint8_t *sb = h ? sbB : sbA;
            const uint2 raw = *reinterpret_cast<const uint2 *>(sb + 8 * lane);
            uint8_t *dst = out_bytes + (size_t)frame * P.bytes_per_frame + (size_t)(s - 1) * 32 * mod + (size_t)lane * mod;
            if (mod == 4) {
                // 16-QAM: wire byte k = (symbol 2k << 4) | symbol 2k+1.  Each word holds four 4-bit symbols, one per byte:
                // fold neighbours into bytes 0 and 2, then close the gap -- a dozen bit operations instead of the loop.
                const unsigned ux = ((raw.x << 4) & 0x00f000f0u) | ((raw.x >> 8) & 0x000f000fu);
                const unsigned uy = ((raw.y << 4) & 0x00f000f0u) | ((raw.y >> 8) & 0x000f000fu);
                const unsigned lo = (ux & 0xffu) | ((ux >> 8) & 0xff00u), hi = (uy & 0xffu) | ((uy >> 8) & 0xff00u);
                *reinterpret_cast<unsigned *>(dst) = lo | (hi << 16);
            } else if (mod == 2) {
                // QPSK: wire byte k = four 2-bit symbols, first symbol in the top bits
                const unsigned b0 = ((raw.x & 3u) << 6) | ((raw.x >> 4) & 0x30u) | ((raw.x >> 14) & 0xcu) | (raw.x >> 24);
                const unsigned b1 = ((raw.y & 3u) << 6) | ((raw.y >> 4) & 0x30u) | ((raw.y >> 14) & 0xcu) | (raw.y >> 24);
                *reinterpret_cast<unsigned short *>(dst) = (unsigned short)(b0 | (b1 << 8));
            } else {
                unsigned long long bits = 0;
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const unsigned sym = ((e < 4 ? raw.x : raw.y) >> (8 * (e & 3))) & 0xffu;
                    bits = (bits << mod) | (unsigned long long)sym;
                }
                for (int bq = 0; bq < mod; bq++) dst[bq] = (uint8_t)(bits >> (8 * (mod - 1 - bq)));
            }
        }
    }
}

// Completes the `synced` debug tap: the fused kernel stored every sample rotated by the fractional-bin
// frequency only; apply the integer-bin part m_s, the per-symbol constant phase and theta so that the
// tap equals the reference's buffer after freq_shift + cp_freq_sinh + pr_phase_sinh.
__global__ void rx_synced_fixup_kernel(const Params P, int n_frames, const RxTaps taps) {
    const int frame = blockIdx.x;
    if (frame >= n_frames || taps.synced == nullptr || taps.scal == nullptr) return;
    const float *sc = taps.scal + (size_t)frame * 48;
    const int nsym = P.n_sym_rx;
    float s_th, c_th;
    sincosf(-sc[3], &s_th, &c_th);
    for (int s = 0; s < nsym; s++) {
        int mi[kRxMaxSym];
        for (int t = 0; t < nsym; t++) mi[t] = (int)sc[16 + t];
        const float psi = sym_turns(sc + 32, mi, s);
        const float ms = sc[16 + s];
        float2 *x = taps.synced + (size_t)frame * P.rx_len + (size_t)s * 640;
        for (int j = threadIdx.x; j < 640; j += blockDim.x) {
            const float2 r = cmul(cis_neg_turns((double)psi + (double)ms * (double)j / 512.0), make_float2(c_th, s_th));
            x[j] = cmul(x[j], r);
        }
    }
}

}  // namespace cofdmk
