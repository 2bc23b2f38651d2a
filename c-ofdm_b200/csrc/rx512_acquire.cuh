// rx512_acquire.cuh -- the "acquire" half of the split receive chain, TWO frames per CTA.
//
// Per frame the acquire step is the preamble only (rx512.cuh, MODE 1): coarse CFO from the 640-point
// spectrum (pilot_freq_sinh, Frame.hpp:285-337), the CP correlation and FFT-512 of the preamble
// (cp_freq_sinh :238-263), pr_phase_sinh (:265-274) and the channel line chan_char_lq (:389-434); it hands
// 48 bytes of scalars (FrameScal) to the demod kernel.  A frame has ONE preamble, so a one-frame CTA leaves
// the high half of every packed f32x2 register pair empty.  Here frame 2i rides in the low half and frame
// 2i+1 in the high half of
//   * the FFT team (warps 0,1): both preambles' FFT-512 in one packed transform, and
//   * the coarse team (warps 2,3): both 640-point spectra in one packed 10x8x8 Stockham transform,
// which halves the instruction count and the shared-memory traffic per frame.  The coarse team takes its
// pass-1 inputs from the same staged copy the FFT team reads (one barrier before the FFT team reuses the
// planes), so the preamble is staged once.  The channel fits of the two frames run side by side on all four
// warps; their serial tails run on warp 0 (frame 2i) and warp 1 (frame 2i+1) at the same time.
#pragma once
#include "rx512.cuh"

namespace cofdmk {

constexpr int kAcqThreads = 128;
constexpr int kAcqC1Slots = 704;      // pass-1 output plane, row stride 11 instead of 10 (bank-conflict-free scatter)

struct AcqMisc {
    uint64_t mbar[2];
    float4 qtab[2][8];               // per FFT warp: Q^r (r<8) -- (reA, reB, imA, imB)
    float4 cpart[2];                 // CP correlation partial sums of the two FFT warps (A.re, A.im, B.re, B.im)
    float2 zs[2][128];               // per frame: CP part of the pr_phase_sinh correlation
    float2 zpart[2][4];
    float ph[2][128];                // raw phases arg(pr[i]/mod_preamble[i])
    float sypart[2][4], sxypart[2][4];
    int jumppart[2][4];
    int amax[2][8];
    int kc[2];
    float theta_t[2];
    float theta[2];
    float2 rot_theta[2];
    double a[2], b[2];
};

COFDM_HD constexpr size_t rx512_acquire_smem_bytes() {
    return (size_t)(kPairSlots + 2 * kAcqC1Slots + 2 * 640) * sizeof(float2) + sizeof(AcqMisc);
}

template <int FMT, bool USE_TMA, bool TAPS>
__global__ void __launch_bounds__(kAcqThreads, 6)
rx_acquire512x2_kernel(const Params P, const void *__restrict__ samples, long long frame_stride /*samples*/,
                       int n_frames, const RxTaps taps, FrameScal *__restrict__ fscal) {
    COFDM_DYN_SMEM(smem_raw);
    constexpr bool RAW16 = FMT == kCI16 && USE_TMA;   // int16 wire data bulk-copied as is, widened when read
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fA = 2 * blockIdx.x;
    if (fA >= n_frames) return;
    const bool hasB = fA + 1 < n_frames;

    float2 *X = reinterpret_cast<float2 *>(smem_raw);
    float2 *Wre = X, *Wim = X + kFft512Slots;
    float2 *C1re = X + kPairSlots, *C1im = C1re + kAcqC1Slots;
    float2 *C2re = C1im + kAcqC1Slots, *C2im = C2re + 640;
    AcqMisc *M = reinterpret_cast<AcqMisc *>(C2im + 640);

    const size_t sample_bytes = (FMT == kCI16) ? 4 : 8;
    const char *srcA = reinterpret_cast<const char *>(samples) + (size_t)fA * (size_t)frame_stride * sample_bytes;
    const char *srcB = srcA + (size_t)frame_stride * sample_bytes;
    float2 *xa = X, *xb = X + 640;

    // ---- stage both preambles (640 samples each) ----
    if (USE_TMA) {
        if (tid == 0) {
            mbar_init(&M->mbar[0], 1);
            mbar_init(&M->mbar[1], 1);
            mbar_fence_init();
            mbar_arrive_expect_tx(&M->mbar[0], 640 * (unsigned)sample_bytes);
            tma_load_1d(xa, srcA, 640 * (unsigned)sample_bytes, &M->mbar[0]);
            if (hasB) {
                mbar_arrive_expect_tx(&M->mbar[1], 640 * (unsigned)sample_bytes);
                tma_load_1d(xb, srcB, 640 * (unsigned)sample_bytes, &M->mbar[1]);
            }
        }
        __syncthreads();
        mbar_wait(&M->mbar[0], 0);
        if (hasB) mbar_wait(&M->mbar[1], 0);
    } else {
        load_symbol_direct<FMT>(xa, srcA, 0, tid, kAcqThreads);
        if (hasB) load_symbol_direct<FMT>(xb, srcB, 0, tid, kAcqThreads);
        __syncthreads();
    }

    const float2 zero2 = make_float2(0.f, 0.f);
    if (warp < 2) {
        // ================= FFT team: the two preambles, this warp owns butterflies t = lane + 32 h =================
        const int h = warp, t = lane + 32 * h;
        float2 ra[8], rb[8], cpa[2], cpb[2];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            ra[r] = staged_sample<RAW16>(xa, 128 + t + 64 * r);
            rb[r] = hasB ? staged_sample<RAW16>(xb, 128 + t + 64 * r) : zero2;
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            cpa[c] = staged_sample<RAW16>(xa, t + 64 * c);
            cpb[c] = hasB ? staged_sample<RAW16>(xb, t + 64 * c) : zero2;
        }
        {
            float2 ca = zero2, cb = zero2;
            cmac_conj(ca, cpa[0], ra[6]); cmac_conj(ca, cpa[1], ra[7]);
            cmac_conj(cb, cpb[0], rb[6]); cmac_conj(cb, cpb[1], rb[7]);
            ca = warp_sum(ca);
            cb = warp_sum(cb);
            if (lane == 0) M->cpart[h] = make_float4(ca.x, ca.y, cb.x, cb.y);
        }
        named_bar_sync(2, kAcqThreads);               // #1: every input has been read; the planes may be reused
        float thA, thB;
        {
            const float4 p0 = M->cpart[0], p1 = M->cpart[1];
            const float2 sel = (lane & 1) ? make_float2(p0.z + p1.z, p0.w + p1.w) : make_float2(p0.x + p1.x, p0.y + p1.y);
            const float ang = fast_atan2_turns(sel.y, sel.x);
            thA = __shfl_sync(0xffffffffu, ang, 0);
            thB = hasB ? __shfl_sync(0xffffffffu, ang, 1) : 0.f;
            if (lane == 0 && h == 0) { M->theta_t[0] = thA; M->theta_t[1] = thB; }
        }
        const float nuA = thA * (1.0f / 512.0f), nuB = thB * (1.0f / 512.0f);
        float4 *qt = M->qtab[h];
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
        const pc Pt = make_pc(cis_neg_turns_f(nuA * (float)(128 + t)), cis_neg_turns_f(nuB * (float)(128 + t)));
        pc v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const float4 qr = qt[r];
            pc Qr; Qr.re = make_float2(qr.x, qr.y); Qr.im = make_float2(qr.z, qr.w);
            const pc w = cmul(Pt, Qr);
            v[r].re = make_float2(ra[r].x * w.re.x - ra[r].y * w.im.x, rb[r].x * w.re.y - rb[r].y * w.im.y);
            v[r].im = make_float2(ra[r].x * w.im.x + ra[r].y * w.re.x, rb[r].x * w.im.y + rb[r].y * w.re.y);
        }
        {
            // CP samples j = t and j = t + 64: exp(-j 2pi nu j) = P(t) conj(Q^2) resp. P(t) conj(Q^1)
            const float4 q1 = qt[1], q2 = qt[2];
            pc Q1, Q2;
            Q1.re = make_float2(q1.x, q1.y); Q1.im = make_float2(-q1.z, -q1.w);
            Q2.re = make_float2(q2.x, q2.y); Q2.im = make_float2(-q2.z, -q2.w);
            const pc w0 = cmul(Pt, Q2), w1 = cmul(Pt, Q1);
            const float2 a0 = cmul(cpa[0], pc_a(w0)), a1 = cmul(cpa[1], pc_a(w1));
            const float2 b0 = cmul(cpb[0], pc_b(w0)), b1 = cmul(cpb[1], pc_b(w1));
            if (TAPS && taps.synced != nullptr) {
                // debug tap, completed by rx_synced_fixup_kernel: samples rotated by the fractional-bin part only
                float2 *da = taps.synced + (size_t)fA * P.rx_len, *db = da + P.rx_len;
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    da[128 + t + 64 * r] = pc_a(v[r]);
                    if (hasB) db[128 + t + 64 * r] = pc_b(v[r]);
                }
                da[t] = a0; da[t + 64] = a1;
                if (hasB) { db[t] = b0; db[t + 64] = b1; }
            }
            // pr_phase_sinh, CP part: conj(ref[j]) x[j] exp(-j 2pi nu' j); the missing factor exp(-j 2pi m_0 j / 512)
            // is applied once the coarse shift is known
            const float2 r0 = __ldg(&P.preamble_td[t]), r1 = __ldg(&P.preamble_td[t + 64]);
            M->zs[0][t] = cmulc(a0, r0); M->zs[0][t + 64] = cmulc(a1, r1);
            M->zs[1][t] = cmulc(b0, r0); M->zs[1][t + 64] = cmulc(b1, r1);
        }
        team_fft512p_head<false>(v, P.tw_p1, t);
        team_fft512p_tail<false, 1, !TAPS>(v, Wre, Wim, P.tw_p2, lane, h, 0);
    } else {
        // ================= coarse team: the two 640-point spectra (CP included), packed =================
        const int ct = tid - 64, cw = warp - 2;
        pc u[10];
#pragma unroll
        for (int q = 0; q < 10; q++) u[q] = make_pc(staged_sample<RAW16>(xa, ct + 64 * q), hasB ? staged_sample<RAW16>(xb, ct + 64 * q) : zero2);
        named_bar_sync(2, kAcqThreads);               // #1
        // pass 1: radix 10, ns = 1; output o = 10 ct + q stored at slot o + o/10 = 11 ct + q
        dft10<false>(u);
#pragma unroll
        for (int q = 0; q < 10; q++) { C1re[11 * ct + q] = u[q].re; C1im[11 * ct + q] = u[q].im; }
        named_bar_sync(1, 64);
        // pass 2: radix 8, ns = 10: inputs i = j + 80 q live at slot i + i/10 = j + g + 88 q, g = j / 10
        for (int j = ct; j < 80; j += 64) {
            const int g = j / 10, k = j - 10 * g;
            pc v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                v[q].re = C1re[j + g + 88 * q]; v[q].im = C1im[j + g + 88 * q];
                if (q > 0) v[q] = cmul(v[q], __ldg(&P.tw_pf[8 * q * k]));        // 8 q k <= 504 < 640
            }
            dft8<false>(v);
#pragma unroll
            for (int q = 0; q < 8; q++) { C2re[80 * g + k + 10 * q] = v[q].re; C2im[80 * g + k + 10 * q] = v[q].im; }
        }
        named_bar_sync(1, 64);
        float2 *mag = C1re;                           // |X|^2 of (frame A, frame B), 640 entries
        for (int j = ct; j < 80; j += 64) {           // pass 3: radix 8, ns = 80
            pc v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                v[q].re = C2re[j + 80 * q]; v[q].im = C2im[j + 80 * q];
                if (q > 0) v[q] = cmul(v[q], __ldg(&P.tw_pf[q * j]));            // q j <= 553 < 640
            }
            dft8<false>(v);
#pragma unroll
            for (int q = 0; q < 8; q++) mag[j + 80 * q] = p_fma(v[q].re, v[q].re, p_mul(v[q].im, v[q].im));
        }
        named_bar_sync(1, 64);
        {   // arg-max of |spectrum| in the pilot windows, first maximum wins (Frame.hpp:311-331)
            const int np = P.num_pilot_subc, half = P.pf_size / 2;
            for (int wi = cw; wi < np; wi += 2) {
                const int win = wi < np / 2 ? wi : wi + 1;             // window np/2 (DC) is skipped
                int lo = P.pf_border0 + win * P.pf_pilot_w;
                const int hi = lo + P.pf_pilot_w;
                if (win == 0 && lo < 0) lo = 0;
                float bestA = -1.0f, bestB = -1.0f;
                int iA = 0x7fffffff, iB = 0x7fffffff;
                for (int ks = lo + lane; ks < hi; ks += 32) {          // ks = fft-shifted index
                    const float2 mv = mag[ks < half ? ks + half : ks - half];
                    if (mv.x > bestA) { bestA = mv.x; iA = ks; }
                    if (mv.y > bestB) { bestB = mv.y; iB = ks; }
                }
                // warp arg-max by two hardware reductions per frame (redux.sync): the magnitudes are non-negative floats, so
                // their bit patterns order like unsigned integers; among the lanes that hold the maximum the smallest index wins
                {
                    const unsigned ba = bestA < 0.f ? 0u : __float_as_uint(bestA), bb = bestB < 0.f ? 0u : __float_as_uint(bestB);   // a lane without an entry
                    const unsigned ma = __reduce_max_sync(0xffffffffu, ba), mb = __reduce_max_sync(0xffffffffu, bb);
                    iA = __reduce_min_sync(0xffffffffu, ba == ma ? iA : 0x7fffffff);
                    iB = __reduce_min_sync(0xffffffffu, bb == mb ? iB : 0x7fffffff);
                }
                if (lane == 0) { M->amax[0][wi] = iA; M->amax[1][wi] = iB; }
            }
        }
        named_bar_sync(1, 64);
        if (ct < 2) {
            int k = 0;
            for (int i = 0; i < P.num_pilot_subc; i++) k += M->amax[ct][i];
            M->kc[ct] = k - P.num_pilot_subc * (P.pf_size / 2);       // shift = kc / pf_den (Frame.hpp:332-334)
        }
    }
    __syncthreads();                                  // #2: both spectra (shifted by the unknown m_0) and kc are ready

    // ---- pr_phase_sinh (Frame.hpp:265-274) and chan_char_lq (Frame.hpp:389-434), both frames side by side:
    //      thread gi owns data sub-carriers gi and 128+gi of each preamble (see rx512.cuh for the algebra) ----
    const int gi = tid;
    const int nfr = hasB ? 2 : 1;
    const float PI_F = 3.14159265358979323846f, TWO_PI_F = 6.28318530717958647692f;
    int m0v[2];
    float2 d0v[2];
    const int db0 = __ldg(&P.data_bin[gi]), db1 = __ldg(&P.data_bin[128 + gi]);
    const float2 mp0 = __ldg(&P.mod_preamble[gi]), mp1 = __ldg(&P.mod_preamble[128 + gi]);
#pragma unroll
    for (int c = 0; c < 2; c++) {
        if (c >= nfr) break;
        const int m0 = (int)ceilf(-(M->theta_t[c] - (float)M->kc[c] * P.pf_bins512) - 0.5f);
        m0v[c] = m0;
        const int s0 = spec_slot((db0 + m0) & 511), s1 = spec_slot((db1 + m0) & 511);
        const float2 y0 = c ? make_float2(Wre[s0].y, Wim[s0].y) : make_float2(Wre[s0].x, Wim[s0].x);
        const float2 y1 = c ? make_float2(Wre[s1].y, Wim[s1].y) : make_float2(Wre[s1].x, Wim[s1].x);
        const float2 d0 = cmulc(mul_negj_pow(y0, m0), mp0);
        const float2 d1 = cmulc(mul_negj_pow(y1, m0), mp1);
        d0v[c] = d0;
        float2 z = cadd(d0, d1);
        if (gi < 8) {
            const int sp = spec_slot((__ldg(&P.pilot_bin[gi]) + m0) & 511);
            const float2 yp = c ? make_float2(Wre[sp].y, Wim[sp].y) : make_float2(Wre[sp].x, Wim[sp].x);
            z = cadd(z, cscale(mul_negj_pow(yp, m0), P.pilot_ampl));
        }
        z = cscale(z, 0.04419417382415922028f);                       // 1/sqrt(512)
        z = cadd(z, cmul(M->zs[c][gi], __ldg(&P.tw_fft[(m0 * gi) & 511])));   // CP part, j = gi < 128
        z = warp_sum(z);
        if (lane == 0) M->zpart[c][warp] = z;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 2; c++) {
        if (c >= nfr) break;
        const float2 z = cadd(cadd(M->zpart[c][0], M->zpart[c][1]), cadd(M->zpart[c][2], M->zpart[c][3]));
        const float inv = rsqrtf(fmaxf(cnorm2(z), 1e-30f));
        const float2 rot = make_float2(z.x * inv, -z.y * inv);
        if (tid == 0) { M->rot_theta[c] = rot; M->theta[c] = TAPS ? atan2f(z.y, z.x) : 0.f; }
        // phase[gi] = arg(pr[gi]/mod_preamble[gi])  (Frame.hpp:403-405)
        const float2 dr = cmul(d0v[c], rot);
        const float ph = fast_atan2_turns(dr.y, dr.x) * TWO_PI_F;
        M->ph[c][gi] = ph;
        // one-step unwrap (Frame.hpp:407-414): nothing moves unless some raw step exceeds pi
        const float prev = __shfl_up_sync(0xffffffffu, ph, 1);
        const bool jump = lane > 0 && fabsf(ph - prev) > PI_F;
        const unsigned jm = __ballot_sync(0xffffffffu, jump);
        const float sy = warp_sum(ph), sxy = warp_sum(ph * (float)gi);
        if (lane == 0) { M->sypart[c][warp] = sy; M->sxypart[c][warp] = sxy; M->jumppart[c][warp] = jm != 0u; }
    }
    __syncthreads();
    if (warp < nfr) {
        const int c = warp;                           // warp 0 finishes frame A, warp 1 frame B
        const float *phc = M->ph[c];
        bool bj = false;                              // steps across the three warp boundaries
        if (lane >= 1 && lane < 4) bj = fabsf(phc[32 * lane] - phc[32 * lane - 1]) > PI_F;
        const bool any = (__ballot_sync(0xffffffffu, bj) != 0u) || M->jumppart[c][0] || M->jumppart[c][1] || M->jumppart[c][2] || M->jumppart[c][3];
        float tsy = (M->sypart[c][0] + M->sypart[c][1]) + (M->sypart[c][2] + M->sypart[c][3]);
        float tsxy = (M->sxypart[c][0] + M->sxypart[c][1]) + (M->sxypart[c][2] + M->sxypart[c][3]);
        if (any) {
            // slow path: the adjustment is a 3-state chain (state = multiple of 2pi carried by the previous element);
            // each lane builds the transition map of its 4 elements for every incoming state, the maps are
            // composed across lanes by a warp scan, then replayed.
            float p4[4];
#pragma unroll
            for (int e = 0; e < 4; e++) p4[e] = phc[4 * lane + e];
            const float prev_raw = __shfl_up_sync(0xffffffffu, p4[3], 1);
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
            float pv = prev_raw, ssy = 0.f, ssxy = 0.f;
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
            tsy = warp_sum(ssy);
            tsxy = warp_sum(ssxy);
        }
        if (lane == 0) {
            // sums of Frame.hpp:416-421 in float, the cancelling final step in double (see rx512.cuh)
            const double n = 128.0, sx = n * (n - 1.0) / 2.0, sx2 = (n - 1.0) * n * (2.0 * n - 1.0) / 6.0;
            const double b = ((double)tsxy - sx * (double)tsy) / (sx2 - sx * sx);   // Frame.hpp:422 (sums, not means)
            const double a = (double)tsy - b * sx;                                  // Frame.hpp:423
            FrameScal o;
            o.kc = M->kc[c]; o.m0 = c ? m0v[1] : m0v[0]; o.th0 = M->theta_t[c]; o.theta = M->theta[c];
            o.rot_theta = M->rot_theta[c]; o.a = a; o.b = b;
            fscal[fA + c] = o;
            if (TAPS) {
                M->a[c] = a; M->b[c] = b;
                if (taps.scal != nullptr) {
                    float *sc = taps.scal + (size_t)(fA + c) * 48;
                    sc[0] = (float)((double)o.kc / (double)P.pf_den); sc[1] = (float)a; sc[2] = (float)b; sc[3] = o.theta;
                    sc[5] = (float)o.kc; sc[6] = 0.f; sc[7] = 0.f;
                    sc[16] = (float)o.m0; sc[32] = o.th0;
                }
            }
        }
    }
    if (TAPS && taps.chan != nullptr) {
        __syncthreads();
        for (int c = 0; c < nfr; c++) {
            const double la = M->a[c], lb = M->b[c];
            for (int i = tid; i < 256; i += kAcqThreads)
                taps.chan[(size_t)(fA + c) * 256 + i] = cis_turns((lb * (double)(i < 128 ? i : i - 256) + la) * 0.15915494309189533577);
        }
    }
}

}  // namespace cofdmk
