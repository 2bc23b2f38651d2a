// generic.cuh -- the any-size path: the same reference chain as rx512n.cuh / tx512w.cuh for configurations
// the fused fft-512 kernels do not cover (e.g. BASELINE.json configs[4]: fft 4096, cp 1024, 1920 data and 128
// pilot sub-carriers, 64-QAM).  One OFDM symbol does not fit a warp's registers and one frame does not fit an
// SM's shared memory here, so the chain is split into five kernels with the spectra kept in HBM between them
// (fp32 complex).  Straightforward, correct and parity-tested; NOT tuned: this path costs about 3x the HBM
// traffic of the fused one.
//
//   gen_coarse_kernel   pilot_freq_sinh (Frame.hpp:285-337): pf_size-point spectrum of the preamble, arg-maxima
//   gen_symbol_kernel   cp correlation (Frame.hpp:251-253), rotation by freq_shift + cp_freq_sinh (Frame.hpp:238-263,
//                       340-348) while loading, FFT (Frame.hpp:276-282)
//   gen_chan_kernel     pr_phase_sinh (Frame.hpp:265-274), chan_char_lq (Frame.hpp:389-434), pilot normaliser (Frame.cpp:76-80)
//   gen_demap_kernel    segment correction (Frame.cpp:87-93), equaliser (rx.cpp:214-216), demod (modulation.cpp:53-87)
//   gen_tx_kernel       FRAME_FORM::write + get / get_int16 (Frame.cpp:185-198,54-70,244-256)
#pragma once
#include "compat.cuh"
#include "params.h"
#include "fft.cuh"
#include "modem.cuh"

namespace cofdmk {

constexpr int kGenThreads = 512;

// per-frame scalars exchanged between the generic kernels (device memory)
struct GenFrame {
    int kc;                 // coarse shift numerator
    float g;                // pilot amplitude normaliser
    float theta;            // pr_phase_sinh angle (radians)
    float2 rot_theta;       // exp(-j theta)
    double a, b;            // chan_char_lq line
    float phit[kGenMaxSym];         // the reference's wrapped CP angle per symbol, in turns
    double psi[kGenMaxSym];         // constant phase carried into symbol s, in turns
};

template <bool INV>
COFDM_DEV void cta_pass(int R, const float2 *in, float2 *out, int n, int ns, const float2 *tw, int tid, int nthr) {
    switch (R) {
        case 16: stockham_pass<16, INV>(in, out, n, ns, tw, tid, nthr); break;
        case 8: stockham_pass<8, INV>(in, out, n, ns, tw, tid, nthr); break;
        case 5: stockham_pass<5, INV>(in, out, n, ns, tw, tid, nthr); break;
        case 4: stockham_pass<4, INV>(in, out, n, ns, tw, tid, nthr); break;
        default: stockham_pass<2, INV>(in, out, n, ns, tw, tid, nthr); break;
    }
}

// n-point FFT of a[0..n) in shared memory with the whole CTA; returns the buffer holding the result
template <bool INV>
COFDM_DEV float2 *cta_fft(float2 *a, float2 *b, int n, const int *radix, int nr, const float2 *tw, int tid, int nthr) {
    int ns = 1;
    for (int p = 0; p < nr; p++) {
        cta_pass<INV>(radix[p], a, b, n, ns, tw, tid, nthr);
        ns *= radix[p];
        __syncthreads();
        float2 *t = a; a = b; b = t;
    }
    return a;
}

template <int FMT>
COFDM_DEV float2 load_sample(const void *base, long long idx) {
    if (FMT == kCI16) {
        const unsigned w = __ldg(reinterpret_cast<const unsigned *>(base) + idx);
        return make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
    }
    return __ldg(reinterpret_cast<const float2 *>(base) + idx);
}

COFDM_DEV float2 block_sum(float2 v, float2 *red /* >= 32 float2 of shared memory */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float2 t = make_float2(0.f, 0.f);
    for (int i = 0; i < nw; i++) t = cadd(t, red[i]);
    return t;
}

// ---- pilot_freq_sinh ---------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(kGenThreads)
gen_coarse_kernel(const Params P, const void *__restrict__ samples, long long frame_stride, int n_frames, GenFrame *__restrict__ gf) {
    COFDM_DYN_SMEM(smem_raw);
    const int frame = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    if (frame >= n_frames) return;
    float2 *A = reinterpret_cast<float2 *>(smem_raw), *B = A + P.pf_size;
    __shared__ int amax[kMaxPilots];
    const long long base = (long long)frame * frame_stride;
    for (int i = tid; i < P.pf_size; i += nthr) A[i] = load_sample<FMT>(samples, base + i);
    __syncthreads();
    float2 *X = cta_fft<false>(A, B, P.pf_size, P.pf_radix, P.pf_nr, P.tw_pf, tid, nthr);
    const int np = P.num_pilot_subc, half = P.pf_size / 2, lane = tid & 31;
    for (int wi = tid >> 5; wi < np; wi += (nthr >> 5)) {
        const int win = wi < np / 2 ? wi : wi + 1;                     // window np/2 (DC) is skipped, Frame.hpp:326
        int lo = P.pf_border0 + win * P.pf_pilot_w;
        const int hi = lo + P.pf_pilot_w;
        if (win == 0 && lo < 0) lo = 0;
        float best = -1.0f;
        int besti = 0x7fffffff;
        for (int ks = lo + lane; ks < hi; ks += 32) {
            const float mv = cnorm2(X[ks < half ? ks + half : ks - half]);
            if (mv > best) { best = mv; besti = ks; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (lane == 0) amax[wi] = besti;
    }
    __syncthreads();
    if (tid == 0) {
        long long k = 0;
        for (int i = 0; i < np; i++) k += amax[i];
        gf[frame].kc = (int)(k - (long long)np * half);               // shift = kc / pf_den (Frame.hpp:332-334)
    }
}

// ---- per symbol: CP correlation, rotation, FFT ------------------------------------------------------------------
// spec[frame][sym][fft_size]; pre_rot[frame][ofdm_len] = fully frequency-corrected preamble samples (symbol 0)
// PRE_ONLY: only symbol 0 is transformed (its spectrum at spec[frame][fft_size]): the acquisition stage of the fft-4096 path
template <int FMT, bool PRE_ONLY = false>
__global__ void __launch_bounds__(kGenThreads)
gen_symbol_kernel(const Params P, const void *__restrict__ samples, long long frame_stride, int n_frames,
                  GenFrame *__restrict__ gf, float2 *__restrict__ spec, float2 *__restrict__ pre_rot, const int sync_less = 0) {
    COFDM_DYN_SMEM(smem_raw);
    const int sym = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
    if (frame >= n_frames) return;
    const int N = P.fft_size, CP = P.cp_size, L = P.ofdm_len;
    float2 *A = reinterpret_cast<float2 *>(smem_raw), *B = A + L;
    __shared__ float2 red[32];
    const long long base = (long long)frame * frame_stride + (long long)sym * L;
    for (int i = tid; i < L; i += nthr) A[i] = load_sample<FMT>(samples, base + i);
    __syncthreads();
    float2 c = make_float2(0.f, 0.f);
    for (int j = tid; j < CP; j += nthr) cmac_conj(c, A[j], A[j + N]);        // Frame.hpp:251-253
    c = block_sum(c, red);
    // (sync-less form, FRAME_FORM::read: no synchronisation stage at all -- the samples are transformed as they stand)
    const double fc = sync_less ? 0.0 : (double)gf[frame].kc / (double)P.pf_den;   // coarse shift, cycles per sample
    // the reference correlates after freq_shift: its angle is Arg(C exp(-j 2pi fc N)) (Frame.hpp:254)
    const float2 cr = cmul(c, cis_neg_turns(fc * (double)N));
    const double phit = sync_less ? 0.0 : (double)atan2f(cr.y, cr.x) * 0.15915494309189533577;
    if (tid == 0) gf[frame].phit[sym] = (float)phit;
    const double nu = fc + phit / (double)N;                                   // total rotation, turns per sample
    __syncthreads();
    {
        // exp(-j 2pi nu j) for j = tid + nthr i: an exact start per thread (double reduction), then a phasor recurrence
        // re-seeded every 8 steps so that its rounding error stays at the 1e-7 level
        for (int j0 = tid; j0 < L; j0 += 8 * nthr) {
            float2 ph = cis_neg_turns(nu * (double)j0);
            const float2 step = cis_neg_turns(nu * (double)nthr);
            for (int e = 0; e < 8; e++) {
                const int j = j0 + e * nthr;
                if (j >= L) break;
                const float2 y = cmul(A[j], ph);
                if (sym < P.num_pr_symb && pre_rot != nullptr) pre_rot[((size_t)frame * P.num_pr_symb + sym) * L + j] = y;   // the preamble's symbols
                if (j >= CP) B[j - CP] = y;                                    // CP strip (Frame.hpp:278-279)
                ph = cmul(ph, step);
            }
        }
    }
    __syncthreads();
    float2 *X = cta_fft<false>(B, A, N, P.fft_radix, P.fft_nr, P.tw_fft, tid, nthr);
    float2 *dst = spec + ((size_t)frame * (PRE_ONLY ? 1 : P.n_sym_rx) + sym) * N;
    for (int k = tid; k < N; k += nthr) dst[k] = X[k];
}

// ---- per frame: theta, channel line, pilot normaliser, per-symbol constant phases ------------------------------------
template <bool PRE_ONLY = false>
__global__ void __launch_bounds__(kGenThreads)
gen_chan_kernel(const Params P, int n_frames, GenFrame *__restrict__ gf, const float2 *__restrict__ spec,
                const float2 *__restrict__ pre_rot, const int sync_less = 0) {
    COFDM_DYN_SMEM(smem_raw);
    const int frame = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    if (frame >= n_frames) return;
    __shared__ float2 red[32];
    const int N = P.fft_size, L = P.ofdm_len, nsym = P.n_sym_rx, ND = P.num_data_subc, NP = P.num_pilot_subc, npr = P.num_pr_symb;
    GenFrame &G = gf[frame];
    __shared__ float2 psi_rot[kGenMaxSym];
    // constant phase carried into symbol s (turns): freq_shift's global index + cp_freq_sinh's accumulated shift
    if (tid == 0) {
        const double fc = (double)G.kc / (double)P.pf_den;
        double acc = 0.0;
        for (int s = 0; s < nsym; s++) {
            G.psi[s] = sync_less ? 0.0 : fc * (double)L * (double)s + acc;
            acc += (double)G.phit[s] * (double)L / (double)N;
            if (s < npr) psi_rot[s] = cis_neg_turns(G.psi[s]);
        }
    }
    __syncthreads();
    // pr_phase_sinh (Frame.hpp:265-274) over the whole preamble: symbol s of it still lacks its constant phase Psi_s (0 for s = 0)
    float2 z = make_float2(1.f, 0.f);
    if (!sync_less) {
        z = make_float2(0.f, 0.f);
        for (int j = tid; j < P.pf_size; j += nthr) cmac_conj(z, __ldg(&P.preamble_td[j]), cmul(pre_rot[(size_t)frame * P.pf_size + j], psi_rot[j / L]));
    }
    z = block_sum(z, red);
    if (sync_less) z = make_float2(1.f, 0.f);
    const float inv = rsqrtf(fmaxf(cnorm2(z), 1e-30f));
    const float2 rot = make_float2(z.x * inv, -z.y * inv);
    // pilot amplitude normaliser over all message symbols (Frame.cpp:76-80)
    float pa = 0.f;
    for (int i = tid; i < (PRE_ONLY ? 0 : (nsym - npr) * NP); i += nthr) {
        const int s = npr + i / NP, p = i % NP;
        pa += sqrtf(cnorm2(spec[((size_t)frame * nsym + s) * N + __ldg(&P.pilot_bin[p])]));
    }
    const float2 pas = block_sum(make_float2(pa, 0.f), red);
    // chan_char_lq (Frame.hpp:389-434): phases of the first ND/2 data sub-carriers of the preamble
    float *ph = reinterpret_cast<float *>(smem_raw);
    const int nph = ND / 2;
    const float2 *S0 = spec + (size_t)frame * (PRE_ONLY ? 1 : nsym) * N;
    for (int i = tid; i < (sync_less ? 0 : nph); i += nthr) {
        const float2 d = cmulc(cmul(S0[__ldg(&P.data_bin[i])], rot), __ldg(&P.mod_preamble[i]));
        ph[i] = atan2f(d.y, d.x);
    }
    __syncthreads();
    if (tid == 0) {
        const float PI_F = 3.14159265358979323846f, TWO_PI_F = 6.28318530717958647692f;
        double sy = 0.0, sxy = 0.0, sx = 0.0, sx2 = 0.0;
        float prev = 0.f;
        for (int i = 0; i < (sync_less ? 0 : nph); i++) {                      // Frame.hpp:407-421
            float v = ph[i];
            if (i > 0) {
                const float d = v - prev;
                if (d > PI_F) v -= TWO_PI_F; else if (d < -PI_F) v += TWO_PI_F;
            }
            prev = v;
            sy += (double)v; sxy += (double)v * (double)i; sx += (double)i; sx2 += (double)i * (double)i;
        }
        const double b = sync_less ? 0.0 : (sxy - sx * sy) / (sx2 - sx * sx);  // Frame.hpp:422 (sums, not means)
        G.b = b;
        G.a = sync_less ? 0.0 : sy - b * sx;                                   // Frame.hpp:423
        G.rot_theta = rot;
        G.theta = atan2f(z.y, z.x);
        G.g = pas.x / ((float)((nsym - npr) * NP) * P.pilot_ampl);
    }
}

// ---- per symbol: segment correction, equaliser, hard demap, bit packing -----------------------------------------------
__global__ void __launch_bounds__(kGenThreads)
gen_demap_kernel(const Params P, int n_frames, const GenFrame *__restrict__ gf, const float2 *__restrict__ spec,
                 uint8_t *__restrict__ out_bytes, unsigned long long *__restrict__ ambiguous, const RxTaps taps) {
    const int npr = P.num_pr_symb;
    const int s = npr + blockIdx.x, frame = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
    if (frame >= n_frames) return;
    const int N = P.fft_size, nsym = P.n_sym_rx, ND = P.num_data_subc, mod = P.mod_type;
    const GenFrame &G = gf[frame];
    const float2 *S1 = spec + ((size_t)frame * nsym + npr) * N, *Ss = spec + ((size_t)frame * nsym + s) * N;
    // constant rotation of message symbol 0, whose pilots are the reference of every segment (Frame.cpp:89)
    const float2 rot1 = cmul(cis_neg_turns(G.psi[npr]), G.rot_theta);
    const DemapK dk = make_demapk(mod);
    const int half = ND / 2;
    int n_amb = 0;
    uint8_t *dst = out_bytes + (size_t)frame * P.bytes_per_frame + (size_t)(s - npr) * (ND * mod / 8);
    for (int grp = tid; grp < ND / 8; grp += nthr) {
        unsigned long long bits = 0;
#pragma unroll 1
        for (int e = 0; e < 8; e++) {
            const int i = 8 * grp + e, p = i / P.seg_size;
            const int pb = __ldg(&P.pilot_bin[p]);
            const float2 p1 = cmul(S1[pb], rot1), ps = Ss[pb];
            const float2 w = cscale(cmulc(p1, ps), 1.0f / (cnorm2(ps) * G.g));
            // 1/H_i with H_i = exp(j(b i' + a)), i' = i (i < ND/2) or i - ND (Frame.hpp:425-430)
            const float2 hc = cis_neg_turns((G.b * (double)(i < half ? i : i - ND) + G.a) * 0.15915494309189533577);
            const float2 zz = cmul(cmul(Ss[__ldg(&P.data_bin[i])], w), hc);
            if (taps.constell != nullptr) taps.constell[((size_t)frame * (nsym - npr) + (s - npr)) * ND + i] = zz;
            bool amb;
            bits = (bits << mod) | (unsigned long long)demap_point(zz, dk, amb);
            n_amb += amb ? 1 : 0;
        }
        for (int bq = 0; bq < mod; bq++) dst[(size_t)grp * mod + bq] = (uint8_t)(bits >> (8 * (mod - 1 - bq)));
    }
    if (ambiguous != nullptr && n_amb) atomicAdd(ambiguous, (unsigned long long)n_amb);
    if (s == npr) {
        if (taps.chan != nullptr)
            for (int i = tid; i < ND; i += nthr)
                taps.chan[(size_t)frame * ND + i] = cis_turns((G.b * (double)(i < half ? i : i - ND) + G.a) * 0.15915494309189533577);
        if (taps.scal != nullptr && tid == 0) {
            float *sc = taps.scal + (size_t)frame * 48;
            sc[0] = (float)((double)G.kc / (double)P.pf_den); sc[1] = (float)G.a; sc[2] = (float)G.b; sc[3] = G.theta;
            sc[4] = G.g; sc[5] = (float)G.kc;
        }
    }
}

// ---- tx: one CTA per (symbol, frame); symbol index num_symb = the constant sync tone + preamble -----------------------
template <int FMT>
COFDM_DEV void gen_store(void *frame_out, long long idx, float2 v, float mult) {
    if (FMT == kCI16) {
        const unsigned pk = ((unsigned)(unsigned short)(short)__float2int_rz(v.x * mult)) | ((unsigned)(unsigned short)(short)__float2int_rz(v.y * mult) << 16);
        reinterpret_cast<unsigned *>(frame_out)[idx] = pk;
    } else {
        reinterpret_cast<float2 *>(frame_out)[idx] = v;
    }
}

template <int FMT>
__global__ void __launch_bounds__(kGenThreads)
gen_tx_kernel(const Params P, const uint8_t *__restrict__ payload, int n_frames, void *__restrict__ frames) {
    COFDM_DYN_SMEM(smem_raw);
    const int s = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
    if (frame >= n_frames) return;
    const int N = P.fft_size, CP = P.cp_size, L = P.ofdm_len;
    const size_t sb = (FMT == kCI16) ? 4 : 8;
    char *fout = reinterpret_cast<char *>(frames) + (size_t)frame * P.frame_len * sb;
    if (s == P.num_symb) {
        for (int i = tid; i < P.t2sin_size + P.pf_size; i += nthr)
            gen_store<FMT>(fout, i, i < P.t2sin_size ? __ldg(&P.t2_tone[i]) : __ldg(&P.preamble_td[i - P.t2sin_size]), P.mult);
        return;
    }
    float2 *A = reinterpret_cast<float2 *>(smem_raw), *B = A + N;
    const uint8_t *pl = payload + (size_t)frame * P.bytes_per_frame;
    const int mod = P.mod_type;
    for (int k = tid; k < N; k += nthr) {
        const int m = __ldg(&P.bin_map[k]);
        float2 v = make_float2(0.f, 0.f);                                       // Frame.cpp:55
        if (m == -2) v = make_float2(P.pilot_ampl, 0.f);                         // Frame.cpp:56-57
        else if (m >= 0) v = __ldg(&P.constell[extract_bits_sw(pl, P.bytes_per_frame, (s * P.num_data_subc + m) * mod, mod)]);
        A[k] = v;
    }
    __syncthreads();
    float2 *X = cta_fft<true>(A, B, N, P.fft_radix, P.fft_nr, P.tw_fft, tid, nthr);   // Frame.cpp:64
    const float sc = 1.0f / sqrtf((float)N);                                     // Frame.cpp:66-68
    const long long base = P.t2sin_size + P.pf_size + (long long)s * L;
    for (int n = tid; n < N; n += nthr) {
        const float2 v = cscale(X[n], sc);
        gen_store<FMT>(fout, base + CP + n, v, P.mult);                           // Frame.cpp:191-192
        if (n >= N - CP) gen_store<FMT>(fout, base + n - (N - CP), v, P.mult);    // Frame.cpp:196-197
    }
}

}  // namespace cofdmk
