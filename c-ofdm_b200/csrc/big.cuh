// big.cuh -- the receive / transmit chain for the fft-4096 geometry (BASELINE.json configs[4]: fft 4096, cp 1024, 64-QAM,
// dense pilots), every sample read from HBM ONCE.
//
// Reference chain: the same as rx512n.cuh (main.cpp:60-80 == rx.cpp:200-220).  One OFDM symbol (5120 samples, 40 KB) fills
// a CTA's shared memory and one frame (9 symbols) fills two SMs, so the frame is spread over a THREAD-BLOCK CLUSTER:
//   big_acquire_kernel  one CTA per frame: the preamble -> FrameScal (coarse CFO, fine CFO, phase lock, channel line)
//   big_demod_kernel    one CLUSTER per frame, one CTA (256 threads) per message symbol: CP correlation, rotation, FFT-4096,
//                       pilots; the two frame-wide quantities of FFT_FORM::read (Frame.cpp:73-96) -- the pilot amplitude
//                       normaliser over ALL symbols and the pilots of message symbol 0, the reference of every segment --
//                       travel between the CTAs through DISTRIBUTED SHARED MEMORY (two cluster barriers), then equalise,
//                       hard demap and bit packing straight from registers.  The spectrum never touches HBM.
//   big_tx_kernel       one CTA per symbol: bits -> grid -> backward FFT-4096 -> CP -> frame
// FFT-4096 = radix 16 x 16 x 16 over 256 threads, 16 values per thread in registers (natural-layout packed f32x2
// arithmetic, ndft16 of fft512w.cuh), two exchanges through the shared memory the staged symbol occupied; first exchange
// padded (16 values per thread at stride 17) so that every access is bank-conflict free.
// Algebra: DESIGN.md 4.1 with N = 4096, L = 5120 (L / N = 5 / 4 as in the fft-512 geometry).
#pragma once
#include "compat.cuh"
#include "params.h"
#include "fft512w.cuh"
#include "modem.cuh"
#include "rx512n.cuh"
#include "generic.cuh"

#ifndef COFDM_EMU
#include <cooperative_groups.h>
#endif

namespace cofdmk {

constexpr int kBigN = 4096, kBigCP = 1024, kBigL = 5120, kBigThreads = 256, kBigMaxSym = 8;
constexpr int kBigExchSlots = 4096 + 256;                  // 16 values per thread at stride 17
constexpr int kBigStageBytes = kBigL * 8;                  // staged symbol (cf32); the exchanges alias it
constexpr int kBigMaxData = 3840;                          // data sub-carriers per symbol the demap buffer holds
static_assert(kBigExchSlots * 8 <= kBigStageBytes, "the exchange must fit the staged symbol's memory");

// the CTA's shared memory behind the staged symbol
struct alignas(16) BigShared {
    float2 pil[kMaxPilots];          // this symbol's pilot bins
    float2 wseg[kMaxPilots];         // segment coefficients
    float2 red[8];                   // per-warp partial sums
    float pabs;                      // sum |pilot| of this symbol (read by the other CTAs of the cluster)
    float pad_[3];
    uint64_t mbar;
    uint64_t pad2_;
    uint8_t sb[kBigMaxData + 16];    // one byte per demapped data sub-carrier
};
COFDM_HD constexpr size_t big_smem_bytes() { return (size_t)kBigStageBytes + sizeof(BigShared); }

// ---- the cluster seen by one CTA.  Under the CPU thread emulator ONE emulated block of nrank * 256 threads stands for the
//      cluster: CTA barrier = a named barrier per rank, cluster barrier = the block barrier, remote shared memory = an offset.
struct BigCtx {
    int rank, nrank, tid, frame;
    unsigned char *smem;
};
#ifdef COFDM_EMU
COFDM_DEV BigCtx big_ctx() {
    BigCtx c;
    c.tid = (int)threadIdx.x % kBigThreads; c.rank = (int)threadIdx.x / kBigThreads; c.nrank = (int)blockDim.x / kBigThreads;
    c.frame = (int)blockIdx.x;
    c.smem = emu::dyn_smem() + (size_t)c.rank * big_smem_bytes();
    return c;
}
COFDM_DEV void big_cta_sync(const BigCtx &c) { emu::named_barrier(1 + c.rank, kBigThreads); }
COFDM_DEV void big_cluster_sync(const BigCtx &) { __syncthreads(); }
template <class T> COFDM_DEV const T *big_remote(const BigCtx &c, const T *p, int r) {
    return reinterpret_cast<const T *>(reinterpret_cast<const unsigned char *>(p) + ((long)r - (long)c.rank) * (long)big_smem_bytes());
}
#else
COFDM_DEV BigCtx big_ctx() {
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    extern __shared__ __align__(128) unsigned char big_smem_raw[];
    BigCtx c;
    c.tid = (int)threadIdx.x; c.rank = (int)cl.block_rank(); c.nrank = (int)cl.num_blocks();
    c.frame = (int)(blockIdx.x / cl.num_blocks());
    c.smem = big_smem_raw;
    return c;
}
COFDM_DEV void big_cta_sync(const BigCtx &) { __syncthreads(); }
COFDM_DEV void big_cluster_sync(const BigCtx &) { cooperative_groups::this_cluster().sync(); }
template <class T> COFDM_DEV const T *big_remote(const BigCtx &, const T *p, int r) {
    return cooperative_groups::this_cluster().map_shared_rank(p, r);
}
#endif

// w^0 .. w^15 of a unit phasor, at most four roundings each
COFDM_DEV void npowers15(float2 w1, float2 (&w)[16]) {
    w[0] = make_float2(1.f, 0.f); w[1] = w1;
    w[2] = nmul(w1, w1); w[3] = nmul(w[2], w1); w[4] = nmul(w[2], w[2]);
    w[5] = nmul(w[4], w1); w[6] = nmul(w[3], w[3]); w[7] = nmul(w[4], w[3]); w[8] = nmul(w[4], w[4]);
    w[9] = nmul(w[8], w1); w[10] = nmul(w[8], w[2]); w[11] = nmul(w[8], w[3]); w[12] = nmul(w[8], w[4]);
    w[13] = nmul(w[8], w[5]); w[14] = nmul(w[8], w[6]); w[15] = nmul(w[8], w[7]);
}

// exp(-j 2 pi (theta + m) J / 4096) for an integer sample index J: the whole-bin part is reduced exactly in integers
COFDM_DEV float2 big_phasor(float theta, int m, int J) {
    return fast_cis_turns(-(theta * ((float)J * (1.0f / 4096.0f)) + (float)((m * J) & 4095) * (1.0f / 4096.0f)));
}

// Forward FFT-4096 over the 256 threads of a CTA.  In: v[u] = x[j + 256 u] (thread j).  Out: v[t] = X[j + 256 t], unnormalised.
// E: kBigExchSlots float2 of shared memory nobody else touches; SYNC(): the CTA barrier.  w256 / w4096: global tables
// exp(-j 2 pi k / 256), exp(-j 2 pi k / 4096).  Stockham autosort passes (radix 16, Ns = 1, 16, 256):
//   pass p: thread j, k = j mod Ns: in[j + 256 t] * W_{16 Ns}^{k t} -> DFT-16 -> out[(j - k) 16 + k + Ns t'].
template <class SYNC>
COFDM_DEV void cta_fft4096(float2 (&v)[16], float2 *E, const float2 *__restrict__ w4096, int j, SYNC sync) {
    ndft16(v);
#pragma unroll
    for (int t = 0; t < 16; t++) E[17 * j + t] = v[t];                       // y[16 j + t] at slot i + (i >> 4)
    sync();
    {
        const float2 *r = E + j + (j >> 4);
#pragma unroll
        for (int t = 0; t < 16; t++) v[t] = r[272 * t];                      // y[j + 256 t]
        float2 w[16];
        npowers15(__ldg(w4096 + 16 * (j & 15)), w);                          // W256^{(j mod 16) t}
#pragma unroll
        for (int t = 1; t < 16; t++) v[t] = nmul(v[t], w[t]);
    }
    ndft16(v);
    sync();                                                                  // everybody has read the first exchange
    {
        float2 *wz = E + ((j >> 4) << 8) + (j & 15);
#pragma unroll
        for (int t = 0; t < 16; t++) wz[16 * t] = v[t];                      // z[(j >> 4) 256 + (j & 15) + 16 t]
    }
    sync();
    {
#pragma unroll
        for (int t = 0; t < 16; t++) v[t] = E[j + 256 * t];                  // z[j + 256 t]
        float2 w[16];
        npowers15(__ldg(w4096 + j), w);                                      // W4096^{j t}
#pragma unroll
        for (int t = 1; t < 16; t++) v[t] = nmul(v[t], w[t]);
    }
    ndft16(v);
}

// ================================================================================================================
// big_demod_kernel: see the head of this file.  Launched with cluster dimension num_symb (<= 8); grid = n_frames * num_symb.
// P.bin_role[k]: >= 0 data index, -1 null, -2 - p pilot number p.
// ================================================================================================================
template <int FMT, bool USE_TMA, bool TAPS>
__global__ void __launch_bounds__(kBigThreads, 4)
big_demod_kernel(const Params P, const void *__restrict__ samples, long long frame_stride /*samples*/, int n_frames,
                 uint8_t *__restrict__ out_bytes, unsigned long long *__restrict__ ambiguous, const RxTaps taps,
                 const FrameScal *__restrict__ fscal) {
    const BigCtx cx = big_ctx();
    const int tid = cx.tid, lane = tid & 31, warp = tid >> 5;
    const int frame = cx.frame;                                   // whole clusters only: every CTA of a cluster sees the same frame
    const int s = cx.rank + 1;                                    // frame symbol index (0 = preamble)
    unsigned char *stage = cx.smem;
    BigShared *M = reinterpret_cast<BigShared *>(cx.smem + kBigStageBytes);
    const size_t sample_bytes = (FMT == kCI16) ? 4 : 8;
    const char *src = reinterpret_cast<const char *>(samples) + ((size_t)frame * (size_t)frame_stride + (size_t)s * kBigL) * sample_bytes;
    auto sync = [&]() { big_cta_sync(cx); };

    // ---- stage the symbol ----
    if (USE_TMA) {
        if (tid == 0) {
            mbar_init(&M->mbar, 1);
            mbar_fence_init();
            mbar_arrive_expect_tx(&M->mbar, kBigL * (unsigned)sample_bytes);
            tma_load_1d(stage, src, kBigL * (unsigned)sample_bytes, &M->mbar);
        }
    } else {
        if (FMT == kCI16) for (int i = tid; i < kBigL; i += kBigThreads) reinterpret_cast<unsigned *>(stage)[i] = __ldg(reinterpret_cast<const unsigned *>(src) + i);
        else for (int i = tid; i < kBigL; i += kBigThreads) reinterpret_cast<float2 *>(stage)[i] = __ldg(reinterpret_cast<const float2 *>(src) + i);
    }
    const FrameScal fs = fscal[frame];
    sync();                                                       // mbarrier initialised / plain loads visible
    if (USE_TMA) mbar_wait(&M->mbar, 0);

    // ---- the thread's 16 body samples v[u] = x[1024 + tid + 256 u] and 4 CP samples cp[c] = x[tid + 256 c] ----
    float2 v[16], cp[4];
#pragma unroll
    for (int u = 0; u < 16; u++) v[u] = staged_at<FMT>(stage, kBigCP + tid + 256 * u);
#pragma unroll
    for (int c = 0; c < 4; c++) cp[c] = staged_at<FMT>(stage, tid + 256 * c);
    // ---- CP correlation (Frame.hpp:251-253): CP sample j pairs with body sample j + 4096, i.e. u = 12 + c ----
    float2 cc = nmac_conj(nmac_conj(make_float2(0.f, 0.f), cp[0], v[12]), cp[1], v[13]);
    cc = nmac_conj(nmac_conj(cc, cp[2], v[14]), cp[3], v[15]);
    cc = warp_sum(cc);
    if (lane == 0) M->red[warp] = cc;
    sync();                                                       // also: every thread has its samples, the stage area is free
    cc = M->red[0];
#pragma unroll
    for (int w = 1; w < 8; w++) cc = nadd(cc, M->red[w]);
    const float theta = fast_atan2_turns(cc.y, cc.x);
    // m_s and the reference's phi_s (Frame.hpp:254): phi = theta - N shift + m in (-0.5, 0.5]
    const int m = (int)ceilf(-(theta - (float)fs.kc * P.pf_binsN) - 0.5f);

    // ---- rotation x[n] *= exp(-j 2 pi (theta + m) n / 4096), n = 1024 + tid + 256 u: P(tid) R^u ----
    {
        float2 rp[16];
        npowers15(big_phasor(theta, m, 256), rp);
        const float2 pl = big_phasor(theta, m, kBigCP + tid);
        v[0] = nmul(v[0], pl);
#pragma unroll
        for (int u = 1; u < 16; u++) v[u] = nmul(nmul(v[u], rp[u]), pl);
    }
    cta_fft4096(v, reinterpret_cast<float2 *>(stage), P.tw_fft, tid, sync);
    // now v[t] = X[tid + 256 t] of the rotated symbol (constant phase Psi_s still on it: it cancels against the segment pilot)

    // ---- pilots and sum |pilot| (Frame.cpp:76-80) ----
    float pm = 0.f;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const int role = (int)__ldg(&P.bin_role[tid + 256 * t]);
        if (role <= -2) { M->pil[-2 - role] = v[t]; pm += sqrtf(cnorm2(v[t])); }
    }
    pm = warp_sum(pm);
    sync();                                                       // red[] has been read by everybody
    if (lane == 0) M->red[warp].x = pm;
    sync();
    if (tid == 0) {
        float tot = 0.f;
        for (int w = 0; w < 8; w++) tot += M->red[w].x;
        M->pabs = tot;
    }
    big_cluster_sync(cx);                                         // pilots and pabs of every symbol of the frame are published

    // ---- frame-wide: g (Frame.cpp:76-80) and the segment coefficients W[e] = P_1[e] conj(P_s[e]) / (|P_s[e]|^2 g) c_1,
    //      c_1 = exp(-j 2 pi Psi_1) exp(-j theta_pr), Psi_1 = 1.25 (theta_0 + m_0) mod 1 (Frame.cpp:89-92 + rx.cpp:214-216) ----
    float g = 0.f;
    for (int r = 0; r < cx.nrank; r++) g += *big_remote(cx, &M->pabs, r);
    g *= P.inv_pilot_norm;
    if (tid < P.num_pilot_subc) {
        float acc = fs.th0 * 1.25f;
        acc -= rintf(acc);
        const float psi1 = acc + (float)((5 * fs.m0) & 3) * 0.25f;
        const float2 c1 = nmul(cis_neg_turns_f(psi1), fs.rot_theta);
        const float2 p1 = big_remote(cx, M->pil, 0)[tid], ps = M->pil[tid];
        const float2 w = nscale(nmulc(p1, ps), __fdividef(1.0f, cnorm2(ps) * g));
        M->wseg[tid] = nmul(w, c1);
    }
    big_cluster_sync(cx);                                         // wseg visible; nobody reads remote shared memory after this point

    // ---- equalise + hard demap (modulation.cpp:53-87) straight from the registers ----
    const int ND = P.num_data_subc, nw = cx.nrank;
    const DemapK dk = make_demapk(P.mod_type);
    const float bt = (float)(fs.b * 0.15915494309189533577), at0 = (float)(fs.a * 0.15915494309189533577 - rint(fs.a * 0.15915494309189533577));
    const float inv_seg = 1.0f / (float)P.seg_size;
    float2 *ctap = (TAPS && taps.constell != nullptr) ? taps.constell + ((size_t)frame * nw + (s - 1)) * ND : nullptr;
    int n_amb = 0;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const int role = (int)__ldg(&P.bin_role[tid + 256 * t]);
        if (role >= 0) {
            const int e = (int)(((float)role + 0.5f) * inv_seg);
            const int ip = role < (ND >> 1) ? role : role - ND;                   // Frame.hpp:425-430
            const float2 hc = cis_neg_turns_f(fmaf(bt, (float)ip, at0));
            const float2 z = nmul(nmul(v[t], M->wseg[e]), hc);
            if (TAPS && ctap != nullptr) ctap[role] = z;
            M->sb[role] = (uint8_t)demap_n<0>(z, dk);
            if (ambiguous != nullptr) n_amb += demap_ambiguous(z, dk) ? 1 : 0;
        }
    }
    if (ambiguous != nullptr) {
        n_amb = (int)warp_sum((float)n_amb);
        if (lane == 0 && n_amb) atomicAdd(ambiguous, (unsigned long long)n_amb);
    }
    if (TAPS && taps.scal != nullptr && tid == 0) {
        float *sc = taps.scal + (size_t)frame * 48;
        if (s == 1) sc[4] = g;
        if (s < 16) { sc[16 + s] = (float)m; sc[32 + s] = theta; }
    }
    sync();
    // ---- pack: 8 consecutive symbols of `mod` bits = `mod` whole bytes, MSB first (modulation.cpp:90-125) ----
    {
        const int mod = P.mod_type;
        uint8_t *dst = out_bytes + (size_t)frame * P.bytes_per_frame + (size_t)(s - 1) * (size_t)(ND * mod / 8);
        for (int grp = tid; grp < ND / 8; grp += kBigThreads) {
            const uint2 raw = *reinterpret_cast<const uint2 *>(M->sb + 8 * grp);
            unsigned long long bits = 0;
#pragma unroll
            for (int e = 0; e < 8; e++) bits = (bits << mod) | (unsigned long long)(((e < 4 ? raw.x : raw.y) >> (8 * (e & 3))) & 0xffu);
            for (int bq = 0; bq < mod; bq++) dst[(size_t)grp * mod + bq] = (uint8_t)(bits >> (8 * (mod - 1 - bq)));
        }
    }
}

// Bridge: the any-size path's per-frame record (generic.cuh GenFrame) -> FrameScal, so that big_demod_kernel can run behind
// the any-size acquisition kernels (COFDM_BIG_ACQUIRE=0).  th0 carries theta_0 + m_0 whole (m0 = 0).
__global__ void big_bridge_kernel(const Params P, int n_frames, const GenFrame *__restrict__ gf, FrameScal *__restrict__ fscal, const RxTaps taps) {
    const int frame = blockIdx.x * blockDim.x + threadIdx.x;
    if (frame >= n_frames) return;
    const GenFrame &G = gf[frame];
    FrameScal f;
    f.kc = G.kc; f.m0 = 0;
    f.th0 = (float)((double)G.phit[0] + (double)P.fft_size * (double)G.kc / (double)P.pf_den);
    f.theta = G.theta; f.rot_theta = G.rot_theta; f.a = G.a; f.b = G.b;
    fscal[frame] = f;
    if (taps.scal != nullptr) {
        float *sc = taps.scal + (size_t)frame * 48;
        sc[0] = (float)((double)G.kc / (double)P.pf_den); sc[1] = (float)G.a; sc[2] = (float)G.b; sc[3] = G.theta; sc[5] = (float)G.kc;
    }
    if (taps.chan != nullptr) {
        const int ND = P.num_data_subc, half = ND / 2;
        for (int i = 0; i < ND; i++) taps.chan[(size_t)frame * ND + i] = cis_turns((G.b * (double)(i < half ? i : i - ND) + G.a) * 0.15915494309189533577);
    }
}

}  // namespace cofdmk
