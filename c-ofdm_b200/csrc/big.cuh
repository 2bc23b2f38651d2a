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
    float2 red[8];                   // per-warp partial sums (CP correlation)
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
COFDM_DEV void big_cluster_arrive(const BigCtx &) {}
COFDM_DEV void big_cluster_arrive_relaxed(const BigCtx &) {}
COFDM_DEV void big_cluster_wait(const BigCtx &) { __syncthreads(); }
template <class T> COFDM_DEV const T *big_remote(const BigCtx &c, const T *p, int r) {
    return reinterpret_cast<const T *>(reinterpret_cast<const unsigned char *>(p) + ((long)r - (long)c.rank) * (long)big_smem_bytes());
}
#else
COFDM_DEV BigCtx big_ctx() {
    extern __shared__ __align__(128) unsigned char big_smem_raw[];
    BigCtx c;
    unsigned rank, nrank, cid;                                  // 1-D clusters: rank in the cluster, its size, its index = the frame
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm("mov.u32 %0, %%cluster_nctarank;" : "=r"(nrank));
    asm("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
    c.tid = (int)threadIdx.x; c.rank = (int)rank; c.nrank = (int)nrank; c.frame = (int)cid;
    c.smem = big_smem_raw;
    return c;
}
COFDM_DEV void big_cta_sync(const BigCtx &) { __syncthreads(); }
COFDM_DEV void big_cluster_sync(const BigCtx &) { cooperative_groups::this_cluster().sync(); }
// split cluster barrier: arrive (release: this thread's shared-memory writes become visible to the cluster) ... independent
// work ... wait (acquire).  The relaxed form orders nothing: it only keeps a CTA's shared memory alive until every CTA of the
// cluster has stopped reading it.
COFDM_DEV void big_cluster_arrive(const BigCtx &) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
COFDM_DEV void big_cluster_arrive_relaxed(const BigCtx &) { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
COFDM_DEV void big_cluster_wait(const BigCtx &) { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
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

// The two per-frame phasors of the demod kernel, computed once by whoever writes the FrameScal (its eb1 / eb8 fields, which only
// the fft-512 kernels use otherwise):
//   eb1 = c_1 = exp(-j 2 pi Psi_1) exp(-j theta_pr), Psi_1 = 1.25 (theta_0 + m_0) mod 1: the constant phase of message symbol 1
//   eb8 = exp(-j b big_dstep): the equaliser's step from one row of a thread's bins to the next
COFDM_DEV void big_frame_phasors(FrameScal &f, const Params &P) {
    float acc = f.th0 * 1.25f;
    acc -= rintf(acc);
    const float psi1 = acc + (float)((5 * f.m0) & 3) * 0.25f;
    f.eb1 = nmul(cis_neg_turns_f(psi1), f.rot_theta);
    f.eb8 = cis_neg_turns_f((float)(f.b * 0.15915494309189533577) * (float)P.big_dstep);
}

// Forward FFT-4096 over the 256 threads of a CTA.  In: v[u] = x[j + 256 u] (thread j).  Out: v[t] = X[j + 256 t], unnormalised.
// E: kBigExchSlots float2 of shared memory nobody else touches; SYNC(): the CTA barrier.  w256 / w4096: global tables
// exp(-j 2 pi k / 256), exp(-j 2 pi k / 4096).  Stockham autosort passes (radix 16, Ns = 1, 16, 256):
//   pass p: thread j, k = j mod Ns: in[j + 256 t] * W_{16 Ns}^{k t} -> DFT-16 -> out[(j - k) 16 + k + Ns t'].
// ROT: the inputs of thread j' still lack a factor F(j') = pq_base * step16^(j' >> 4) ... supplied by the caller as
// (f0, fstep): register t of the pass-1 reader j comes from writer (j >> 4) + 16 t, whose factor is f0 * fstep^t; it rides on
// the pass-1 twiddle chain (one sequential product per register instead of a power table + a separate rotation).
template <bool ROT, class SYNC>
COFDM_DEV void cta_fft4096(float2 (&v)[16], float2 *E, const float2 *__restrict__ w4096, int j, SYNC sync,
                           float2 f0 = make_float2(1.f, 0.f), float2 fstep = make_float2(1.f, 0.f)) {
    ndft16(v);
#pragma unroll
    for (int t = 0; t < 16; t++) E[17 * j + t] = v[t];                       // y[16 j + t] at slot i + (i >> 4)
    sync();
    {
        const float2 *r = E + j + (j >> 4);
#pragma unroll
        for (int t = 0; t < 16; t++) v[t] = r[272 * t];                      // y[j + 256 t]
        if (ROT) {
            const float2 stp = nmul(__ldg(w4096 + 16 * (j & 15)), fstep);    // W256^{j mod 16} * fstep
            float2 g = f0;
            v[0] = nmul(v[0], g);
#pragma unroll
            for (int t = 1; t < 16; t++) { g = nmul(g, stp); v[t] = nmul(v[t], g); }
        } else {
            float2 w[16];
            npowers15(__ldg(w4096 + 16 * (j & 15)), w);                      // W256^{(j mod 16) t}
#pragma unroll
            for (int t = 1; t < 16; t++) v[t] = nmul(v[t], w[t]);
        }
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
// MOD: modulation order the instance is specialised for (6), or 0 = any (read from the configuration)
// LAY: the instance is compiled for the row layout kBigLayData / kBigLayUsed (P.big_lay; the production geometry's 1920 + 128
//      sub-carriers): unused rows of the last FFT pass are never computed, the equaliser runs branch-free from P.big_eq
constexpr unsigned kBigLayData = 0xF00Fu, kBigLayUsed = 0xF01Fu;
template <int FMT, bool USE_TMA, bool TAPS, int MOD, bool LAY = false>
__global__ void __launch_bounds__(kBigThreads, 4)
big_demod_kernel(const Params P, const void *__restrict__ samples, long long frame_stride /*samples*/, int n_frames,
                 uint8_t *__restrict__ out_bytes, unsigned long long *__restrict__ ambiguous, const RxTaps taps,
                 const FrameScal *__restrict__ fscal) {
    static_assert(!(LAY && TAPS), "the layout-specialised instance has no taps");
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
    //      R^u is applied here; P rides on the pass-1 twiddles of the thread that reads this thread's pass-0 outputs: reader j gets
    //      register t from writer (j >> 4) + 16 t, i.e. the factor P((j >> 4)) * P-step^t with P-step = exp(-j 2 pi beta 16 / 4096)
    {
        const float2 R = big_phasor(theta, m, 256);
        float2 g = R;
        v[1] = nmul(v[1], g);
#pragma unroll
        for (int u = 2; u < 16; u++) { g = nmul(g, R); v[u] = nmul(v[u], g); }
    }
    cta_fft4096<true>(v, reinterpret_cast<float2 *>(stage), P.tw_fft, tid, sync, big_phasor(theta, m, kBigCP + (tid >> 4)), big_phasor(theta, m, 16));
    // now v[t] = X[tid + 256 t] of the rotated symbol (constant phase Psi_s still on it: it cancels against the segment pilot)

    // ---- the thread's 16 roles (bins tid + 256 t), two 128-bit loads: int16 each, >= 0 data index, -1 null, -2 - p pilot p ----
    const uint4 ra = __ldg(&P.big_roles[2 * tid]), rb = __ldg(&P.big_roles[2 * tid + 1]);
    const unsigned rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#define COFDM_ROLE(t) ((int)(short)((rw[(t) >> 1] >> (16 * ((t) & 1))) & 0xffffu))
    const unsigned tmask = LAY ? kBigLayUsed : (unsigned)P.big_tmask;   // rows t in which ANY thread holds a used bin (uniform)
    // ---- pilots (Frame.cpp:76-80): the few threads whose bins are pilots put them side by side; warp 0 then sums |pilot| ----
    if (((ra.x | ra.y | ra.z | ra.w | rb.x | rb.y | rb.z | rb.w) & 0x80008000u) != 0u) {
#pragma unroll
        for (int t = 0; t < 16; t++) {
            if (!((tmask >> t) & 1u)) continue;
            const int role = COFDM_ROLE(t);
            if (role <= -2) M->pil[-2 - role] = v[t];
        }
    }
    sync();                                                       // pil[] complete
    if (warp == 0) {
        float pm = 0.f;
        for (int q = lane; q < P.num_pilot_subc; q += 32) {
            const float n2 = cnorm2(M->pil[q]);
            pm = fmaf(n2, rsqrtf(fmaxf(n2, 1e-30f)), pm);          // |pilot| (2 ulp: it enters g, a sum of 1024 terms)
        }
        pm = warp_sum(pm);
        if (lane == 0) M->pabs = pm;                              // (published by this warp's own arrive.release)
    }
    // this symbol's pilots and pabs are published: warp 0 arrives with release semantics -- it wrote pabs, and the other warps'
    // pilots reached it through the CTA barrier above (release is cumulative) -- so the other warps need no fence of their own
    if (warp == 0) big_cluster_arrive(cx); else big_cluster_arrive_relaxed(cx);
    const int ND = P.num_data_subc, nw = cx.nrank;
    const float bt = (float)(fs.b * 0.15915494309189533577), at0 = (float)(fs.a * 0.15915494309189533577 - rint(fs.a * 0.15915494309189533577));
    const int dstep = P.big_dstep;
    const float2 hstep = fs.eb8, c1 = fs.eb1;                     // exp(-j b dstep), c_1 (big_frame_phasors)
    uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
    if (LAY) {
        // The row layout is known: the thread's data sit in rows 0..3 and 12..15 and its abscissa advances by big_dstep inside each
        // group; one table word per row says where the symbol goes (a dump slot when the bin holds no data) and which segment
        // coefficient it takes.  The equaliser's own factor exp(-j (b i' + a)) needs nothing from the other symbols: it is applied
        // here, while the cluster barrier completes -- two phasors evaluated, six by recurrence, no branches.
        qa = __ldg(&P.big_eq[3 * tid]); qb = __ldg(&P.big_eq[3 * tid + 1]);
        const unsigned hw = __ldg(&P.big_eq[3 * tid + 2].x);
        float2 hc = cis_neg_turns_f(fmaf(bt, (float)(short)(hw & 0xffffu), at0));
        v[0] = nmul(v[0], hc);  hc = nmul(hc, hstep); v[1] = nmul(v[1], hc);
        hc = nmul(hc, hstep);   v[2] = nmul(v[2], hc);  hc = nmul(hc, hstep); v[3] = nmul(v[3], hc);
        hc = cis_neg_turns_f(fmaf(bt, (float)((int)hw >> 16), at0));
        v[12] = nmul(v[12], hc); hc = nmul(hc, hstep); v[13] = nmul(v[13], hc);
        hc = nmul(hc, hstep);    v[14] = nmul(v[14], hc); hc = nmul(hc, hstep); v[15] = nmul(v[15], hc);
    }
    big_cluster_wait(cx);                                         // pilots and pabs of every symbol of the frame are visible

    // ---- frame-wide: g (Frame.cpp:76-80) and the segment coefficients W[e] = P_1[e] conj(P_s[e]) / (|P_s[e]|^2 g) c_1 (Frame.cpp:89-92) ----
    //      (lane r fetches symbol r's sum; a fixed tree over the 8 lanes, the same in every CTA of the frame)
    float g = lane < cx.nrank ? *big_remote(cx, &M->pabs, lane) : 0.f;
    g += __shfl_xor_sync(0xffffffffu, g, 4);
    g += __shfl_xor_sync(0xffffffffu, g, 2);
    g += __shfl_xor_sync(0xffffffffu, g, 1);
    g = __shfl_sync(0xffffffffu, g, 0) * P.inv_pilot_norm;
    if (tid < P.num_pilot_subc) {
        const float2 p1 = big_remote(cx, M->pil, 0)[tid], ps = M->pil[tid];
        const float2 w = nscale(nmulc(p1, ps), __fdividef(1.0f, cnorm2(ps) * g));
        M->wseg[tid] = nmul(w, c1);
    }
    big_cluster_arrive_relaxed(cx);                               // this CTA reads no remote shared memory from here on (waited for at the end)
    sync();                                                       // wseg visible

    // ---- equaliser exp(-j (b i' + a)), i' = i (i < ND / 2) or i - ND (rx.cpp:214-216, Frame.hpp:425-430), segment correction and
    //      hard demap (modulation.cpp:53-87) straight from the registers.  A thread's data indices advance by a constant step from
    //      row to row (big_dstep, e.g. 240 = 256 * 15 / 16): the phasor of the next row is the previous one times exp(-j b step);
    //      any other step is evaluated directly. ----
    const DemapK dk = make_demapk(P.mod_type);
    const float inv_seg = 1.0f / (float)P.seg_size;
    float2 *ctap = (TAPS && taps.constell != nullptr) ? taps.constell + ((size_t)frame * nw + (s - 1)) * ND : nullptr;
    int n_amb = 0;
    if (LAY) {
        // segment correction and hard demap of the 8 rows (equalised above)
        const unsigned char *wsb = reinterpret_cast<const unsigned char *>(M->wseg);
#define COFDM_BIG_Z(T, W) nmul(v[T], *reinterpret_cast<const float2 *>(wsb + ((W) >> 16)))
#define COFDM_BIG_ROWS(F) { F(0, qa.x); F(1, qa.y); F(2, qa.z); F(3, qa.w); F(12, qb.x); F(13, qb.y); F(14, qb.z); F(15, qb.w); }
#define COFDM_BIG_EQ(T, W) M->sb[(W) & 0xffffu] = (uint8_t)demap_n<MOD>(COFDM_BIG_Z(T, W), dk)
        COFDM_BIG_ROWS(COFDM_BIG_EQ)
        if (ambiguous != nullptr) {
            // optional count of boundary-ambiguous decisions: the points are recomputed, off the fast path
#define COFDM_BIG_AMB(T, W) n_amb += (((W) & 0xffffu) < (unsigned)ND && demap_ambiguous(COFDM_BIG_Z(T, W), dk)) ? 1 : 0
            COFDM_BIG_ROWS(COFDM_BIG_AMB)
#undef COFDM_BIG_AMB
        }
#undef COFDM_BIG_EQ
#undef COFDM_BIG_ROWS
#undef COFDM_BIG_Z
    } else {
        float2 hc = make_float2(1.f, 0.f);
        int prev_ip = -0x40000000;
#pragma unroll
        for (int t = 0; t < 16; t++) {
            if (!((tmask >> t) & 1u)) continue;
            const int role = COFDM_ROLE(t);
            if (role >= 0) {
                const int ip = role < (ND >> 1) ? role : role - ND;
                if (ip - prev_ip == dstep) hc = nmul(hc, hstep);
                else hc = cis_neg_turns_f(fmaf(bt, (float)ip, at0));
                prev_ip = ip;
                const int e = (int)(((float)role + 0.5f) * inv_seg);
                const float2 z = nmul(nmul(v[t], hc), M->wseg[e]);
                if (TAPS && ctap != nullptr) ctap[role] = z;
                M->sb[role] = (uint8_t)demap_n<MOD>(z, dk);
                if (ambiguous != nullptr) n_amb += demap_ambiguous(z, dk) ? 1 : 0;
            }
        }
    }
#undef COFDM_ROLE
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
        const int mod = MOD ? MOD : P.mod_type;
        uint8_t *dst = out_bytes + (size_t)frame * P.bytes_per_frame + (size_t)(s - 1) * (size_t)(ND * mod / 8);
        for (int grp = tid; grp < ND / 8; grp += kBigThreads) {
            const uint2 raw = *reinterpret_cast<const uint2 *>(M->sb + 8 * grp);
            if (MOD == 6 && (reinterpret_cast<uintptr_t>(dst) & 1) == 0) {
                // 64-QAM: 8 symbols of 6 bits = 3 big-endian 16-bit words
                const unsigned s0 = raw.x & 63u, s1 = (raw.x >> 8) & 63u, s2 = (raw.x >> 16) & 63u, s3 = raw.x >> 24;
                const unsigned s4 = raw.y & 63u, s5 = (raw.y >> 8) & 63u, s6 = (raw.y >> 16) & 63u, s7 = raw.y >> 24;
                const unsigned w0 = (s0 << 10) | (s1 << 4) | (s2 >> 2), w1 = ((s2 & 3u) << 14) | (s3 << 8) | (s4 << 2) | (s5 >> 4),
                               w2 = ((s5 & 15u) << 12) | (s6 << 6) | s7;
                unsigned short *d16 = reinterpret_cast<unsigned short *>(dst + (size_t)grp * 6);
                d16[0] = (unsigned short)((w0 >> 8) | ((w0 & 0xffu) << 8)); d16[1] = (unsigned short)((w1 >> 8) | ((w1 & 0xffu) << 8));
                d16[2] = (unsigned short)((w2 >> 8) | ((w2 & 0xffu) << 8));
            } else {
                unsigned long long bits = 0;
#pragma unroll
                for (int e = 0; e < 8; e++) bits = (bits << mod) | (unsigned long long)(((e < 4 ? raw.x : raw.y) >> (8 * (e & 3))) & 0xffu);
                for (int bq = 0; bq < mod; bq++) dst[(size_t)grp * mod + bq] = (uint8_t)(bits >> (8 * (mod - 1 - bq)));
            }
        }
    }
    big_cluster_wait(cx);                                         // nobody reads this CTA's shared memory any more
}

// forward 20-point DFT in place, natural layout: X[k], X[k + 10] = E[k] +- W20^k O[k], E / O = DFT-10 of the even / odd inputs
COFDM_DEV void ndft20(float2 (&v)[20]) {
    float2 e[10], o[10];
#pragma unroll
    for (int q = 0; q < 10; q++) { e[q] = v[2 * q]; o[q] = v[2 * q + 1]; }
    ndft10(e);
    ndft10(o);
    o[1] = nmul(o[1], make_float2(0.95105651629515353118f, -0.30901699437494739575f));
    o[2] = nmul(o[2], make_float2(0.80901699437494745126f, -0.58778525229247313710f));
    o[3] = nmul(o[3], make_float2(0.58778525229247313710f, -0.80901699437494745126f));
    o[4] = nmul(o[4], make_float2(0.30901699437494745126f, -0.95105651629515353118f));
    o[5] = make_float2(o[5].y, -o[5].x);                                      // W20^5 = -j
    o[6] = nmul(o[6], make_float2(-0.30901699437494734024f, -0.95105651629515364220f));
    o[7] = nmul(o[7], make_float2(-0.58778525229247302608f, -0.80901699437494745126f));
    o[8] = nmul(o[8], make_float2(-0.80901699437494734024f, -0.58778525229247324813f));
    o[9] = nmul(o[9], make_float2(-0.95105651629515353118f, -0.30901699437494750677f));
#pragma unroll
    for (int k = 0; k < 10; k++) { v[k] = nadd(e[k], o[k]); v[k + 10] = nsub(e[k], o[k]); }
}

// ================================================================================================================
// big_acquire_kernel: the preamble of one frame per CTA (256 threads) -> FrameScal.
//   coarse CFO   pilot_freq_sinh (Frame.hpp:285-337): 5120-point spectrum of the received preamble, CP included, as radix
//                20 x 16 x 16 (Stockham; first exchange padded 20 -> 21 so that every access is conflict-free; the radix-16
//                passes have 320 butterflies: 256 + 64), |X|^2, arg-max in the pilot windows -> kc
//   fine CFO     cp_freq_sinh (:238-263) on the preamble: CP correlation -> theta_0, m_0; rotation; FFT-4096
//   phase lock   pr_phase_sinh (:265-274): theta = arg sum conj(ref) y over the 5120 rotated samples
//   channel fit  chan_char_lq (:389-434): ND / 2 phases, the reference's one-step unwrap, the (bug-compatible) line
// Shared memory: Y (43008 B: first coarse exchange -> |X|^2 -> FFT-4096 exchange) and Z (40960 B: staged preamble -> second
// coarse exchange -> staged preamble again (second bulk copy, an L2 hit) -> the phases).
// ================================================================================================================
constexpr int kBigAcqY = (5120 + 256) * 8, kBigAcqZ = 5120 * 8;
struct alignas(16) BigAcqShared {
    float2 red[8];
    int ksum;
    int anyjump;                     // some neighbouring phases differ by more than pi: the unwrap's slow path
    int pad_[2];
    uint64_t mbar[2];
    double dsum[8][2];               // per-warp partial sums of the channel-line fit
};
COFDM_HD constexpr size_t big_acquire_smem_bytes() { return (size_t)kBigAcqY + kBigAcqZ + sizeof(BigAcqShared); }

// one radix-16 Stockham butterfly of a 5120-point transform: bf in [0, 320), Ns = 20 (PASS 1) or 320 (PASS 2)
template <int PASS>
COFDM_DEV void big_coarse_bf(const float2 *in, float2 *out, float *mag, const float2 *__restrict__ w5120, int bf) {
    float2 t[16], w[16];
    if (PASS == 1) {
        const int k = bf % 20;
        const float2 *r = in + bf + bf / 20;                                 // y[bf + 320 q] at slot i + i / 20
#pragma unroll
        for (int q = 0; q < 16; q++) t[q] = r[336 * q];
        npowers15(__ldg(w5120 + 16 * k), w);                                 // W320^{k q}
#pragma unroll
        for (int q = 1; q < 16; q++) t[q] = nmul(t[q], w[q]);
        ndft16(t);
        float2 *o = out + (bf - k) * 16 + k;
#pragma unroll
        for (int q = 0; q < 16; q++) o[20 * q] = t[q];
    } else {
#pragma unroll
        for (int q = 0; q < 16; q++) t[q] = in[bf + 320 * q];
        npowers15(__ldg(w5120 + bf), w);                                     // W5120^{bf q}
#pragma unroll
        for (int q = 1; q < 16; q++) t[q] = nmul(t[q], w[q]);
        ndft16(t);
#pragma unroll
        for (int q = 0; q < 16; q++) { const float2 sq = p_mul(t[q], t[q]); mag[bf + 320 * q] = sq.x + sq.y; }
    }
}

template <int FMT, bool USE_TMA, bool TAPS>
__global__ void __launch_bounds__(kBigThreads, 2)
big_acquire_kernel(const Params P, const void *__restrict__ samples, long long frame_stride /*samples*/, int n_frames,
                   const RxTaps taps, FrameScal *__restrict__ fscal) {
    COFDM_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.x;
    if (frame >= n_frames) return;
    float2 *Y = reinterpret_cast<float2 *>(smem_raw);
    unsigned char *Zb = reinterpret_cast<unsigned char *>(smem_raw) + kBigAcqY;
    float2 *Z = reinterpret_cast<float2 *>(Zb);
    BigAcqShared *M = reinterpret_cast<BigAcqShared *>(Zb + kBigAcqZ);
    const size_t sample_bytes = (FMT == kCI16) ? 4 : 8;
    const char *src = reinterpret_cast<const char *>(samples) + (size_t)frame * (size_t)frame_stride * sample_bytes;
    auto stage_issue = [&](int which) {
        if (USE_TMA) {
            if (tid == 0) {
                mbar_init(&M->mbar[which], 1);
                mbar_fence_init();
                mbar_arrive_expect_tx(&M->mbar[which], kBigL * (unsigned)sample_bytes);
                tma_load_1d(Zb, src, kBigL * (unsigned)sample_bytes, &M->mbar[which]);
            }
        } else {
            if (FMT == kCI16) for (int i = tid; i < kBigL; i += kBigThreads) reinterpret_cast<unsigned *>(Zb)[i] = __ldg(reinterpret_cast<const unsigned *>(src) + i);
            else for (int i = tid; i < kBigL; i += kBigThreads) Z[i] = __ldg(reinterpret_cast<const float2 *>(src) + i);
        }
    };
    stage_issue(0);
    if (tid == 0) { M->ksum = 0; M->anyjump = 0; }
    __syncthreads();
    if (USE_TMA) mbar_wait(&M->mbar[0], 0);

    // ================= coarse CFO + the preamble's CP correlation =================
    float2 cc;
    {
        float2 t[20];
#pragma unroll
        for (int u = 0; u < 20; u++) t[u] = staged_at<FMT>(Zb, tid + 256 * u);
        // CP sample j = tid + 256 c pairs with sample j + 4096 (u = 16 + c)  (Frame.hpp:251-253)
        cc = nmac_conj(nmac_conj(make_float2(0.f, 0.f), t[0], t[16]), t[1], t[17]);
        cc = nmac_conj(nmac_conj(cc, t[2], t[18]), t[3], t[19]);
        ndft20(t);
#pragma unroll
        for (int u = 0; u < 20; u++) Y[21 * tid + u] = t[u];                  // y[20 j + u] at slot i + i / 20
    }
    cc = warp_sum(cc);
    if (lane == 0) M->red[warp] = cc;
    __syncthreads();
    big_coarse_bf<1>(Y, Z, nullptr, P.tw_pf, tid);
    if (tid < 64) big_coarse_bf<1>(Y, Z, nullptr, P.tw_pf, tid + 256);
    __syncthreads();
    float *mag = reinterpret_cast<float *>(Y);
    big_coarse_bf<2>(Z, nullptr, mag, P.tw_pf, tid);
    if (tid < 64) big_coarse_bf<2>(Z, nullptr, mag, P.tw_pf, tid + 256);
    __syncthreads();
    stage_issue(1);                                                           // the raw preamble again (Z is free; the copy hits L2)
    {
        // arg-max of |spectrum| in the pilot windows, first maximum wins (Frame.hpp:311-331); window np/2 (DC) is skipped.
        // The windows are short (pf_pilot_w = 20 bins here): TWO THREADS per window scan one half each, the lower half wins ties.
        const int np = P.num_pilot_subc, half = P.pf_size / 2;
        const int wi = tid >> 1, part = tid & 1;
        float best = -1.0f;
        int besti = 0;
        if (wi < np) {
            const int win = wi < np / 2 ? wi : wi + 1;
            int lo = P.pf_border0 + win * P.pf_pilot_w;
            const int hi = lo + P.pf_pilot_w;
            if (win == 0 && lo < 0) lo = 0;
            const int mid = lo + ((hi - lo + 1) >> 1);
            const int k0 = part ? mid : lo, k1 = part ? hi : mid;
            for (int ks = k0; ks < k1; ks++) {
                const float mv = mag[ks < half ? ks + half : ks - half];
                if (mv > best) { best = mv; besti = ks; }
            }
        }
        const float obest = __shfl_xor_sync(0xffffffffu, best, 1);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, 1);
        if (obest > best) besti = oi;
        const int acc = __reduce_add_sync(0xffffffffu, (part == 0 && wi < np) ? besti : 0);
        if (lane == 0) atomicAdd(&M->ksum, acc);
    }
    __syncthreads();
    if (USE_TMA) mbar_wait(&M->mbar[1], 0);
    const int kc = M->ksum - P.num_pilot_subc * (P.pf_size / 2);              // shift = kc / pf_den (Frame.hpp:332-334)
    cc = M->red[0];
#pragma unroll
    for (int w = 1; w < 8; w++) cc = nadd(cc, M->red[w]);
    const float theta0 = fast_atan2_turns(cc.y, cc.x);
    const int m0 = (int)ceilf(-(theta0 - (float)kc * P.pf_binsN) - 0.5f);

    // ================= rotation, pr_phase_sinh partial sums, FFT-4096 =================
    float2 v[16];
    float2 z = make_float2(0.f, 0.f);
    {
        float2 rp[16];
        npowers15(big_phasor(theta0, m0, 256), rp);
        const float2 pl = big_phasor(theta0, m0, kBigCP + tid);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            // CP sample j = tid + 256 c: exp(-j 2 pi beta j / 4096) = P(tid) conj(R^(4 - c))
            const float2 y = nmul(nmulc(staged_at<FMT>(Zb, tid + 256 * c), rp[4 - c]), pl);
            z = nmac_conj(z, __ldg(&P.preamble_td[tid + 256 * c]), y);
        }
#pragma unroll
        for (int u = 0; u < 16; u++) {
            v[u] = nmul(staged_at<FMT>(Zb, kBigCP + tid + 256 * u), pl);
            if (u) v[u] = nmul(v[u], rp[u]);
            z = nmac_conj(z, __ldg(&P.preamble_td[kBigCP + tid + 256 * u]), v[u]);
        }
    }
    z = warp_sum(z);
    // the roles of the thread's 16 bins (needed after the transform; fetched here so that nothing waits for them there)
    const uint4 ra = __ldg(&P.big_roles[2 * tid]), rb = __ldg(&P.big_roles[2 * tid + 1]);
    __syncthreads();                                                          // red[] and mag have been read by everybody
    if (lane == 0) M->red[warp] = z;
    auto sync = [&]() { __syncthreads(); };
    cta_fft4096<false>(v, Y, P.tw_fft, tid, sync);                            // (its first barrier also publishes red[])
    z = M->red[0];
#pragma unroll
    for (int w = 1; w < 8; w++) z = nadd(z, M->red[w]);
    const float inv = rsqrtf(fmaxf(cnorm2(z), 1e-30f));
    const float2 rot = make_float2(z.x * inv, -z.y * inv);                    // exp(-j theta)
    const float theta = TAPS ? atan2f(z.y, z.x) : 0.f;

    // ================= chan_char_lq: phase[i] = arg(pr[i] / mod_preamble[i]), i < ND / 2 (Frame.hpp:403-405) =================
    const float TWO_PI_F = 6.28318530717958647692f, PI_F = 3.14159265358979323846f;
    const int nph = P.num_data_subc / 2;
    float *phs = reinterpret_cast<float *>(Zb);                              // the staged preamble has been consumed (barriers of the FFT)
    const unsigned rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
    const unsigned phmask = (unsigned)P.big_phmask;                          // rows in which ANY thread holds one of these bins (uniform)
#pragma unroll
    for (int t = 0; t < 16; t++) {
        if (!((phmask >> t) & 1u)) continue;
        const int role = (int)(short)((rw[t >> 1] >> (16 * (t & 1))) & 0xffffu);
        if (role >= 0 && role < nph) {
            const float2 d = nmulc(nmul(v[t], rot), __ldg(&P.mod_preamble[role]));
            phs[role] = fast_atan2_turns(d.y, d.x) * TWO_PI_F;
        }
    }
    __syncthreads();
    // The one-step unwrap (Frame.hpp:407-414) changes nothing unless two neighbouring phases differ by more than pi.  All 256
    // threads look for such a pair in their own stretch and add up the sums of Frame.hpp:416-421 as if there were none -- the
    // common case, finished here in parallel; otherwise warp 0 redoes the sums with the unwrap's 3-state chain below.
    double tsy, tsxy;
    {
        const int E8 = (nph + kBigThreads - 1) / kBigThreads, a0 = tid * E8, a1 = min(nph, a0 + E8);
        float pv = a0 > 0 && a0 < nph ? phs[a0 - 1] : 0.f;
        bool jump = false;
        double ssy = 0.0, ssxy = 0.0;
        for (int i = a0; i < a1; i++) {
            const float p = phs[i];
            if (i > 0 && fabsf(p - pv) > PI_F) jump = true;
            pv = p;
            ssy += (double)p;
            ssxy += (double)p * (double)i;
        }
        if (jump) M->anyjump = 1;
        ssy = warp_sum(ssy); ssxy = warp_sum(ssxy);
        if (lane == 0) { M->dsum[warp][0] = ssy; M->dsum[warp][1] = ssxy; }
    }
    __syncthreads();
    if (warp != 0) return;
    if (M->anyjump == 0) {
        tsy = warp_sum(lane < kBigThreads / 32 ? M->dsum[lane][0] : 0.0);
        tsxy = warp_sum(lane < kBigThreads / 32 ? M->dsum[lane][1] : 0.0);
    } else {
        // by one warp: lane l owns elements [l E, (l + 1) E).  The adjustment is a 3-state chain (state = multiple of 2 pi
        // carried by the previous element): every lane maps the three possible entry states of its range to exit states,
        // the maps are composed across the lanes, then each lane replays its range from its true entry state.
        const int E = (nph + 31) / 32, i0 = lane * E, i1 = min(nph, i0 + E);
        const float prev_raw = i0 > 0 && i0 < nph ? phs[i0 - 1] : 0.f;
        int cstart = 0;                                                       // state entering this lane's range
        {
            unsigned map = 0;
            for (int cin = 0; cin < 3; cin++) {
                int cs = cin - 1;
                float pv = prev_raw;
                for (int i = i0; i < i1; i++) {
                    const float p = phs[i];
                    if (i == 0) { cs = 0; pv = p; continue; }
                    const float dlt = p - (pv + (float)cs * TWO_PI_F);
                    cs = dlt > PI_F ? -1 : (dlt < -PI_F ? 1 : 0);
                    pv = p;
                }
                map |= (unsigned)(cs + 1) << (2 * cin);
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
            cstart = lane == 0 ? 0 : (int)((before >> 2) & 3u) - 1;        // the chain starts in state 0 (entry [1] of the composed map)
        }
        double ssy = 0.0, ssxy = 0.0;
        {
            int cs = cstart;
            float pv = prev_raw;
            for (int i = i0; i < i1; i++) {
                const float p = phs[i];
                float val = p;
                if (i > 0) {
                    const float dlt = p - (pv + (float)cs * TWO_PI_F);
                    cs = dlt > PI_F ? -1 : (dlt < -PI_F ? 1 : 0);
                    val = p + (float)cs * TWO_PI_F;
                } else {
                    cs = 0;
                }
                pv = p;
                ssy += (double)val;
                ssxy += (double)val * (double)i;
            }
        }
        tsy = warp_sum(ssy); tsxy = warp_sum(ssxy);
    }
    const double n = (double)nph, sx1 = n * (n - 1.0) / 2.0, sx2 = (n - 1.0) * n * (2.0 * n - 1.0) / 6.0;
    const double lb = (tsxy - sx1 * tsy) / (sx2 - sx1 * sx1);               // Frame.hpp:422 (sums, not means)
    const double la = tsy - lb * sx1;                                        // Frame.hpp:423
    if (lane == 0) {
        FrameScal f;
        f.kc = kc; f.m0 = m0; f.th0 = theta0; f.theta = theta; f.rot_theta = rot; f.a = la; f.b = lb;
        big_frame_phasors(f, P);
        fscal[frame] = f;
    }
    if (TAPS) {
        if (taps.scal != nullptr && lane == 0) {
            float *sc = taps.scal + (size_t)frame * 48;
            sc[0] = (float)((double)kc / (double)P.pf_den); sc[1] = (float)la; sc[2] = (float)lb; sc[3] = theta;
            sc[5] = (float)kc; sc[6] = 0.f; sc[7] = 0.f; sc[16] = (float)m0; sc[32] = theta0;
        }
        if (taps.chan != nullptr) {
            const int ND = P.num_data_subc;
            for (int i = lane; i < ND; i += 32)
                taps.chan[(size_t)frame * ND + i] = cis_turns((lb * (double)(i < nph ? i : i - ND) + la) * 0.15915494309189533577);
        }
    }
}

// ================================================================================================================
// big_tx_kernel: FRAME_FORM::write + get / get_int16 (Frame.cpp:185-198, 54-70, 244-256) for the fft-4096 geometry.
// grid (num_symb + 1, n_frames): CTA (s, f) builds message symbol s of frame f -- the thread's 16 grid points (bins j + 256 u:
// data point, pilot or null, from the descriptor table P.big_txd) conjugated, conj(FFT(conj G)) = the backward transform, / sqrt(4096), coalesced stores of body and cyclic prefix straight from the
// registers (thread j holds samples j + 256 t: a warp stores 32 consecutive samples); CTA (num_symb, f) copies the constant
// sync tone + preamble.
// ================================================================================================================
// shared memory behind the exchange: the point table (up to 64-QAM: 66 entries x 16 lane slots, so that lane l of a half-warp
// always reads bank pair l & 15 -- no conflicts whatever the data; 256-QAM: 258 plain entries), the staged payload, the mbarrier
constexpr int kBigTxTabBytes = 66 * 16 * 8, kBigTxPayMax = kBigMaxData + 16;
COFDM_HD constexpr size_t big_tx_smem_bytes() { return (size_t)kBigExchSlots * 8 + kBigTxTabBytes + kBigTxPayMax + 16; }

template <int FMT>
COFDM_DEV void big_store(void *frame_out, long long idx, float2 v, float mult) {
    if (FMT == kCI16) {
        // Frame.cpp:252: int16(trunc(re*mult)), int16(trunc(im*mult))
        const unsigned pk = ((unsigned)(unsigned short)(short)__float2int_rz(v.x * mult)) | ((unsigned)(unsigned short)(short)__float2int_rz(v.y * mult) << 16);
        reinterpret_cast<unsigned *>(frame_out)[idx] = pk;
    } else {
        reinterpret_cast<float2 *>(frame_out)[idx] = v;
    }
}

// MOD: modulation order the instance is specialised for (6), or 0 = any; LAY: compiled for the row layout kBigLayUsed (rows
// 5..11 of the grid are empty: the first FFT pass never touches them).
// P.big_txd: per thread 16 words, one per row: [15:0] byte of the symbol's payload where the sub-carrier's bits start,
// [18:16] bit offset in that byte, [25:24] 1 = pilot, 2 = null.  The bytes come straight from global memory (two byte loads per
// row, neighbouring threads read neighbouring bytes; all of a thread's loads are in flight together) -- or, when the symbol's
// bytes are 16-byte aligned (the production geometry: 1440 bytes per symbol), from a copy one bulk load (TMA) put in shared
// memory; the points from a small table in shared memory: conj(constellation), then the pilot and zero.
template <int FMT, int MOD = 0, bool LAY = false>
__global__ void __launch_bounds__(kBigThreads, 4)
big_tx_kernel(const Params P, const uint8_t *__restrict__ payload, int n_frames, void *__restrict__ frames) {
    COFDM_DYN_SMEM(smem_raw);
    const int s = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x;
    if (frame >= n_frames) return;
    const size_t sb = (FMT == kCI16) ? 4 : 8;
    char *fout = reinterpret_cast<char *>(frames) + (size_t)frame * P.frame_len * sb;
    if (s == P.num_symb) {
        for (int i = tid; i < P.t2sin_size + P.pf_size; i += kBigThreads)
            big_store<FMT>(fout, i, i < P.t2sin_size ? __ldg(&P.t2_tone[i]) : __ldg(&P.preamble_td[i - P.t2sin_size]), P.mult);
        return;
    }
    float2 *E = reinterpret_cast<float2 *>(smem_raw);
    float2 *ct = reinterpret_cast<float2 *>(smem_raw + (size_t)kBigExchSlots * 8);
    uint8_t *pl = reinterpret_cast<uint8_t *>(smem_raw) + (size_t)kBigExchSlots * 8 + kBigTxTabBytes;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(pl + kBigTxPayMax);
    const int mod = MOD ? MOD : P.mod_type, npts = 1 << mod, ND = P.num_data_subc, sym_bytes = ND * mod / 8;
    const uint8_t *src = payload + (size_t)frame * P.bytes_per_frame + (size_t)s * sym_bytes;
    const bool staged = ((reinterpret_cast<uintptr_t>(src) | (unsigned)sym_bytes) & 15u) == 0u && sym_bytes <= kBigTxPayMax;
    if (staged && tid == 0) {
        mbar_init(mbar, 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(mbar, (unsigned)sym_bytes);
        tma_load_1d(pl, src, (unsigned)sym_bytes, mbar);
    }
    // the point table; up to 64-QAM every entry is repeated for the 16 lane slots (entry e, slot q at e * 16 + q)
    const bool skew = mod <= 6;
    if (skew) {
        // one load per thread: entry tid / 4, lane slots 4 (tid % 4) .. + 3; the pilot and zero by the first warp
        const int e = tid >> 2;
        if (e < npts) {
            float2 c = __ldg(&P.constell[e]);
            c.y = -c.y;                                                            // conjugated: the backward transform is conj(FFT(conj G))
            float4 *d = reinterpret_cast<float4 *>(ct + e * 16 + (tid & 3) * 4);
            d[0] = make_float4(c.x, c.y, c.x, c.y); d[1] = d[0];
        }
        if (tid < 32) ct[(npts + (tid >> 4)) * 16 + (tid & 15)] = make_float2((tid >> 4) ? 0.f : P.pilot_ampl, 0.f);   // Frame.cpp:55-57
    } else {
        for (int i = tid; i < npts + 2; i += kBigThreads) {
            float2 c = make_float2(0.f, 0.f);                                      // Frame.cpp:55
            if (i < npts) { c = __ldg(&P.constell[i]); c.y = -c.y; }
            else if (i == npts) c = make_float2(P.pilot_ampl, 0.f);                // Frame.cpp:56-57
            ct[i] = c;
        }
    }
    const unsigned tmask = LAY ? kBigLayUsed : (unsigned)P.big_tmask;
    const uint4 da = __ldg(&P.big_txd[4 * tid]), db = __ldg(&P.big_txd[4 * tid + 1]), dc = __ldg(&P.big_txd[4 * tid + 2]), dd = __ldg(&P.big_txd[4 * tid + 3]);
    const unsigned dw[16] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w, dc.x, dc.y, dc.z, dc.w, dd.x, dd.y, dd.z, dd.w};
    unsigned wb[16];
    // (the second byte is read at a clamped position: when the bits end inside the first byte it is shifted out anyway)
    if (!staged) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            wb[u] = 0u;
            if (!((tmask >> u) & 1u)) continue;
            const int b0 = (int)(dw[u] & 0xffffu);
            wb[u] = ((unsigned)__ldg(src + b0) << 8) | (unsigned)__ldg(src + min(b0 + 1, sym_bytes - 1));
        }
    }
    __syncthreads();                                                              // ct[] complete, the mbarrier initialised
    if (staged) {
        mbar_wait(mbar, 0);
#pragma unroll
        for (int u = 0; u < 16; u++) {
            wb[u] = 0u;
            if (!((tmask >> u) & 1u)) continue;
            const int b0 = (int)(dw[u] & 0xffffu);
            wb[u] = ((unsigned)pl[b0] << 8) | (unsigned)pl[min(b0 + 1, sym_bytes - 1)];
        }
    }
    float2 v[16];
    const float2 *ctl = skew ? ct + (tid & 15) : ct;
    const int csh = skew ? 4 : 0;
#pragma unroll
    for (int u = 0; u < 16; u++) {
        v[u] = make_float2(0.f, 0.f);
        if (!((tmask >> u) & 1u)) continue;
        const unsigned off = (dw[u] >> 16) & 7u, flag = dw[u] >> 24;
        const unsigned sym = (wb[u] >> (16u - (unsigned)mod - off)) & (unsigned)(npts - 1);      // Frame.cpp:59-62 + modulation.cpp:39-50
        v[u] = ctl[(flag ? (unsigned)npts - 1u + flag : sym) << csh];
    }
    auto sync = [&]() { __syncthreads(); };
    cta_fft4096<false>(v, E, P.tw_fft, tid, sync);
    const float sc = 0.015625f;                                                    // 1 / sqrt(4096)  (Frame.cpp:66-68)
    const float2 scj = make_float2(sc, -sc);
    const long long base = P.t2sin_size + P.pf_size + (long long)s * kBigL;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const float2 y = p_mul(v[t], scj);
        big_store<FMT>(fout, base + kBigCP + tid + 256 * t, y, P.mult);              // Frame.cpp:191-192
        if (t >= 12) big_store<FMT>(fout, base + tid + 256 * (t - 12), y, P.mult);   // Frame.cpp:196-197
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
    big_frame_phasors(f, P);
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
