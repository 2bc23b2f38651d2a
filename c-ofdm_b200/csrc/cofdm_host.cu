// cofdm_host.cu -- implementation of the C ABI in include/cofdm.h: handle management, host-side
// constant tables (host_consts.hpp), kernel launches for sm_100a, and the COFDM_HOST staging path
// (chunked H2D -> kernel -> D2H, double-buffered over two streams).  No CPU compute fallback exists: without a
// usable CUDA device every computing entry point returns COFDM_ERR_CUDA.
#include "../../include/cofdm.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "big.cuh"
#include "stream.cuh"
#include "host_consts.hpp"

using namespace cofdmk;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(COFDM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));      \
    } while (0)

constexpr int kPipe = 8;   // maximum streams / buffer sets of the COFDM_HOST pipeline (depth in use: cofdm::pipe_depth)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
// pinned host staging (small results of per-frame calls: one asynchronous copy instead of several synchronous ones)
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaHostAlloc(&p, n, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

struct cofdm {
    int device = 0;
    HostTables T;
    Params P{};
    std::vector<void *> table_allocs;
    const float2 *constell_dev[9] = {};
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t pipe_stream[kPipe] = {};
    DevBuf pipe_in[kPipe], pipe_out[kPipe];
    DevBuf scratch_a, scratch_b, scratch_c;
    DevBuf ring;                                 // cofdm_ring_load: the receiver's int16 ring, resident
    DevBuf coll;                                 // cofdm_allreduce_counters staging
    PinBuf pin;                                  // pinned staging of the tapped rx call's results
    // intermediates between the kernels of one rx pass, ONE SET PER PIPELINE SLOT: the COFDM_HOST pipeline runs consecutive
    // chunks on different streams, so chunk c + 1's first kernel must not overwrite what chunk c's last kernel still reads
    DevBuf gen_frames[kPipe], gen_spec[kPipe], gen_pre[kPipe];   // any-size path
    DevBuf fscal[kPipe];                         // per-frame scalars handed from the acquire to the demod kernel
    int big_acquire = 1;                         // ... including their own acquisition kernel (env COFDM_BIG_ACQUIRE=0: the any-size kernels + bridge)
    int big_cluster_policy = 2;                  // ... cluster scheduling policy preference of the demod launch: load balancing (measured 2.44 vs 2.48 ms
                                                 //     per 16 384 frames against the default and "spread"; env COFDM_BIG_CLUSTER_POLICY = 0 / 1 / 2)
    int big_lay_on = 1;                          // ... the layout-specialised demod instance where the map allows it (env COFDM_BIG_LAY=0: off)
    int big_on = 1;                              // fft-4096 configurations use the cluster kernels of big.cuh (env COFDM_BIG=0: the any-size path)
    int tx_ctas = 148 * 4;                       // CTAs of the persistent tx kernel (SMs x resident CTAs per SM, measured at create)
    int tx_bulk = 1;                             // tx: symbols leave the SM as TMA bulk stores (env COFDM_TX_BULK=0: register stores)
    int pipe_depth = 2;                          // streams in flight (env COFDM_PIPE_DEPTH, <= kPipe); measured on B200:
                                                 // 2 reaches the PCIe full-duplex ceiling, 3 and more lose 10-15 %
    size_t pipe_chunk = 2048;                    // frames per chunk of the COFDM_HOST pipeline (env COFDM_PIPE_CHUNK)
    unsigned long long *amb_dev = nullptr;       // ambiguity counter
    unsigned long long *pos_dev = nullptr;       // find_t2sin result
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timing = false, timed = false;
    // per-stage device times of the last call (cofdm_last_stage_ms), CUDA events on the launching stream
    cudaEvent_t sev[8] = {};
    float stage_ms[COFDM_STAGE_COUNT] = {};
    bool rx_events_pending = false;
    unsigned long long launches = 0;
};

namespace {

template <class T>
int upload(cofdm *h, const std::vector<T> &v, const T **dst) {
    void *d = nullptr;
    const size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    CU_TRY(cudaMalloc(&d, bytes));
    h->table_allocs.push_back(d);
    if (!v.empty()) CU_TRY(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dst = reinterpret_cast<const T *>(d);
    return COFDM_OK;
}

struct Timed {
    cofdm *h;
    explicit Timed(cofdm *h_) : h(h_) { if (h->timing) cudaEventRecord(h->ev0, h->stream); }
    ~Timed() { if (h->timing) { cudaEventRecord(h->ev1, h->stream); h->timed = true; } }
};

// adds the acquire / demod kernel times of the last launch_rx to stage_ms (blocks until that launch has finished)
void collect_rx_stage(cofdm *h) {
    if (!h->rx_events_pending) return;
    h->rx_events_pending = false;
    float a = 0.f, d = 0.f;
    if (cudaEventSynchronize(h->sev[2]) != cudaSuccess) return;
    if (cudaEventElapsedTime(&a, h->sev[0], h->sev[1]) == cudaSuccess) h->stage_ms[COFDM_STAGE_ACQUIRE] += a;
    if (cudaEventElapsedTime(&d, h->sev[1], h->sev[2]) == cudaSuccess) h->stage_ms[COFDM_STAGE_DEMOD] += d;
}
void add_stage(cofdm *h, int stage, cudaEvent_t e0, cudaEvent_t e1) {
    float ms = 0.f;
    if (cudaEventSynchronize(e1) == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) h->stage_ms[stage] += ms;
}

int check_launch(cofdm *h, const char *what) {
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(COFDM_ERR_CUDA, std::string(what) + " launch: " + cudaGetErrorString(e));
    return COFDM_OK;
}

size_t sample_bytes(int fmt) { return fmt == COFDM_CI16 ? 4 : 8; }

// ---- device-side launches (all pointers are device pointers, stream given) ----------------------
int launch_rx_generic(cofdm *h, cudaStream_t st, const void *samples, int fmt, size_t n_frames, size_t stride,
                      uint8_t *bytes, unsigned long long *amb, const RxTaps &taps, int slot, int sync_less = 0);
int launch_rx_big(cofdm *h, cudaStream_t st, const void *samples, int fmt, size_t n_frames, size_t stride,
                  uint8_t *bytes, unsigned long long *amb, const RxTaps &taps, int slot);
int launch_tx_generic(cofdm *h, cudaStream_t st, const uint8_t *payload, size_t n_frames, void *frames, int fmt);
// frame_pos (device, optional): record f starts at sample frame_pos[f] of `samples` (a buffer of buf_samples samples) instead
// of at f * stride: the frames a scanner found are demodulated in place
int launch_rx(cofdm *h, cudaStream_t st, const void *samples, int fmt, size_t n_frames, size_t stride,
              uint8_t *bytes, unsigned long long *amb, const RxTaps &taps, int sync_less = 0, int slot = 0,
              const long long *frame_pos = nullptr, size_t buf_samples = 0) {
    if (n_frames == 0) return COFDM_OK;
    if (!h->T.fused512_ok) {
        if (h->T.big_ok && h->big_on && !sync_less) return launch_rx_big(h, st, samples, fmt, n_frames, stride, bytes, amb, taps, slot);
        if (h->T.generic_ok) return launch_rx_generic(h, st, samples, fmt, n_frames, stride, bytes, amb, taps, slot, sync_less);
        return fail(COFDM_ERR_UNSUPPORTED, "rx: configuration outside both the fused fft-512 path and the generic path (see DESIGN.md section 7)");
    }
    if ((uintptr_t)bytes & 3) return fail(COFDM_ERR_ARG, "rx: the output byte buffer must be 4-byte aligned");
    const Params &P = h->P;
    const bool want = taps.scal || taps.grid || taps.chan || taps.constell || taps.synced;
    // records that are 16-byte aligned are staged by TMA bulk copies (cf32, or raw int16 wire data widened when read);
    // a frame cut out of a capture at an arbitrary sample is staged by the warp with plain loads
    const size_t sb = sample_bytes(fmt);
    // int16 records always take the TMA instances: the kernels copy from the aligned address below an unaligned record
    bool al = ((uintptr_t)samples & 15) == 0 && (stride * sb) % 16 == 0 && frame_pos == nullptr;
    if (fmt == COFDM_CI16 && ((uintptr_t)samples & 3) == 0) al = true;
    if (frame_pos != nullptr && fmt != COFDM_CI16) return fail(COFDM_ERR_ARG, "rx: in-place frame positions need int16 records");
    RxSrc rs{};
    rs.frame_pos = frame_pos;
    rs.lo = (const char *)samples;
    rs.hi = (const char *)samples + (frame_pos != nullptr ? buf_samples : (n_frames - 1) * stride + (size_t)P.rx_len) * sb;
    // the acquire kernel's hand-over (56 bytes per frame), in the buffer of this pipeline slot
    FrameScal *fsc = nullptr;
    if (!sync_less) {
        CU_TRY(h->fscal[slot].reserve(n_frames * sizeof(FrameScal)));
        fsc = (FrameScal *)h->fscal[slot].p;
    }
    // ---- acquire: the preamble -> 56 bytes of scalars per frame, one warp per frame.  The sync-less form (FRAME_FORM::read)
    //      has no synchronisation stage at all; its preamble is only looked at for the chan_char tap ----
    if (h->timing) { collect_rx_stage(h); cudaEventRecord(h->sev[0], st); }
    if (!sync_less || taps.chan != nullptr) {
        const unsigned g4 = (unsigned)((n_frames + kAcqwWarps - 1) / kAcqwWarps);
#define COFDM_ACQW(F, T) \
        do { if (want) rx_acquire512w_kernel<F, T, true><<<g4, 32 * kAcqwWarps, rx_acquire512w_smem_bytes(), st>>>(P, samples, (long long)stride, (int)n_frames, taps, fsc, sync_less, rs); \
             else rx_acquire512w_kernel<F, T, false><<<g4, 32 * kAcqwWarps, rx_acquire512w_smem_bytes(), st>>>(P, samples, (long long)stride, (int)n_frames, taps, fsc, sync_less, rs); } while (0)
        if (fmt == COFDM_CI16) { if (al) COFDM_ACQW(kCI16, true); else COFDM_ACQW(kCI16, false); }
        else { if (al) COFDM_ACQW(kCF32, true); else COFDM_ACQW(kCF32, false); }
#undef COFDM_ACQW
        if (int rc = check_launch(h, "rx_acquire512w")) return rc;
    }
    if (h->timing) cudaEventRecord(h->sev[1], st);
    // ---- demod: the message symbols -> payload bytes, one warp per symbol ----
    {
        const size_t sm = rx_demod512_smem_bytes(P.num_symb);
        const unsigned thr = 32u * (unsigned)P.num_symb;
#define COFDM_DM(F, T, W, MW, MD) rx_demod512_kernel<F, T, W, MW, MD><<<(unsigned)n_frames, thr, sm, st>>>(P, samples, (long long)stride, (int)n_frames, bytes, amb, taps, sync_less, fsc, rs)
        // production instances are specialised on the modulation order (QPSK, 16-QAM); everything else reads it from the configuration
#define COFDM_DM_PICK(F, T) do { if (P.num_symb <= 8) { if (want) COFDM_DM(F, T, true, 8, 0); else if (P.mod_type == 4) COFDM_DM(F, T, false, 8, 4); \
                                                       else if (P.mod_type == 2) COFDM_DM(F, T, false, 8, 2); else COFDM_DM(F, T, false, 8, 0); } \
                             else { if (want) COFDM_DM(F, T, true, kMaxFusedSymb, 0); else COFDM_DM(F, T, false, kMaxFusedSymb, 0); } } while (0)
        if (fmt == COFDM_CI16) { if (al) COFDM_DM_PICK(kCI16, true); else COFDM_DM_PICK(kCI16, false); }
        else { if (al) COFDM_DM_PICK(kCF32, true); else COFDM_DM_PICK(kCF32, false); }
#undef COFDM_DM_PICK
#undef COFDM_DM
        if (int rc = check_launch(h, "rx_demod512")) return rc;
    }
    if (h->timing) { cudaEventRecord(h->sev[2], st); h->rx_events_pending = true; }
    if (taps.synced != nullptr && taps.scal != nullptr && !sync_less) {
        rx_synced_fixup_kernel<<<(unsigned)n_frames, 128, 0, st>>>(P, (int)n_frames, taps);
        return check_launch(h, "rx_synced_fixup");
    }
    return COFDM_OK;
}

// the fft-4096 path (big.cuh): acquisition -> FrameScal, then one cluster of num_symb CTAs per frame
template <int FMT, bool TMA, bool TAPS, int MOD, bool LAY = false>
int launch_big_demod(cofdm *h, cudaStream_t st, const void *samples, size_t n_frames, size_t stride, uint8_t *bytes,
                     unsigned long long *amb, const RxTaps &taps, const FrameScal *fsc) {
    const Params &P = h->P;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(n_frames * (size_t)P.num_symb));
    cfg.blockDim = dim3(kBigThreads);
    cfg.dynamicSmemBytes = big_smem_bytes();
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)P.num_symb; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeClusterSchedulingPolicyPreference;     // 0 the driver's default, 1 spread, 2 load balancing
    attr[1].val.clusterSchedulingPolicyPreference = (cudaClusterSchedulingPolicy)h->big_cluster_policy;
    cfg.attrs = attr; cfg.numAttrs = h->big_cluster_policy ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, big_demod_kernel<FMT, TMA, TAPS, MOD, LAY>, P, samples, (long long)stride, (int)n_frames, bytes, amb, taps, fsc);
    if (e != cudaSuccess) return fail(COFDM_ERR_CUDA, std::string("big_demod launch: ") + cudaGetErrorString(e));
    return check_launch(h, "big_demod");
}

int launch_rx_big(cofdm *h, cudaStream_t st, const void *samples, int fmt, size_t n_frames, size_t stride,
                  uint8_t *bytes, unsigned long long *amb, const RxTaps &taps, int slot) {
    const Params &P = h->P;
    const size_t N = (size_t)P.fft_size, L = (size_t)P.ofdm_len;
    const size_t sb = sample_bytes(fmt);
    const bool want = taps.scal || taps.chan || taps.constell;
    const bool al = ((uintptr_t)samples & 15) == 0 && (stride * sb) % 16 == 0;
    CU_TRY(h->fscal[slot].reserve(n_frames * sizeof(FrameScal)));
    FrameScal *fsc = (FrameScal *)h->fscal[slot].p;
    if (h->timing) { collect_rx_stage(h); cudaEventRecord(h->sev[0], st); }
    if (h->big_acquire) {
#define COFDM_BACQ(F, T) do { if (want) big_acquire_kernel<F, T, true><<<(unsigned)n_frames, kBigThreads, big_acquire_smem_bytes(), st>>>(P, samples, (long long)stride, (int)n_frames, taps, fsc); \
                              else big_acquire_kernel<F, T, false><<<(unsigned)n_frames, kBigThreads, big_acquire_smem_bytes(), st>>>(P, samples, (long long)stride, (int)n_frames, taps, fsc); } while (0)
        if (fmt == COFDM_CI16) { if (al) COFDM_BACQ(kCI16, true); else COFDM_BACQ(kCI16, false); }
        else { if (al) COFDM_BACQ(kCF32, true); else COFDM_BACQ(kCF32, false); }
#undef COFDM_BACQ
        if (int rc = check_launch(h, "big_acquire")) return rc;
    } else {
        // acquisition by the any-size kernels on the preamble only (sub-batches bound their scratch memory)
        const size_t sub = 4096, nb = std::min(sub, n_frames);
        CU_TRY(h->gen_frames[slot].reserve(nb * sizeof(GenFrame)));
        CU_TRY(h->gen_spec[slot].reserve(nb * N * sizeof(float2)));
        CU_TRY(h->gen_pre[slot].reserve(nb * L * sizeof(float2)));
        GenFrame *gf = (GenFrame *)h->gen_frames[slot].p;
        float2 *spec = (float2 *)h->gen_spec[slot].p, *pre = (float2 *)h->gen_pre[slot].p;
        const size_t sm_c = 2 * (size_t)P.pf_size * sizeof(float2), sm_s = (L + N) * sizeof(float2), sm_h = (size_t)P.num_data_subc / 2 * sizeof(float) + 16;
        for (size_t f0 = 0; f0 < n_frames; f0 += sub) {
            const int n = (int)std::min(sub, n_frames - f0);
            const void *src = (const char *)samples + f0 * stride * sb;
            RxTaps t = taps;
            if (t.scal) t.scal += f0 * 48;
            if (t.chan) t.chan += f0 * (size_t)P.num_data_subc;
            if (fmt == COFDM_CI16) {
                gen_coarse_kernel<kCI16><<<n, kGenThreads, sm_c, st>>>(P, src, (long long)stride, n, gf);
                if (int rc = check_launch(h, "gen_coarse")) return rc;
                gen_symbol_kernel<kCI16, true><<<dim3(1, n), kGenThreads, sm_s, st>>>(P, src, (long long)stride, n, gf, spec, pre);
            } else {
                gen_coarse_kernel<kCF32><<<n, kGenThreads, sm_c, st>>>(P, src, (long long)stride, n, gf);
                if (int rc = check_launch(h, "gen_coarse")) return rc;
                gen_symbol_kernel<kCF32, true><<<dim3(1, n), kGenThreads, sm_s, st>>>(P, src, (long long)stride, n, gf, spec, pre);
            }
            if (int rc = check_launch(h, "gen_symbol")) return rc;
            gen_chan_kernel<true><<<n, kGenThreads, sm_h, st>>>(P, n, gf, spec, pre);
            if (int rc = check_launch(h, "gen_chan")) return rc;
            big_bridge_kernel<<<(n + 127) / 128, 128, 0, st>>>(P, n, gf, fsc + f0, t);
            if (int rc = check_launch(h, "big_bridge")) return rc;
        }
    }
    if (h->timing) cudaEventRecord(h->sev[1], st);
    int rc;
    // the production instance is specialised on 64-QAM and, when the sub-carrier map has it (P.big_lay), on the row layout;
    // everything else reads the modulation order and the roles from the configuration
    const bool lay = P.big_lay && h->big_lay_on;
#define COFDM_BIG(F, T) (want ? launch_big_demod<F, T, true, 0>(h, st, samples, n_frames, stride, bytes, amb, taps, fsc) \
                              : (P.mod_type == 6 ? (lay ? launch_big_demod<F, T, false, 6, true>(h, st, samples, n_frames, stride, bytes, amb, taps, fsc) \
                                                        : launch_big_demod<F, T, false, 6>(h, st, samples, n_frames, stride, bytes, amb, taps, fsc)) \
                                                 : launch_big_demod<F, T, false, 0>(h, st, samples, n_frames, stride, bytes, amb, taps, fsc)))
    if (fmt == COFDM_CI16) rc = al ? COFDM_BIG(kCI16, true) : COFDM_BIG(kCI16, false);
    else rc = al ? COFDM_BIG(kCF32, true) : COFDM_BIG(kCF32, false);
#undef COFDM_BIG
    if (h->timing) { cudaEventRecord(h->sev[2], st); h->rx_events_pending = true; }
    return rc;
}

// the any-size path (generic.cuh): five kernels per sub-batch with the spectra in HBM between them
int launch_rx_generic(cofdm *h, cudaStream_t st, const void *samples, int fmt, size_t n_frames, size_t stride,
                      uint8_t *bytes, unsigned long long *amb, const RxTaps &taps, int slot, int sync_less) {
    const Params &P = h->P;
    if (sync_less && taps.chan != nullptr)
        return fail(COFDM_ERR_UNSUPPORTED, "read: the chan_char output is only built for the fft-512 geometry");
    const size_t sub = 512, N = (size_t)P.fft_size, L = (size_t)P.ofdm_len, nsym = (size_t)P.n_sym_rx;
    const size_t nb = std::min(sub, n_frames);
    CU_TRY(h->gen_frames[slot].reserve(nb * sizeof(GenFrame)));
    CU_TRY(h->gen_spec[slot].reserve(nb * nsym * N * sizeof(float2)));
    CU_TRY(h->gen_pre[slot].reserve(nb * (size_t)P.pf_size * sizeof(float2)));
    GenFrame *gf = (GenFrame *)h->gen_frames[slot].p;
    float2 *spec = (float2 *)h->gen_spec[slot].p, *pre = (float2 *)h->gen_pre[slot].p;
    const size_t sm_c = 2 * (size_t)P.pf_size * sizeof(float2), sm_s = (L + N) * sizeof(float2), sm_h = (size_t)P.num_data_subc / 2 * sizeof(float) + 16;
    const size_t sb = sample_bytes(fmt);
    for (size_t f0 = 0; f0 < n_frames; f0 += sub) {
        const int n = (int)std::min(sub, n_frames - f0);
        const void *src = (const char *)samples + f0 * stride * sb;
        RxTaps t = taps;
        if (t.scal) t.scal += f0 * 48;
        if (t.chan) t.chan += f0 * (size_t)P.num_data_subc;
        if (t.constell) t.constell += f0 * (size_t)P.num_data_subc * P.num_symb;
        // (the sync-less form, FRAME_FORM::read, has no coarse-CFO stage and no rotation)
        if (fmt == COFDM_CI16) {
            if (!sync_less) gen_coarse_kernel<kCI16><<<n, kGenThreads, sm_c, st>>>(P, src, (long long)stride, n, gf);
            if (int rc = check_launch(h, "gen_coarse")) return rc;
            gen_symbol_kernel<kCI16><<<dim3((unsigned)nsym, n), kGenThreads, sm_s, st>>>(P, src, (long long)stride, n, gf, spec, pre, sync_less);
        } else {
            if (!sync_less) gen_coarse_kernel<kCF32><<<n, kGenThreads, sm_c, st>>>(P, src, (long long)stride, n, gf);
            if (int rc = check_launch(h, "gen_coarse")) return rc;
            gen_symbol_kernel<kCF32><<<dim3((unsigned)nsym, n), kGenThreads, sm_s, st>>>(P, src, (long long)stride, n, gf, spec, pre, sync_less);
        }
        if (int rc = check_launch(h, "gen_symbol")) return rc;
        gen_chan_kernel<false><<<n, kGenThreads, sm_h, st>>>(P, n, gf, spec, pre, sync_less);
        if (int rc = check_launch(h, "gen_chan")) return rc;
        gen_demap_kernel<<<dim3((unsigned)P.num_symb, n), kGenThreads, 0, st>>>(P, n, gf, spec, bytes + f0 * (size_t)P.bytes_per_frame, amb, t);
        if (int rc = check_launch(h, "gen_demap")) return rc;
    }
    return COFDM_OK;
}

int launch_tx_generic(cofdm *h, cudaStream_t st, const uint8_t *payload, size_t n_frames, void *frames, int fmt) {
    const Params &P = h->P;
    if (h->T.big_ok && h->big_on) {
        // the fft-4096 geometry (big.cuh): one 256-thread CTA per symbol, FFT in registers
        for (size_t f0 = 0; f0 < n_frames; f0 += 32768) {              // grid.y limit
            const int n = (int)std::min<size_t>(32768, n_frames - f0);
            const uint8_t *pl = payload + f0 * (size_t)P.bytes_per_frame;
            void *out = (char *)frames + f0 * (size_t)P.frame_len * sample_bytes(fmt);
            const dim3 grid((unsigned)P.num_symb + 1, n);
            // the production instance is specialised on 64-QAM and on the row layout (P.big_lay); everything else reads both from the configuration
            const bool spec = P.mod_type == 6 && P.big_lay && h->big_lay_on;
            if (fmt == COFDM_CI16) { if (spec) big_tx_kernel<kCI16, 6, true><<<grid, kBigThreads, big_tx_smem_bytes(), st>>>(P, pl, n, out);
                                     else big_tx_kernel<kCI16><<<grid, kBigThreads, big_tx_smem_bytes(), st>>>(P, pl, n, out); }
            else { if (spec) big_tx_kernel<kCF32, 6, true><<<grid, kBigThreads, big_tx_smem_bytes(), st>>>(P, pl, n, out);
                   else big_tx_kernel<kCF32><<<grid, kBigThreads, big_tx_smem_bytes(), st>>>(P, pl, n, out); }
            if (int rc = check_launch(h, "big_tx")) return rc;
        }
        return COFDM_OK;
    }
    const size_t sm = 2 * (size_t)P.fft_size * sizeof(float2);
    for (size_t f0 = 0; f0 < n_frames; f0 += 32768) {                  // grid.y limit
        const int n = (int)std::min<size_t>(32768, n_frames - f0);
        const uint8_t *pl = payload + f0 * (size_t)P.bytes_per_frame;
        void *out = (char *)frames + f0 * (size_t)P.frame_len * sample_bytes(fmt);
        const dim3 grid((unsigned)P.num_symb + 1, n);
        if (fmt == COFDM_CI16) gen_tx_kernel<kCI16><<<grid, kGenThreads, sm, st>>>(P, pl, n, out);
        else gen_tx_kernel<kCF32><<<grid, kGenThreads, sm, st>>>(P, pl, n, out);
        if (int rc = check_launch(h, "gen_tx")) return rc;
    }
    return COFDM_OK;
}

int launch_tx(cofdm *h, cudaStream_t st, const uint8_t *payload, size_t n_frames, void *frames, int fmt) {
    if (n_frames == 0) return COFDM_OK;
    if (!h->T.fused512_ok) {
        if (h->T.generic_ok) return launch_tx_generic(h, st, payload, n_frames, frames, fmt);
        return fail(COFDM_ERR_UNSUPPORTED, "tx: configuration outside both the fused fft-512 path and the generic path (see DESIGN.md section 7)");
    }
    // frames leave the SM as TMA bulk stores when the buffer is 16-byte aligned (frame and symbol sizes are multiples of 16 bytes)
    const bool bulk = h->tx_bulk && ((uintptr_t)frames & 15) == 0 && ((size_t)h->P.frame_len * sample_bytes(fmt)) % 16 == 0 &&
                      ((size_t)(h->P.t2sin_size + h->P.pf_size) * sample_bytes(fmt)) % 16 == 0;
    {
        // one warp per symbol (tx512w.cuh)
        // persistent CTAs: as many as fit on the device at once, each walks over frames
        const size_t smw = tx512w_smem_bytes(h->P.num_symb, h->P.t2sin_size + h->P.pf_size);
        const unsigned thr = (unsigned)tx512w_threads(h->P.num_symb);
        const unsigned grd = (unsigned)std::min<size_t>(n_frames, (size_t)h->tx_ctas);
#define COFDM_TXW(F, B) do { if (h->P.num_symb <= 8) tx512w_kernel<F, B, 8><<<grd, thr, smw, st>>>(h->P, payload, (int)n_frames, frames); \
                             else tx512w_kernel<F, B, kMaxFusedSymb><<<grd, thr, smw, st>>>(h->P, payload, (int)n_frames, frames); } while (0)
        if (fmt == COFDM_CI16) { if (bulk) COFDM_TXW(kCI16, true); else COFDM_TXW(kCI16, false); }
        else { if (bulk) COFDM_TXW(kCF32, true); else COFDM_TXW(kCF32, false); }
#undef COFDM_TXW
        return check_launch(h, "tx512w");
    }
}

int launch_t2(cofdm *h, cudaStream_t st, const void *samples, int fmt, size_t start, size_t n_blocks, float *rel) {
    if (n_blocks == 0) return COFDM_OK;
    if (h->P.t2sin_size != 256) {
        // any power-of-two block size from 16 to 1024: one warp per block, every bin evaluated
        const int n = h->P.t2sin_size;
        if (n < 16 || n > 1024 || (n & (n - 1)) != 0) return fail(COFDM_ERR_UNSUPPORTED, "t2sin: T2sin_size must be a power of two in [16, 1024]");
        const unsigned grid = (unsigned)((n_blocks + kT2AnyWarps - 1) / kT2AnyWarps);
        if (fmt == COFDM_CI16) t2sin_metric_any_kernel<kCI16><<<grid, 32 * kT2AnyWarps, t2sin_any_smem_bytes(n), st>>>(h->P, samples, (long long)start, (long long)n_blocks, rel);
        else t2sin_metric_any_kernel<kCF32><<<grid, 32 * kT2AnyWarps, t2sin_any_smem_bytes(n), st>>>(h->P, samples, (long long)start, (long long)n_blocks, rel);
        return check_launch(h, "t2sin_metric_any");
    }
    if (n_blocks >= 64) {                            // two blocks per warp, packed arithmetic
        const unsigned g2 = (unsigned)(((n_blocks + 1) / 2 + kT2PairWarps - 1) / kT2PairWarps);
        if (fmt == COFDM_CI16) t2sin_metric2_kernel<kCI16><<<g2, 32 * kT2PairWarps, 0, st>>>(h->P, samples, (long long)start, (long long)n_blocks, rel);
        else t2sin_metric2_kernel<kCF32><<<g2, 32 * kT2PairWarps, 0, st>>>(h->P, samples, (long long)start, (long long)n_blocks, rel);
        return check_launch(h, "t2sin_metric2");
    }
    const unsigned grid = (unsigned)((n_blocks + kT2WarpsPerCta - 1) / kT2WarpsPerCta);
    if (fmt == COFDM_CI16) t2sin_metric_kernel<kCI16><<<grid, 32 * kT2WarpsPerCta, 0, st>>>(h->P, samples, (long long)start, (long long)n_blocks, rel);
    else t2sin_metric_kernel<kCF32><<<grid, 32 * kT2WarpsPerCta, 0, st>>>(h->P, samples, (long long)start, (long long)n_blocks, rel);
    return check_launch(h, "t2sin_metric");
}

int launch_pc(cofdm *h, cudaStream_t st, const void *samples, int fmt, size_t n_samples, const long long *starts,
              size_t n_starts, float *cor, long long *first) {
    if (n_starts == 0) return COFDM_OK;
    const size_t sm = (size_t)(h->P.cor_size + 2 * h->P.pr_sin_len) * sizeof(float2);
    if (sm > 200 * 1024) return fail(COFDM_ERR_UNSUPPORTED, "preamble search window exceeds shared memory");
    if ((h->P.pr_sin_len % 4) == 0 && (h->P.cor_size % 4) == 0) {
        // 4 consecutive lags per thread on a transposed window (kernels.cuh corr4_lags)
        const size_t sm4 = preamble_corr4_smem_bytes(h->P.cor_size, h->P.pr_sin_len);
        if (fmt == COFDM_CI16) preamble_corr4_kernel<kCI16><<<(unsigned)n_starts, kPc4Threads, sm4, st>>>(h->P, samples, (long long)n_samples, starts, (int)n_starts, cor, first);
        else preamble_corr4_kernel<kCF32><<<(unsigned)n_starts, kPc4Threads, sm4, st>>>(h->P, samples, (long long)n_samples, starts, (int)n_starts, cor, first);
        return check_launch(h, "preamble_corr4");
    }
    if (fmt == COFDM_CI16) preamble_corr_kernel<kCI16><<<(unsigned)n_starts, kPcThreads, sm, st>>>(h->P, samples, (long long)n_samples, starts, (int)n_starts, cor, first);
    else preamble_corr_kernel<kCF32><<<(unsigned)n_starts, kPcThreads, sm, st>>>(h->P, samples, (long long)n_samples, starts, (int)n_starts, cor, first);
    return check_launch(h, "preamble_corr");
}

int set_device(const cofdm *h) {
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return fail(COFDM_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return COFDM_OK;
}

}  // namespace

extern "C" {

const char *cofdm_last_error(void) { return g_err.c_str(); }
const char *cofdm_version(void) { return "cofdm_b200 0.1 (sm_100a)"; }

void cofdm_destroy(cofdm_t *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (void *p : h->table_allocs) cudaFree(p);
    for (int i = 0; i < kPipe; i++) {
        h->pipe_in[i].release(); h->pipe_out[i].release();
        if (h->pipe_stream[i]) cudaStreamDestroy(h->pipe_stream[i]);
    }
    h->scratch_a.release(); h->scratch_b.release(); h->scratch_c.release(); h->ring.release(); h->coll.release(); h->pin.release();
    for (int i = 0; i < kPipe; i++) { h->gen_frames[i].release(); h->gen_spec[i].release(); h->gen_pre[i].release(); h->fscal[i].release(); }
    if (h->amb_dev) cudaFree(h->amb_dev);
    if (h->pos_dev) cudaFree(h->pos_dev);
    for (auto &e : h->sev) if (e) cudaEventDestroy(e);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

int cofdm_create(const char *config_path, int device, cofdm_t **out) {
    if (!config_path || !out) return fail(COFDM_ERR_ARG, "cofdm_create: null argument");
    *out = nullptr;
    cofdm *h = new cofdm;
    try {
        h->T = build_tables(parse_config_file(config_path));
    } catch (const std::exception &e) {
        delete h;
        return fail(COFDM_ERR_CONFIG, e.what());
    }
    h->device = device;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        delete h;
        return fail(COFDM_ERR_CUDA, std::string("no usable CUDA device ") + std::to_string(device) + ": " +
                                        (e != cudaSuccess ? cudaGetErrorString(e) : "ordinal out of range") +
                                        " (this library has no CPU fallback)");
    }
    auto bail = [&](int rc) { cofdm_destroy(h); return rc; };
    if (set_device(h)) return bail(COFDM_ERR_CUDA);
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10)
        return bail(fail(COFDM_ERR_CUDA, "device is not sm_100 class; this library is built for sm_100a only"));
    HostTables &T = h->T;
    h->P = T.p;
    Params &P = h->P;
    int rc = 0;
    rc |= upload(h, T.tw_fft, &P.tw_fft);
    rc |= upload(h, T.tw_pf, &P.tw_pf);     rc |= upload(h, T.tw_t2, &P.tw_t2);   rc |= upload(h, T.t2_mask, &P.t2_mask);
    rc |= upload(h, T.t2_tone, &P.t2_tone); rc |= upload(h, T.preamble_td, &P.preamble_td);
    rc |= upload(h, T.matched, &P.matched); rc |= upload(h, T.mod_preamble, &P.mod_preamble);
    rc |= upload(h, T.bin_role, &P.bin_role); rc |= upload(h, T.big_roles, &P.big_roles); rc |= upload(h, T.big_eq, &P.big_eq); rc |= upload(h, T.big_txd, &P.big_txd); rc |= upload(h, T.bin_map, &P.bin_map); rc |= upload(h, T.data_bin, &P.data_bin); rc |= upload(h, T.pilot_bin, &P.pilot_bin);
    rc |= upload(h, T.lane_desc, &P.lane_desc); rc |= upload(h, T.lane_aux, &P.lane_aux); rc |= upload(h, T.acq_desc, &P.acq_desc);
    rc |= upload(h, T.grid_lane, &P.grid_lane); rc |= upload(h, T.tx_desc, &P.tx_desc);
    for (int m : {1, 2, 4, 6, 8}) rc |= upload(h, T.constell[m], &h->constell_dev[m]);
    if (rc) return bail(COFDM_ERR_CUDA);
    P.constell = h->constell_dev[P.mod_type];
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(COFDM_ERR_CUDA, "stream create"));
    h->stream = h->own_stream;
    for (int i = 0; i < kPipe; i++)
        if (cudaStreamCreateWithFlags(&h->pipe_stream[i], cudaStreamNonBlocking) != cudaSuccess) return bail(fail(COFDM_ERR_CUDA, "stream create"));
    if (cudaMalloc(&h->amb_dev, sizeof(unsigned long long)) != cudaSuccess) return bail(fail(COFDM_ERR_CUDA, "cudaMalloc"));
    if (cudaMalloc(&h->pos_dev, sizeof(unsigned long long)) != cudaSuccess) return bail(fail(COFDM_ERR_CUDA, "cudaMalloc"));
    for (auto &e : h->sev) cudaEventCreate(&e);
    cudaEventCreate(&h->ev0);
    cudaEventCreate(&h->ev1);
    {
        // host pipeline: chunks of about 2048 default-size frames (49 MB of int16 samples), whatever the frame size
        h->pipe_chunk = std::max<size_t>(1, (size_t)2048 * 6016 / (size_t)std::max(1, P.frame_len));
        const char *c = std::getenv("COFDM_PIPE_CHUNK");
        if (c && std::atoll(c) > 0) h->pipe_chunk = (size_t)std::atoll(c);
        const char *d = std::getenv("COFDM_PIPE_DEPTH");
        if (d && std::atoi(d) > 0) h->pipe_depth = std::min(std::atoi(d), kPipe);
    }
    if (T.fused512_ok) {
        cudaError_t a = cudaSuccess, b = cudaSuccess;
        {
            const char *tb = std::getenv("COFDM_TX_BULK");
            if (tb) h->tx_bulk = std::atoi(tb) != 0;
        }
        // maximum shared-memory carve-out: occupancy of both rx kernels is bounded by shared memory and registers, not by L1
        {
            const int smd = (int)rx_demod512_smem_bytes(P.num_symb);
#define COFDM_DM_ATTR(F, T, W, MW, MD) \
            if (a == cudaSuccess) a = cudaFuncSetAttribute(rx_demod512_kernel<F, T, W, MW, MD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smd); \
            if (a == cudaSuccess) a = cudaFuncSetAttribute(rx_demod512_kernel<F, T, W, MW, MD>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)
#define COFDM_DM_ATTR_ALL(F, T) COFDM_DM_ATTR(F, T, true, 8, 0); COFDM_DM_ATTR(F, T, false, 8, 0); COFDM_DM_ATTR(F, T, false, 8, 2); COFDM_DM_ATTR(F, T, false, 8, 4); \
                                COFDM_DM_ATTR(F, T, true, kMaxFusedSymb, 0); COFDM_DM_ATTR(F, T, false, kMaxFusedSymb, 0)
            COFDM_DM_ATTR_ALL(kCF32, true); COFDM_DM_ATTR_ALL(kCF32, false); COFDM_DM_ATTR_ALL(kCI16, true); COFDM_DM_ATTR_ALL(kCI16, false);
#undef COFDM_DM_ATTR_ALL
#undef COFDM_DM_ATTR
        }
#define COFDM_ACQW_ATTR(F, T, W) \
        if (a == cudaSuccess) a = cudaFuncSetAttribute(rx_acquire512w_kernel<F, T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rx_acquire512w_smem_bytes()); \
        if (a == cudaSuccess) a = cudaFuncSetAttribute(rx_acquire512w_kernel<F, T, W>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)
        COFDM_ACQW_ATTR(kCF32, true, true); COFDM_ACQW_ATTR(kCF32, true, false); COFDM_ACQW_ATTR(kCF32, false, true); COFDM_ACQW_ATTR(kCF32, false, false);
        COFDM_ACQW_ATTR(kCI16, true, true); COFDM_ACQW_ATTR(kCI16, true, false); COFDM_ACQW_ATTR(kCI16, false, true); COFDM_ACQW_ATTR(kCI16, false, false);
#undef COFDM_ACQW_ATTR
        {
            const int smw = (int)tx512w_smem_bytes(P.num_symb, P.t2sin_size + P.pf_size);
#define COFDM_TXW_ATTR(F, B, MW) \
            if (a == cudaSuccess) a = cudaFuncSetAttribute(tx512w_kernel<F, B, MW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smw); \
            if (a == cudaSuccess) a = cudaFuncSetAttribute(tx512w_kernel<F, B, MW>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)
            COFDM_TXW_ATTR(kCF32, true, 8); COFDM_TXW_ATTR(kCF32, false, 8); COFDM_TXW_ATTR(kCI16, true, 8); COFDM_TXW_ATTR(kCI16, false, 8);
            COFDM_TXW_ATTR(kCF32, true, kMaxFusedSymb); COFDM_TXW_ATTR(kCF32, false, kMaxFusedSymb);
            COFDM_TXW_ATTR(kCI16, true, kMaxFusedSymb); COFDM_TXW_ATTR(kCI16, false, kMaxFusedSymb);
#undef COFDM_TXW_ATTR
            int per_sm = 0;
            if (a == cudaSuccess) {
                if (P.num_symb <= 8) a = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tx512w_kernel<kCF32, true, 8>, tx512w_threads(P.num_symb), smw);
                else a = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tx512w_kernel<kCF32, true, kMaxFusedSymb>, tx512w_threads(P.num_symb), smw);
            }
            const char *tc = std::getenv("COFDM_TX_CTAS_PER_SM");
            if (tc && std::atoi(tc) > 0) per_sm = std::atoi(tc);
            h->tx_ctas = prop.multiProcessorCount * std::max(per_sm, 1);
        }
        if (a != cudaSuccess || b != cudaSuccess)
            return bail(fail(COFDM_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(a != cudaSuccess ? a : b)));
    }
    if (!T.fused512_ok && T.generic_ok) {
        const int sm_c = (int)(2 * (size_t)P.pf_size * sizeof(float2)), sm_s = (int)((size_t)(P.ofdm_len + P.fft_size) * sizeof(float2));
        const int sm_t = (int)(2 * (size_t)P.fft_size * sizeof(float2));
        cudaFuncSetAttribute(gen_coarse_kernel<kCF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_c);
        cudaFuncSetAttribute(gen_coarse_kernel<kCI16>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_c);
        cudaFuncSetAttribute(gen_symbol_kernel<kCF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_s);
        cudaFuncSetAttribute(gen_symbol_kernel<kCI16>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_s);
        cudaFuncSetAttribute(gen_tx_kernel<kCF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_t);
        cudaFuncSetAttribute(gen_tx_kernel<kCI16>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_t);
        if (T.big_ok) {
            const char *bg = std::getenv("COFDM_BIG");
            if (bg) h->big_on = std::atoi(bg) != 0;
            cudaFuncSetAttribute(gen_symbol_kernel<kCF32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_s);
            cudaFuncSetAttribute(gen_symbol_kernel<kCI16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_s);
            const char *bl = std::getenv("COFDM_BIG_LAY");           // 0: the layout-specialised demod instance off (A/B runs)
            if (bl) h->big_lay_on = std::atoi(bl) != 0;
            const char *bp = std::getenv("COFDM_BIG_CLUSTER_POLICY");
            if (bp) h->big_cluster_policy = std::max(0, std::min(2, std::atoi(bp)));
            const char *ba = std::getenv("COFDM_BIG_ACQUIRE");
            if (ba) h->big_acquire = std::atoi(ba) != 0;
            const int sma = (int)big_acquire_smem_bytes();
#define COFDM_BACQ_ATTR(F, T, W) \
            cudaFuncSetAttribute(big_acquire_kernel<F, T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, sma); \
            cudaFuncSetAttribute(big_acquire_kernel<F, T, W>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)
            COFDM_BACQ_ATTR(kCF32, true, true); COFDM_BACQ_ATTR(kCF32, true, false); COFDM_BACQ_ATTR(kCF32, false, true); COFDM_BACQ_ATTR(kCF32, false, false);
            COFDM_BACQ_ATTR(kCI16, true, true); COFDM_BACQ_ATTR(kCI16, true, false); COFDM_BACQ_ATTR(kCI16, false, true); COFDM_BACQ_ATTR(kCI16, false, false);
#undef COFDM_BACQ_ATTR
#define COFDM_BTX_ATTR(...) \
            cudaFuncSetAttribute(big_tx_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_tx_smem_bytes()); \
            cudaFuncSetAttribute(big_tx_kernel<__VA_ARGS__>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)
            COFDM_BTX_ATTR(kCF32); COFDM_BTX_ATTR(kCI16); COFDM_BTX_ATTR(kCF32, 6, true); COFDM_BTX_ATTR(kCI16, 6, true);
#undef COFDM_BTX_ATTR
            const int smb = (int)big_smem_bytes();
#define COFDM_BIG_ATTR1(F, T, W, MD) \
            cudaFuncSetAttribute(big_demod_kernel<F, T, W, MD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smb); \
            cudaFuncSetAttribute(big_demod_kernel<F, T, W, MD>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)
#define COFDM_BIG_ATTR(F, T) COFDM_BIG_ATTR1(F, T, true, 0); COFDM_BIG_ATTR1(F, T, false, 0); COFDM_BIG_ATTR1(F, T, false, 6); \
            cudaFuncSetAttribute(big_demod_kernel<F, T, false, 6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smb); \
            cudaFuncSetAttribute(big_demod_kernel<F, T, false, 6, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)
            COFDM_BIG_ATTR(kCF32, true); COFDM_BIG_ATTR(kCF32, false); COFDM_BIG_ATTR(kCI16, true); COFDM_BIG_ATTR(kCI16, false);
#undef COFDM_BIG_ATTR1
#undef COFDM_BIG_ATTR
        }
    }
    if (P.t2sin_size != 256 && P.t2sin_size >= 16 && P.t2sin_size <= 1024) {
        cudaFuncSetAttribute(t2sin_metric_any_kernel<kCF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t2sin_any_smem_bytes(P.t2sin_size));
        cudaFuncSetAttribute(t2sin_metric_any_kernel<kCI16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t2sin_any_smem_bytes(P.t2sin_size));
    }
    cudaFuncSetAttribute(stream_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stream_scan_smem_bytes(P.cor_size, P.pr_sin_len));
    {
        const int smp = (int)((size_t)(P.cor_size + 2 * P.pr_sin_len) * sizeof(float2));
        const int smp4 = (int)preamble_corr4_smem_bytes(P.cor_size, P.pr_sin_len);
        if (smp4 > 48 * 1024) {
            cudaFuncSetAttribute(preamble_corr4_kernel<kCF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smp4);
            cudaFuncSetAttribute(preamble_corr4_kernel<kCI16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smp4);
        }
        if (smp > 48 * 1024) {
            cudaFuncSetAttribute(preamble_corr_kernel<kCF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smp);
            cudaFuncSetAttribute(preamble_corr_kernel<kCI16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smp);
        }
    }
    *out = h;
    return COFDM_OK;
}

int cofdm_query(const cofdm_t *h, cofdm_sizes *o) {
    if (!h || !o) return fail(COFDM_ERR_ARG, "cofdm_query: null argument");
    const Params &P = h->P;
    std::memset(o, 0, sizeof *o);
    o->fft_size = P.fft_size; o->num_data_subc = P.num_data_subc; o->num_pilot_subc = P.num_pilot_subc;
    o->cp_size = P.cp_size; o->num_symb = P.num_symb; o->num_pr_symb = P.num_pr_symb;
    o->pr_sin_len = P.pr_sin_len; o->t2sin_size = P.t2sin_size; o->mod_type = P.mod_type;
    o->ofdm_len = P.ofdm_len; o->rx_len = P.rx_len; o->output_size = P.frame_len;
    o->usefull_size = P.bytes_per_frame; o->constell_size = P.num_data_subc * P.num_symb; o->cor_size = P.cor_size;
    o->mult = (int)P.mult; o->rx_buf_size = h->T.rx_buf_size; o->iterations = h->T.iterations;
    o->fused_path = h->T.fused512_ok ? 1 : (h->T.generic_ok ? 0 : -1); o->device = h->device;
    return COFDM_OK;
}

int cofdm_set_stream(cofdm_t *h, void *cuda_stream) {
    if (!h) return fail(COFDM_ERR_ARG, "null handle");
    h->stream = (cudaStream_t)cuda_stream;       // NULL is CUDA's (legacy) default stream, as everywhere in CUDA
    return COFDM_OK;
}

void *cofdm_own_stream(const cofdm_t *h) { return h ? (void *)h->own_stream : nullptr; }

int cofdm_synchronize(cofdm_t *h) {
    if (!h) return fail(COFDM_ERR_ARG, "null handle");
    if (set_device(h)) return COFDM_ERR_CUDA;
    CU_TRY(cudaStreamSynchronize(h->stream));
    return COFDM_OK;
}

int cofdm_enable_timing(cofdm_t *h, int on) {
    if (!h) return fail(COFDM_ERR_ARG, "null handle");
    h->timing = on != 0; h->timed = false; h->rx_events_pending = false;
    for (float &v : h->stage_ms) v = 0.f;
    return COFDM_OK;
}
int cofdm_last_stage_ms(cofdm_t *h, float *out, int n) {
    if (!h || !out || n < 0) return fail(COFDM_ERR_ARG, "cofdm_last_stage_ms: bad argument");
    if (!h->timing) return fail(COFDM_ERR_ARG, "cofdm_last_stage_ms: timing is off (cofdm_enable_timing)");
    collect_rx_stage(h);
    for (int i = 0; i < n; i++) out[i] = i < COFDM_STAGE_COUNT ? h->stage_ms[i] : 0.f;
    return COFDM_OK;
}

float cofdm_last_kernel_ms(const cofdm_t *h) {
    if (!h || !h->timed) return -1.0f;
    float ms = -1.0f;
    if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0f;
    return ms;
}
unsigned long long cofdm_launch_count(const cofdm_t *h) { return h ? h->launches : 0; }

int cofdm_get_constants(const cofdm_t *h, double *t2sin_tone, uint8_t *preamble_bytes, double *ofdm_preamble,
                        double *mod_preamble, double *matched, double *constell) {
    if (!h) return fail(COFDM_ERR_ARG, "null handle");
    const HostTables &T = h->T;
    auto put = [](double *dst, const std::vector<cd> &v) { if (dst) std::memcpy(dst, v.data(), v.size() * sizeof(cd)); };
    put(t2sin_tone, T.t2_tone_d); put(ofdm_preamble, T.preamble_td_d); put(mod_preamble, T.mod_preamble_d);
    put(matched, T.matched_d); put(constell, T.constell_d[T.p.mod_type]);
    if (preamble_bytes) std::memcpy(preamble_bytes, T.preamble_bytes.data(), T.preamble_bytes.size());
    return COFDM_OK;
}

static bool valid_mod(int m) { return m == 1 || m == 2 || m == 4 || m == 6 || m == 8; }

int cofdm_mod(cofdm_t *h, int mod, const uint8_t *bytes, size_t n_bytes, float *points, int space) {
    if (!h || !bytes || !points || !valid_mod(mod)) return fail(COFDM_ERR_ARG, "cofdm_mod: bad argument");
    if (set_device(h)) return COFDM_ERR_CUDA;
    const size_t n_pts = (n_bytes * 8 + mod - 1) / mod;
    if (n_pts == 0) return COFDM_OK;
    const uint8_t *din = bytes; float2 *dout = reinterpret_cast<float2 *>(points);
    if (space == COFDM_HOST) {
        CU_TRY(h->scratch_a.reserve(n_bytes)); CU_TRY(h->scratch_b.reserve(n_pts * sizeof(float2)));
        CU_TRY(cudaMemcpyAsync(h->scratch_a.p, bytes, n_bytes, cudaMemcpyHostToDevice, h->stream));
        din = (const uint8_t *)h->scratch_a.p; dout = (float2 *)h->scratch_b.p;
    }
    {
        Timed t(h);
        mod_kernel<<<(unsigned)((n_pts + 255) / 256), 256, 0, h->stream>>>(h->constell_dev[mod], mod, din, (long long)n_bytes, dout, (long long)n_pts);
        if (int rc = check_launch(h, "mod")) return rc;
    }
    if (space == COFDM_HOST) {
        CU_TRY(cudaMemcpyAsync(points, dout, n_pts * sizeof(float2), cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
    }
    return COFDM_OK;
}

int cofdm_demod(cofdm_t *h, int mod, const float *points, size_t n_points, uint8_t *bytes,
                unsigned long long *ambiguous, int space) {
    if (!h || !points || !bytes || !valid_mod(mod)) return fail(COFDM_ERR_ARG, "cofdm_demod: bad argument");
    if (set_device(h)) return COFDM_ERR_CUDA;
    const size_t n_bytes = (n_points * mod + 7) / 8;
    if (n_points == 0) return COFDM_OK;
    const float2 *din = reinterpret_cast<const float2 *>(points); uint8_t *dout = bytes;
    if (space == COFDM_HOST) {
        CU_TRY(h->scratch_a.reserve(n_points * sizeof(float2))); CU_TRY(h->scratch_b.reserve(n_bytes));
        CU_TRY(cudaMemcpyAsync(h->scratch_a.p, points, n_points * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
        din = (const float2 *)h->scratch_a.p; dout = (uint8_t *)h->scratch_b.p;
    }
    CU_TRY(cudaMemsetAsync(h->amb_dev, 0, sizeof(unsigned long long), h->stream));
    {
        Timed t(h);
        const size_t groups = (n_points + 7) / 8;
        demod_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, h->stream>>>(mod, din, (long long)n_points, dout, (long long)n_bytes, h->amb_dev);
        if (int rc = check_launch(h, "demod")) return rc;
    }
    if (space == COFDM_HOST) CU_TRY(cudaMemcpyAsync(bytes, dout, n_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (ambiguous || space == COFDM_HOST) {
        unsigned long long a = 0;
        CU_TRY(cudaMemcpyAsync(&a, h->amb_dev, sizeof a, cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        if (ambiguous) *ambiguous += a;
    }
    return COFDM_OK;
}

int cofdm_tx_batch(cofdm_t *h, const uint8_t *payload, size_t n_frames, void *frames, int fmt, int space) {
    if (!h || !payload || !frames || (fmt != COFDM_CF32 && fmt != COFDM_CI16)) return fail(COFDM_ERR_ARG, "cofdm_tx_batch: bad argument");
    if (set_device(h)) return COFDM_ERR_CUDA;
    if (space == COFDM_DEVICE) {
        Timed t(h);
        return launch_tx(h, h->stream, payload, n_frames, frames, fmt);
    }
    const size_t bpf = (size_t)h->P.bytes_per_frame, fb = (size_t)h->P.frame_len * sample_bytes(fmt);
    const size_t chunk = std::min<size_t>(n_frames, h->pipe_chunk);
    for (int i = 0; i < h->pipe_depth; i++) { CU_TRY(h->pipe_in[i].reserve(chunk * bpf)); CU_TRY(h->pipe_out[i].reserve(chunk * fb)); }
    auto run_chunks = [&]() -> int {
        size_t c = 0;
        for (size_t f0 = 0; f0 < n_frames; f0 += chunk, c++) {
            const size_t n = std::min(chunk, n_frames - f0);
            const int s = (int)(c % h->pipe_depth);
            cudaStream_t st = h->pipe_stream[s];
            CU_TRY(cudaMemcpyAsync(h->pipe_in[s].p, payload + f0 * bpf, n * bpf, cudaMemcpyHostToDevice, st));
            if (int rc = launch_tx(h, st, (const uint8_t *)h->pipe_in[s].p, n, h->pipe_out[s].p, fmt)) return rc;
            CU_TRY(cudaMemcpyAsync((char *)frames + f0 * fb, h->pipe_out[s].p, n * fb, cudaMemcpyDeviceToHost, st));
        }
        return COFDM_OK;
    };
    const int rc_chunks = run_chunks();       // on failure the queued copies are drained before returning (they touch the caller's buffers)
    for (int i = 0; i < kPipe; i++) {
        const cudaError_t e = cudaStreamSynchronize(h->pipe_stream[i]);
        if (e != cudaSuccess && rc_chunks == COFDM_OK) return fail(COFDM_ERR_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e));
    }
    return rc_chunks;
}

static int rx_batch_impl(cofdm_t *h, const void *samples, int fmt, size_t n_frames, size_t frame_stride,
                         uint8_t *bytes, unsigned long long *ambiguous, const cofdm_rx_taps *taps, int space, int sync_less);

int cofdm_rx_aligned_batch(cofdm_t *h, const void *samples, int fmt, size_t n_frames, size_t frame_stride,
                           uint8_t *bytes, unsigned long long *ambiguous, const cofdm_rx_taps *taps, int space) {
    return rx_batch_impl(h, samples, fmt, n_frames, frame_stride, bytes, ambiguous, taps, space, 0);
}

int cofdm_read_batch(cofdm_t *h, const void *frames, int fmt, size_t n_frames, uint8_t *bytes,
                     unsigned long long *ambiguous, float *restored, float *chan_char, int space) {
    if (!h || !frames) return fail(COFDM_ERR_ARG, "cofdm_read_batch: bad argument");
    cofdm_rx_taps t{nullptr, nullptr, chan_char, restored, nullptr};
    // FRAME_FORM::read copies a whole frame (Frame.cpp:240); the chain starts at the preamble slot
    const char *p = (const char *)frames + (size_t)h->P.t2sin_size * sample_bytes(fmt);
    return rx_batch_impl(h, p, fmt, n_frames, (size_t)h->P.frame_len, bytes, ambiguous, (restored || chan_char) ? &t : nullptr, space, 1);
}

static int rx_batch_impl(cofdm_t *h, const void *samples, int fmt, size_t n_frames, size_t frame_stride,
                         uint8_t *bytes, unsigned long long *ambiguous, const cofdm_rx_taps *taps, int space, int sync_less) {
    if (!h || !samples || !bytes || (fmt != COFDM_CF32 && fmt != COFDM_CI16)) return fail(COFDM_ERR_ARG, "cofdm_rx_aligned_batch: bad argument");
    if (frame_stride < (size_t)h->P.rx_len) return fail(COFDM_ERR_ARG, "frame_stride < rx_len");
    if (set_device(h)) return COFDM_ERR_CUDA;
    const Params &P = h->P;
    const bool want_taps = taps && (taps->scal || taps->grid || taps->chan || taps->constell || taps->synced);
    if (space == COFDM_DEVICE) {
        RxTaps t{};
        if (want_taps) {
            t.scal = taps->scal; t.grid = (float2 *)taps->grid; t.chan = (float2 *)taps->chan;
            t.constell = (float2 *)taps->constell; t.synced = (float2 *)taps->synced;
        }
        if (ambiguous) CU_TRY(cudaMemsetAsync(h->amb_dev, 0, sizeof(unsigned long long), h->stream));
        {
            Timed tm(h);
            if (int rc = launch_rx(h, h->stream, samples, fmt, n_frames, frame_stride, bytes, ambiguous ? h->amb_dev : nullptr, t, sync_less)) return rc;
        }
        if (ambiguous) {
            unsigned long long a = 0;
            CU_TRY(cudaMemcpyAsync(&a, h->amb_dev, sizeof a, cudaMemcpyDeviceToHost, h->stream));
            CU_TRY(cudaStreamSynchronize(h->stream));
            *ambiguous += a;
        }
        return COFDM_OK;
    }
    const size_t sb = sample_bytes(fmt), bpf = (size_t)P.bytes_per_frame;
    if (n_frames == 0) return COFDM_OK;
    if (want_taps) {
        // ---- host results WITH taps (parity checks, the per-frame facade): one stream, every result into one device block,
        //      ONE asynchronous copy into pinned staging, one synchronisation ----
        const size_t n_sc = 48, n_grid = (size_t)P.num_symb * P.fft_size, n_ch = (size_t)P.num_data_subc,
                     n_con = (size_t)P.num_data_subc * P.num_symb, n_syn = (size_t)P.rx_len;
        size_t off = 0;
        auto take = [&](bool on, size_t bytes_) { const size_t o = off; if (on) off += (bytes_ + 15) & ~(size_t)15; return o; };
        const size_t o_by = take(true, n_frames * bpf), o_amb = take(true, 16), o_sc = take(taps->scal || taps->synced, n_frames * n_sc * sizeof(float)),
                     o_gr = take(taps->grid, n_frames * n_grid * sizeof(float2)), o_ch = take(taps->chan, n_frames * n_ch * sizeof(float2)),
                     o_co = take(taps->constell, n_frames * n_con * sizeof(float2)), o_sy = take(taps->synced, n_frames * n_syn * sizeof(float2));
        CU_TRY(h->scratch_c.reserve(off));
        CU_TRY(h->pin.reserve(off));
        char *b = (char *)h->scratch_c.p, *hp = (char *)h->pin.p;
        cudaStream_t st = h->stream;
        const void *dsamp = samples;
        if (space != COFDM_DEVICE_IN) {
            const size_t in_bytes = ((n_frames - 1) * frame_stride + (size_t)P.rx_len) * sb;
            CU_TRY(h->pipe_in[0].reserve(in_bytes));
            CU_TRY(cudaMemcpyAsync(h->pipe_in[0].p, samples, in_bytes, cudaMemcpyHostToDevice, st));
            dsamp = h->pipe_in[0].p;
        }
        RxTaps t{};
        if (taps->scal || taps->synced) t.scal = (float *)(b + o_sc);
        if (taps->grid) t.grid = (float2 *)(b + o_gr);
        if (taps->chan) t.chan = (float2 *)(b + o_ch);
        if (taps->constell) t.constell = (float2 *)(b + o_co);
        if (taps->synced) t.synced = (float2 *)(b + o_sy);
        unsigned long long *amb = ambiguous ? (unsigned long long *)(b + o_amb) : nullptr;
        if (amb) CU_TRY(cudaMemsetAsync(amb, 0, sizeof(unsigned long long), st));
        if (int rc = launch_rx(h, st, dsamp, fmt, n_frames, frame_stride, (uint8_t *)(b + o_by), amb, t, sync_less, 0)) return rc;
        CU_TRY(cudaMemcpyAsync(hp, b, off, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        std::memcpy(bytes, hp + o_by, n_frames * bpf);
        if (ambiguous) *ambiguous += *(const unsigned long long *)(hp + o_amb);
        if (taps->scal) std::memcpy(taps->scal, hp + o_sc, n_frames * n_sc * sizeof(float));
        if (taps->grid) std::memcpy(taps->grid, hp + o_gr, n_frames * n_grid * sizeof(float2));
        if (taps->chan) std::memcpy(taps->chan, hp + o_ch, n_frames * n_ch * sizeof(float2));
        if (taps->constell) std::memcpy(taps->constell, hp + o_co, n_frames * n_con * sizeof(float2));
        if (taps->synced) std::memcpy(taps->synced, hp + o_sy, n_frames * n_syn * sizeof(float2));
        return COFDM_OK;
    }
    // ---- COFDM_HOST without taps: chunked, double-buffered pipeline ----------
    const size_t chunk = std::min<size_t>(n_frames, h->pipe_chunk);
    for (int i = 0; i < h->pipe_depth; i++) {
        if (space != COFDM_DEVICE_IN) CU_TRY(h->pipe_in[i].reserve(chunk * frame_stride * sb));
        CU_TRY(h->pipe_out[i].reserve(chunk * bpf));
    }
    CU_TRY(cudaMemsetAsync(h->amb_dev, 0, sizeof(unsigned long long), h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    RxTaps t{};
    // (on any failure the copies already queued on the other pipe streams are drained before returning: they touch the
    //  caller's buffers and the handle's staging buffers)
    auto run_chunks = [&]() -> int {
        size_t c = 0;
        for (size_t f0 = 0; f0 < n_frames; f0 += chunk, c++) {
            const size_t n = std::min(chunk, n_frames - f0);
            const int s = (int)(c % h->pipe_depth);
            cudaStream_t st = h->pipe_stream[s];
            // the last record only needs rx_len samples (the caller's buffer may end there)
            const size_t in_bytes = ((n - 1) * frame_stride + (size_t)P.rx_len) * sb;
            const void *dsamp = (const char *)samples + f0 * frame_stride * sb;
            if (space != COFDM_DEVICE_IN) {
                CU_TRY(cudaMemcpyAsync(h->pipe_in[s].p, dsamp, in_bytes, cudaMemcpyHostToDevice, st));
                dsamp = h->pipe_in[s].p;
            }
            if (int rc = launch_rx(h, st, dsamp, fmt, n, frame_stride, (uint8_t *)h->pipe_out[s].p, h->amb_dev, t, sync_less, s)) return rc;
            CU_TRY(cudaMemcpyAsync(bytes + f0 * bpf, h->pipe_out[s].p, n * bpf, cudaMemcpyDeviceToHost, st));
        }
        return COFDM_OK;
    };
    const int rc_chunks = run_chunks();
    for (int i = 0; i < kPipe; i++) {
        const cudaError_t e = cudaStreamSynchronize(h->pipe_stream[i]);
        if (e != cudaSuccess && rc_chunks == COFDM_OK) return fail(COFDM_ERR_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e));
    }
    if (rc_chunks) return rc_chunks;
    if (ambiguous) {
        unsigned long long a = 0;
        CU_TRY(cudaMemcpy(&a, h->amb_dev, sizeof a, cudaMemcpyDeviceToHost));
        *ambiguous += a;
    }
    return COFDM_OK;
}

int cofdm_t2sin_metric(cofdm_t *h, const void *samples, int fmt, size_t n_samples, size_t start, float *rel, int space) {
    if (!h || !samples || !rel || (fmt != COFDM_CF32 && fmt != COFDM_CI16)) return fail(COFDM_ERR_ARG, "cofdm_t2sin_metric: bad argument");
    if (set_device(h)) return COFDM_ERR_CUDA;
    if (start > n_samples || h->P.t2sin_size <= 0) return fail(COFDM_ERR_ARG, "start beyond the capture");
    const size_t n_blocks = (n_samples - start) / (size_t)h->P.t2sin_size;        // Frame.hpp:152
    if (n_blocks == 0) return COFDM_OK;
    if (space == COFDM_DEVICE) {
        Timed t(h);
        return launch_t2(h, h->stream, samples, fmt, start, n_blocks, rel);
    }
    const size_t sb = sample_bytes(fmt);
    CU_TRY(h->scratch_b.reserve(n_blocks * sizeof(float)));
    const void *dsamp = samples;
    if (space != COFDM_DEVICE_IN) {
        CU_TRY(h->scratch_a.reserve(n_samples * sb));
        CU_TRY(cudaMemcpyAsync(h->scratch_a.p, samples, n_samples * sb, cudaMemcpyHostToDevice, h->stream));
        dsamp = h->scratch_a.p;
    }
    if (int rc = launch_t2(h, h->stream, dsamp, fmt, start, n_blocks, (float *)h->scratch_b.p)) return rc;
    CU_TRY(cudaMemcpyAsync(rel, h->scratch_b.p, n_blocks * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return COFDM_OK;
}

int cofdm_find_t2sin(cofdm_t *h, const void *samples, int fmt, size_t n_samples, size_t start, long long *pos, int space) {
    if (!h || !samples || !pos || (fmt != COFDM_CF32 && fmt != COFDM_CI16)) return fail(COFDM_ERR_ARG, "cofdm_find_t2sin: bad argument");
    if (set_device(h)) return COFDM_ERR_CUDA;
    *pos = COFDM_NOT_FOUND_T2SIN;
    if (start > n_samples || h->P.t2sin_size <= 0) return COFDM_OK;
    const size_t n_blocks = (n_samples - start) / (size_t)h->P.t2sin_size;
    if (n_blocks == 0) return COFDM_OK;
    const void *dsamp = samples;
    const size_t sb = sample_bytes(fmt);
    if (space == COFDM_HOST) {
        CU_TRY(h->scratch_a.reserve(n_samples * sb));
        CU_TRY(cudaMemcpyAsync(h->scratch_a.p, samples, n_samples * sb, cudaMemcpyHostToDevice, h->stream));
        dsamp = h->scratch_a.p;
    }
    CU_TRY(h->scratch_b.reserve(n_blocks * sizeof(float)));
    CU_TRY(cudaMemsetAsync(h->pos_dev, 0xff, sizeof(unsigned long long), h->stream));
    if (int rc = launch_t2(h, h->stream, dsamp, fmt, start, n_blocks, (float *)h->scratch_b.p)) return rc;
    first_above_kernel<<<(unsigned)((n_blocks + 255) / 256), 256, 0, h->stream>>>((const float *)h->scratch_b.p, (long long)n_blocks, h->P.t2_level,
                                                                                 (long long)start, h->P.t2sin_size, h->pos_dev);
    if (int rc = check_launch(h, "first_above")) return rc;
    unsigned long long v = 0;
    CU_TRY(cudaMemcpyAsync(&v, h->pos_dev, sizeof v, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    *pos = v == ~0ull ? COFDM_NOT_FOUND_T2SIN : (long long)v;
    return COFDM_OK;
}

int cofdm_preamble_search(cofdm_t *h, const void *samples, int fmt, size_t n_samples, const long long *starts,
                          size_t n_starts, float *cor, long long *first, int space) {
    if (!h || !samples || !starts || (fmt != COFDM_CF32 && fmt != COFDM_CI16)) return fail(COFDM_ERR_ARG, "cofdm_preamble_search: bad argument");
    if (set_device(h)) return COFDM_ERR_CUDA;
    if (n_starts == 0) return COFDM_OK;
    if (space == COFDM_DEVICE) {
        Timed t(h);
        return launch_pc(h, h->stream, samples, fmt, n_samples, starts, n_starts, cor, first);
    }
    const size_t sb = sample_bytes(fmt), ncor = (size_t)h->P.cor_size;
    const size_t o_first = n_starts * sizeof(long long), o_cor = 2 * o_first;
    CU_TRY(h->scratch_b.reserve(o_cor + n_starts * ncor * sizeof(float)));
    char *b = (char *)h->scratch_b.p;
    const void *dsamp = samples;
    if (space != COFDM_DEVICE_IN) {
        CU_TRY(h->scratch_a.reserve(n_samples * sb));
        CU_TRY(cudaMemcpyAsync(h->scratch_a.p, samples, n_samples * sb, cudaMemcpyHostToDevice, h->stream));
        dsamp = h->scratch_a.p;
    }
    CU_TRY(cudaMemcpyAsync(b, starts, o_first, cudaMemcpyHostToDevice, h->stream));
    if (int rc = launch_pc(h, h->stream, dsamp, fmt, n_samples, (const long long *)b, n_starts,
                           cor ? (float *)(b + o_cor) : nullptr, (long long *)(b + o_first))) return rc;
    if (first) CU_TRY(cudaMemcpyAsync(first, b + o_first, o_first, cudaMemcpyDeviceToHost, h->stream));
    if (cor) CU_TRY(cudaMemcpyAsync(cor, b + o_cor, n_starts * ncor * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return COFDM_OK;
}

// The host-sequenced form of the loop (one search launch + one 8-byte read-back per step).  Kept for
// configurations the device scanner does not cover and as an independent cross-check (env COFDM_STREAM_HOSTSEQ=1).
static int rx_stream_hostseq(cofdm_t *h, const int16_t *capture, size_t n_samples, size_t max_frames,
                             long long *pr_begin_abs, uint8_t *bytes, size_t *n_found) {
    const Params &P = h->P;
    const long long out_sz = P.frame_len, block = out_sz * h->T.rx_buf_size;   // SDR::rx_buf_size, sdr.hpp:141
    const long long ring = out_sz * (h->T.rx_buf_size + 1);                     // from_sdr_buf.size(), Frame.cpp:221
    const long long n_blocks = block > 0 ? (long long)n_samples / block : 0;
    if (n_blocks == 0 || max_frames == 0) return COFDM_OK;
    cudaStream_t st = h->stream;
    // device ring (int16 I,Q = 4 bytes per sample) + batch of frames awaiting demodulation
    const size_t batch_cap = 1024;
    CU_TRY(h->scratch_a.reserve((size_t)ring * 4));
    CU_TRY(h->scratch_b.reserve((size_t)ring / (size_t)std::max(16, P.t2sin_size) * sizeof(float) + 64));
    CU_TRY(h->pipe_in[0].reserve(batch_cap * (size_t)P.rx_len * 4));
    CU_TRY(h->pipe_out[0].reserve(batch_cap * (size_t)P.bytes_per_frame));
    CU_TRY(h->scratch_c.reserve(64));
    char *d_ring = (char *)h->scratch_a.p;
    CU_TRY(cudaMemsetAsync(d_ring, 0, (size_t)ring * 4, st));
    long long next_block = 0, cur_block = -1;
    auto buf_update = [&]() -> int {                                           // rx.cpp:73-91
        if (next_block >= n_blocks) return 0;
        if (cudaMemcpyAsync(d_ring + out_sz * 4, capture + 2 * next_block * block, (size_t)block * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
        cur_block = next_block++;
        return 1;
    };
    const long long threshold = ring - out_sz;                                // rx.cpp:116
    auto carry = [&]() {                                                       // rx.cpp:149-153 / 182-186
        return cudaMemcpyAsync(d_ring, d_ring + threshold * 4, (size_t)out_sz * 4, cudaMemcpyDeviceToDevice, st);
    };
    size_t found = 0, in_batch = 0, flushed = 0;
    auto flush = [&]() -> int {
        if (in_batch == 0) return COFDM_OK;
        RxTaps none{};
        if (int rc = launch_rx(h, st, h->pipe_in[0].p, COFDM_CI16, in_batch, (size_t)P.rx_len, (uint8_t *)h->pipe_out[0].p, nullptr, none)) return rc;
        if (bytes && cudaMemcpyAsync(bytes + flushed * P.bytes_per_frame, h->pipe_out[0].p, in_batch * P.bytes_per_frame, cudaMemcpyDeviceToHost, st) != cudaSuccess)
            return fail(COFDM_ERR_CUDA, "rx_stream: D2H");
        if (cudaStreamSynchronize(st) != cudaSuccess) return fail(COFDM_ERR_CUDA, "rx_stream: sync");
        flushed += in_batch;
        in_batch = 0;
        return COFDM_OK;
    };
    // T2SIN_FORM::find_t2sin on the ring from `start` (Frame.hpp:150-197), in windows of 64 blocks
    auto find_t2 = [&](long long start, long long *pos) -> int {
        *pos = -1;
        const long long t2n = P.t2sin_size, cycles = (ring - start) / t2n;
        for (long long c0 = 0; c0 < cycles; c0 += 64) {
            const long long nb = std::min<long long>(64, cycles - c0);
            if (cudaMemsetAsync(h->pos_dev, 0xff, sizeof(unsigned long long), st) != cudaSuccess) return fail(COFDM_ERR_CUDA, "memset");
            if (int rc = launch_t2(h, st, d_ring, COFDM_CI16, (size_t)(start + c0 * t2n), (size_t)nb, (float *)h->scratch_b.p)) return rc;
            first_above_kernel<<<1, 64, 0, st>>>((const float *)h->scratch_b.p, nb, P.t2_level, start + c0 * t2n, (int)t2n, h->pos_dev);
            if (int rc = check_launch(h, "first_above")) return rc;
            unsigned long long v = 0;
            if (cudaMemcpyAsync(&v, h->pos_dev, sizeof v, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
                return fail(COFDM_ERR_CUDA, "rx_stream: D2H");
            if (v != ~0ull) { *pos = (long long)v; return COFDM_OK; }
        }
        return COFDM_OK;
    };
    auto find_pre = [&](long long start, long long *first) -> int {           // Frame.cpp:338-378
        long long *d_start = (long long *)h->scratch_c.p, *d_first = d_start + 1;
        if (cudaMemcpyAsync(d_start, &start, sizeof start, cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(COFDM_ERR_CUDA, "H2D");
        if (int rc = launch_pc(h, st, d_ring, COFDM_CI16, (size_t)ring, d_start, 1, nullptr, d_first)) return rc;
        if (cudaMemcpyAsync(first, d_first, sizeof *first, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
            return fail(COFDM_ERR_CUDA, "rx_stream: D2H");
        return COFDM_OK;
    };
#define COFDM_UPDATE_OR_BREAK() { int u_ = buf_update(); if (u_ < 0) return fail(COFDM_ERR_CUDA, "rx_stream: H2D"); if (u_ == 0) break; }
    if (buf_update() <= 0) return COFDM_OK;                                    // rx.cpp:103-112
    long long pos = 0;
    const long long cycles = h->T.iterations;                                  // rx.cpp:124
    for (long long it = 0; it < cycles && found < max_frames; it++) {          // rx.cpp:126
        if (int rc = find_t2(pos, &pos)) return rc;                            // :133
        if (pos == -1) {                                                       // :137-145
            pos = out_sz;
            COFDM_UPDATE_OR_BREAK();
            continue;
        }
        if (pos >= threshold) {                                                // :147-156
            pos -= threshold;
            CU_TRY(carry());
            COFDM_UPDATE_OR_BREAK();
        }
        long long first = -10;
        if (int rc = find_pre(pos, &first)) return rc;
        const long long preamble_begin = first + 1;                            // :158
        if (preamble_begin < -2) { pos += (long long)P.ofdm_len * P.num_symb; continue; }   // :160-166
        pos = preamble_begin;                                                  // :168
        if (pos == -1) {                                                       // :170-178
            pos = out_sz;
            COFDM_UPDATE_OR_BREAK();
            continue;
        }
        if (pos >= threshold + P.t2sin_size) {                                 // :180-189
            pos -= threshold;
            CU_TRY(carry());
            COFDM_UPDATE_OR_BREAK();
        }
        // rx.cpp:192-196: the frame's rx_len samples go to the demodulation batch
        CU_TRY(cudaMemcpyAsync((char *)h->pipe_in[0].p + in_batch * (size_t)P.rx_len * 4, d_ring + pos * 4, (size_t)P.rx_len * 4, cudaMemcpyDeviceToDevice, st));
        if (pr_begin_abs) pr_begin_abs[found] = cur_block * block + pos - out_sz;
        pos += (long long)P.ofdm_len * P.num_symb;                             // :198
        found++;
        if (++in_batch == batch_cap) { if (int rc = flush()) return rc; }
    }
#undef COFDM_UPDATE_OR_BREAK
    if (int rc = flush()) return rc;
    *n_found = found;
    return COFDM_OK;
}

// merge of per-shard chains (same rule as c-ofdm_b200/stream.py::merge_shards): follow the earlier shard's chain
// (the true one) until it meets a preamble position the later shard also found, then switch to the later shard's list;
// a frame is owned by the shard whose block range contains its preamble.
static void merge_stream_shards(const std::vector<std::vector<long long>> &lists, const std::vector<long long> &own_end,
                                std::vector<long long> &out, size_t *unmerged) {
    std::vector<long long> carry;
    *unmerged = 0;
    for (size_t r = 0; r < lists.size(); r++) {
        const std::vector<long long> &pos = lists[r];
        size_t start = 0;
        if (!carry.empty()) {
            size_t k = carry.size(), j = 0;
            for (size_t a = 0; a < carry.size() && k == carry.size(); a++) {
                auto it = std::lower_bound(pos.begin(), pos.end(), carry[a]);
                if (it != pos.end() && *it == carry[a]) { k = a; j = (size_t)(it - pos.begin()); }
            }
            if (k < carry.size()) {
                out.insert(out.end(), carry.begin(), carry.begin() + (long)k);
                start = j;
            } else {
                (*unmerged)++;
                out.insert(out.end(), carry.begin(), carry.end());
                start = (size_t)(std::upper_bound(pos.begin(), pos.end(), carry.back()) - pos.begin());
            }
        }
        size_t n_own = start;
        for (size_t a = start; a < pos.size(); a++) if (pos[a] < own_end[r]) n_own = a + 1;
        out.insert(out.end(), pos.begin() + (long)start, pos.begin() + (long)n_own);
        carry.assign(pos.begin() + (long)n_own, pos.end());
    }
    out.insert(out.end(), carry.begin(), carry.end());
}

int cofdm_rx_stream_sharded(cofdm_t *h, const int16_t *capture, size_t n_samples, int space, int n_shards, size_t max_frames,
                            long long *pr_begin_abs, uint8_t *bytes, size_t *n_found, size_t *n_unmerged) {
    if (!h || !capture || !n_found || n_shards < 1 || (space != COFDM_HOST && space != COFDM_DEVICE && space != COFDM_DEVICE_IN))
        return fail(COFDM_ERR_ARG, "cofdm_rx_stream_sharded: bad argument");
    const bool bytes_on_device = space == COFDM_DEVICE;           // the payloads stay in device memory (4-byte aligned buffer)
    if (bytes_on_device && bytes && ((uintptr_t)bytes & 3)) return fail(COFDM_ERR_ARG, "cofdm_rx_stream_sharded: device byte buffer must be 4-byte aligned");
    *n_found = 0;
    if (n_unmerged) *n_unmerged = 0;
    if (set_device(h)) return COFDM_ERR_CUDA;
    const Params &P = h->P;
    if (!h->T.fused512_ok && !h->T.generic_ok) return fail(COFDM_ERR_UNSUPPORTED, "rx_stream: configuration outside both receive paths");
    // the device scanner is built for the fft-512 geometry with T2sin_size = 256; everything else (other sync-tone sizes, the
    // any-size / fft-4096 receive paths) runs the host-sequenced form of the same loop (one search launch per step)
    const bool scanner_ok = h->T.fused512_ok && P.t2sin_size == 256 && (P.pr_sin_len % 4) == 0 && (P.cor_size % 4) == 0 && h->T.rx_buf_size >= 1;
    static const bool force_hostseq = [] { const char *e = std::getenv("COFDM_STREAM_HOSTSEQ"); return e && std::atoi(e) != 0; }();
    if ((!scanner_ok || force_hostseq) && space == COFDM_HOST && n_shards == 1)
        return rx_stream_hostseq(h, capture, n_samples, max_frames, pr_begin_abs, bytes, n_found);
    if (!scanner_ok) return fail(COFDM_ERR_UNSUPPORTED, "rx_stream: the device scanner needs pr_sin_len and the lag count to be multiples of 4");
    const long long out_sz = P.frame_len, block = out_sz * h->T.rx_buf_size;
    const long long total_blocks = (long long)n_samples / block;
    if (total_blocks == 0 || max_frames == 0) return COFDM_OK;
    cudaStream_t st = h->stream;
    const unsigned *d_cap = reinterpret_cast<const unsigned *>(capture);
    const bool tm = h->timing;
    if (tm) { collect_rx_stage(h); for (float &v : h->stage_ms) v = 0.f; cudaEventRecord(h->sev[3], st); }
    if (space == COFDM_HOST) {
        CU_TRY(h->scratch_a.reserve((size_t)total_blocks * block * 4));
        CU_TRY(cudaMemcpyAsync(h->scratch_a.p, capture, (size_t)total_blocks * block * 4, cudaMemcpyHostToDevice, st));
        d_cap = reinterpret_cast<const unsigned *>(h->scratch_a.p);
    }
    if (tm) cudaEventRecord(h->sev[4], st);
    const long long ns = std::min<long long>(n_shards, total_blocks);
    std::vector<StreamShard> shards((size_t)ns);
    std::vector<long long> own_end((size_t)ns);
    const long long msg = (long long)P.ofdm_len * P.num_symb;
    long long max_per = 0;
    for (long long r = 0; r < ns; r++) {                       // cofdm_b200.dist.shard_range + one overlap block
        const long long base = total_blocks / ns, rem = total_blocks % ns;
        const long long b0 = r * base + std::min(r, rem), b1 = b0 + base + (r < rem ? 1 : 0);
        const long long b1x = std::min(total_blocks, b1 + (b1 < total_blocks ? 1 : 0));
        shards[(size_t)r] = StreamShard{b0 * block, b1x - b0, b1 - b0};
        own_end[(size_t)r] = b1 * block;
        max_per = std::max(max_per, (b1x - b0) * block / msg + 2);
    }
    if (ns == 1) max_per = std::min<long long>(max_per, (long long)max_frames);
    const size_t list_bytes = (size_t)ns * (size_t)max_per * sizeof(long long);
    CU_TRY(h->scratch_b.reserve(list_bytes + (size_t)ns * sizeof(int) + (size_t)ns * sizeof(StreamShard) + 64));
    long long *d_pos = (long long *)h->scratch_b.p;
    StreamShard *d_sh = (StreamShard *)((char *)h->scratch_b.p + list_bytes);
    int *d_cnt = (int *)(d_sh + ns);
    CU_TRY(cudaMemcpyAsync(d_sh, shards.data(), (size_t)ns * sizeof(StreamShard), cudaMemcpyHostToDevice, st));
    {
        Timed t(h);
        stream_scan_kernel<<<(unsigned)ns, kScanThreads, stream_scan_smem_bytes(P.cor_size, P.pr_sin_len), st>>>(
            P, d_cap, d_sh, (int)ns, h->T.rx_buf_size, (long long)h->T.iterations, d_pos, (int)max_per, d_cnt);
        if (int rc = check_launch(h, "stream_scan")) return rc;
    }
    if (tm) cudaEventRecord(h->sev[5], st);
    const auto wall = [] { return std::chrono::steady_clock::now(); };
    const auto t_merge0 = wall();
    std::vector<std::vector<long long>> lists((size_t)ns);
    {
        // ONE read-back of the shards' lists and counts into pinned memory, one synchronisation (a few hundred KB at most:
        // a frame occupies at least message.size samples of its shard)
        const size_t tail_bytes = (size_t)ns * sizeof(StreamShard) + (size_t)ns * sizeof(int);
        CU_TRY(h->pin.reserve(list_bytes + tail_bytes));
        CU_TRY(cudaMemcpyAsync(h->pin.p, d_pos, list_bytes + tail_bytes, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        const long long *all = (const long long *)h->pin.p;
        const int *cnt = (const int *)((const char *)h->pin.p + list_bytes + (size_t)ns * sizeof(StreamShard));
        for (long long r = 0; r < ns; r++)
            lists[(size_t)r].assign(all + (size_t)r * (size_t)max_per, all + (size_t)r * (size_t)max_per + (size_t)cnt[(size_t)r]);
    }
    std::vector<long long> merged;
    size_t unmerged = 0;
    merge_stream_shards(lists, own_end, merged, &unmerged);
    if (tm) {
        add_stage(h, COFDM_STAGE_UPLOAD, h->sev[3], h->sev[4]);
        add_stage(h, COFDM_STAGE_SCAN, h->sev[4], h->sev[5]);
        // list read-back + merge on the host: wall clock from the end of the scan (the stream was idle in between)
        float scan_ms = h->stage_ms[COFDM_STAGE_SCAN] + h->stage_ms[COFDM_STAGE_UPLOAD];
        const float since = std::chrono::duration<float, std::milli>(wall() - t_merge0).count();
        h->stage_ms[COFDM_STAGE_MERGE] += std::max(0.f, since - scan_ms);
    }
    if (merged.size() > max_frames) merged.resize(max_frames);
    if (n_unmerged) *n_unmerged = unmerged;
    const size_t found = merged.size();
    if (pr_begin_abs) std::copy(merged.begin(), merged.end(), pr_begin_abs);
    if (bytes && found) {
        // gather the frames found and demodulate them in batches
        const size_t batch_cap = 32768;
        CU_TRY(h->scratch_c.reserve(found * sizeof(long long)));
        CU_TRY(cudaMemcpyAsync(h->scratch_c.p, merged.data(), found * sizeof(long long), cudaMemcpyHostToDevice, st));
        CU_TRY(h->pipe_in[0].reserve(std::min(found, batch_cap) * (size_t)P.rx_len * 4));
        if (!bytes_on_device) CU_TRY(h->pipe_out[0].reserve(std::min(found, batch_cap) * (size_t)P.bytes_per_frame));
        RxTaps none{};
        // Frames that lie wholly inside the capture are demodulated IN PLACE (the rx kernels take the position list; int16
        // records need no alignment); a frame that sticks out of the capture (at most one at either end) goes through the
        // gather kernel, which pads it with zeros like the reference's calloc'd ring.
        static const bool inplace_on = [] { const char *e = std::getenv("COFDM_STREAM_INPLACE"); return !(e && std::atoi(e) == 0); }();
        const long long total_samples = total_blocks * block;
        size_t i0 = 0, i1 = found;
        while (i0 < found && merged[i0] < 0) i0++;
        while (i1 > i0 && merged[i1 - 1] + (long long)P.rx_len > total_samples) i1--;
        if (!inplace_on) i0 = i1 = found;
        auto run = [&](size_t f0, size_t n, bool inplace) -> int {
            uint8_t *dbytes = bytes_on_device ? bytes + f0 * (size_t)P.bytes_per_frame : (uint8_t *)h->pipe_out[0].p;
            const long long *dpos = (const long long *)h->scratch_c.p + f0;
            if (inplace) {
                if (int rc = launch_rx(h, st, d_cap, COFDM_CI16, n, (size_t)P.rx_len, dbytes, nullptr, none, 0, 0, dpos, (size_t)total_samples)) return rc;
            } else {
                if (tm) cudaEventRecord(h->sev[6], st);
                stream_gather_kernel<<<(unsigned)n, 256, 0, st>>>(d_cap, total_samples, dpos, (int)n, P.rx_len, (unsigned *)h->pipe_in[0].p);
                if (int rc = check_launch(h, "stream_gather")) return rc;
                if (tm) cudaEventRecord(h->sev[7], st);
                if (int rc = launch_rx(h, st, h->pipe_in[0].p, COFDM_CI16, n, (size_t)P.rx_len, dbytes, nullptr, none)) return rc;
            }
            if (tm) cudaEventRecord(h->sev[3], st);
            if (!bytes_on_device)
                CU_TRY(cudaMemcpyAsync(bytes + f0 * (size_t)P.bytes_per_frame, h->pipe_out[0].p, n * (size_t)P.bytes_per_frame, cudaMemcpyDeviceToHost, st));
            if (tm) cudaEventRecord(h->sev[4], st);
            CU_TRY(cudaStreamSynchronize(st));
            if (tm) { if (!inplace) add_stage(h, COFDM_STAGE_GATHER, h->sev[6], h->sev[7]); collect_rx_stage(h); add_stage(h, COFDM_STAGE_D2H, h->sev[3], h->sev[4]); }
            return COFDM_OK;
        };
        for (size_t f0 = 0; f0 < i0; f0 += batch_cap) if (int rc = run(f0, std::min(batch_cap, i0 - f0), false)) return rc;
        for (size_t f0 = i0; f0 < i1; f0 += batch_cap) if (int rc = run(f0, std::min(batch_cap, i1 - f0), true)) return rc;
        for (size_t f0 = i1; f0 < found; f0 += batch_cap) if (int rc = run(f0, std::min(batch_cap, found - f0), false)) return rc;
    }
    *n_found = found;
    return COFDM_OK;
}

int cofdm_rx_stream(cofdm_t *h, const int16_t *capture, size_t n_samples, size_t max_frames,
                    long long *pr_begin_abs, uint8_t *bytes, size_t *n_found) {
    return cofdm_rx_stream_sharded(h, capture, n_samples, COFDM_HOST, 1, max_frames, pr_begin_abs, bytes, n_found, nullptr);
}

int cofdm_i16_to_cf32(cofdm_t *h, const int16_t *in, float *out, size_t n, int space) {
    if (!h || !in || !out) return fail(COFDM_ERR_ARG, "cofdm_i16_to_cf32: bad argument");
    if (set_device(h)) return COFDM_ERR_CUDA;
    if (n == 0) return COFDM_OK;
    const unsigned *din = (const unsigned *)in; float2 *dout = (float2 *)out;
    if (space == COFDM_HOST) {
        CU_TRY(h->scratch_a.reserve(n * 4)); CU_TRY(h->scratch_b.reserve(n * 8));
        CU_TRY(cudaMemcpyAsync(h->scratch_a.p, in, n * 4, cudaMemcpyHostToDevice, h->stream));
        din = (const unsigned *)h->scratch_a.p; dout = (float2 *)h->scratch_b.p;
    }
    {
        Timed t(h);
        i16_to_cf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(din, dout, (long long)n);
        if (int rc = check_launch(h, "i16_to_cf32")) return rc;
    }
    if (space == COFDM_HOST) {
        CU_TRY(cudaMemcpyAsync(out, dout, n * 8, cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
    }
    return COFDM_OK;
}

int cofdm_ring_load(cofdm_t *h, const int16_t *ring_host, size_t n_samples, const int16_t **ring_dev) {
    if (!h || !ring_host || !ring_dev) return fail(COFDM_ERR_ARG, "cofdm_ring_load: bad argument");
    if (set_device(h)) return COFDM_ERR_CUDA;
    // (+ one frame of slack: the searches read up to cor_size + pr_sin_len samples past their start, like the reference)
    CU_TRY(h->ring.reserve((n_samples + (size_t)h->P.frame_len) * 4));
    CU_TRY(cudaMemcpyAsync(h->ring.p, ring_host, n_samples * 4, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(cudaMemsetAsync((char *)h->ring.p + n_samples * 4, 0, (size_t)h->P.frame_len * 4, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));                     // the caller may overwrite its ring right away
    *ring_dev = (const int16_t *)h->ring.p;
    return COFDM_OK;
}

// ncclAllReduce resolved at run time from the libnccl.so.2 of the process: the library itself links the CUDA runtime only
namespace {
typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int /*ncclDataType_t*/, int /*ncclRedOp_t*/, void * /*ncclComm_t*/, cudaStream_t);
nccl_allreduce_fn resolve_nccl_allreduce() {
    static nccl_allreduce_fn fn = [] {
        void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW);
        return lib ? (nccl_allreduce_fn)dlsym(lib, "ncclAllReduce") : (nccl_allreduce_fn) nullptr;
    }();
    return fn;
}
}  // namespace

int cofdm_allreduce_counters(cofdm_t *h, void *nccl_comm, unsigned long long *sum_counters, size_t n_sum, double *max_values, size_t n_max) {
    if (!h || !nccl_comm || (n_sum && !sum_counters) || (n_max && !max_values)) return fail(COFDM_ERR_ARG, "cofdm_allreduce_counters: bad argument");
    if (set_device(h)) return COFDM_ERR_CUDA;
    nccl_allreduce_fn ar = resolve_nccl_allreduce();
    if (!ar) return fail(COFDM_ERR_UNSUPPORTED, "cofdm_allreduce_counters: libnccl.so.2 (ncclAllReduce) not found in this process");
    CU_TRY(h->coll.reserve((n_sum + n_max) * 8 + 16));
    unsigned long long *ds = (unsigned long long *)h->coll.p;
    double *dm = (double *)(ds + n_sum);
    if (n_sum) CU_TRY(cudaMemcpyAsync(ds, sum_counters, n_sum * 8, cudaMemcpyHostToDevice, h->stream));
    if (n_max) CU_TRY(cudaMemcpyAsync(dm, max_values, n_max * 8, cudaMemcpyHostToDevice, h->stream));
    const int kNcclUint64 = 5, kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;      // nccl.h: ncclDataType_t / ncclRedOp_t
    if (n_sum) if (int rc = ar(ds, ds, n_sum, kNcclUint64, kNcclSum, nccl_comm, h->stream)) return fail(COFDM_ERR_CUDA, "ncclAllReduce(sum) failed: " + std::to_string(rc));
    if (n_max) if (int rc = ar(dm, dm, n_max, kNcclFloat64, kNcclMax, nccl_comm, h->stream)) return fail(COFDM_ERR_CUDA, "ncclAllReduce(max) failed: " + std::to_string(rc));
    if (n_sum) CU_TRY(cudaMemcpyAsync(sum_counters, ds, n_sum * 8, cudaMemcpyDeviceToHost, h->stream));
    if (n_max) CU_TRY(cudaMemcpyAsync(max_values, dm, n_max * 8, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return COFDM_OK;
}

}  // extern "C"
