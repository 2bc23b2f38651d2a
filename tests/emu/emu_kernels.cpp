// emu_kernels.cpp -- compiles the REAL kernel source (c-ofdm_b200/csrc/kernels.cuh) for the CPU
// thread emulator (cuda_emu.h) and exposes it to pytest through a C ABI.  TEST INFRASTRUCTURE ONLY;
// see the header of cuda_emu.h.  Built by tests/emu/build_emu.py into tests/emu/libcofdm_emu.so.
#define COFDM_EMU 1
#include "cuda_emu.h"
#include "kernels.cuh"
#include "big.cuh"
#include "stream.cuh"
#include "host_consts.hpp"

using namespace cofdmk;

struct EmuHandle {
    HostTables T;
    Params P;
};

static void wire(EmuHandle *h) {
    HostTables &T = h->T;
    h->P = T.p;
    Params &P = h->P;
    P.tw_fft = T.tw_fft.data();
    P.tw_pf = T.tw_pf.data(); P.tw_t2 = T.tw_t2.data(); P.t2_mask = T.t2_mask.data();
    P.t2_tone = T.t2_tone.data(); P.preamble_td = T.preamble_td.data(); P.matched = T.matched.data();
    P.mod_preamble = T.mod_preamble.data(); P.constell = T.constell[T.p.mod_type].data();
    P.bin_role = T.bin_role.data(); P.big_roles = T.big_roles.data(); P.big_eq = T.big_eq.data(); P.big_txd = T.big_txd.data(); P.bin_map = T.bin_map.data(); P.data_bin = T.data_bin.data(); P.pilot_bin = T.pilot_bin.data();
    P.lane_desc = T.lane_desc.data(); P.lane_aux = T.lane_aux.data(); P.acq_desc = T.acq_desc.data(); P.grid_lane = T.grid_lane.data(); P.tx_desc = T.tx_desc.data();
}

// the one-warp-per-frame acquire kernel (rx512n.cuh)
template <bool TAPS>
static void emu_acquire512w(const Params &P, const void *samples, int fmt, int use_tma, int n_frames, long long stride,
                            const RxTaps &taps, FrameScal *fs, int sync_less) {
    const dim3 grid((n_frames + kAcqwWarps - 1) / kAcqwWarps), block(32 * kAcqwWarps);
    const size_t sm = rx_acquire512w_smem_bytes();
    const size_t sbytes = fmt == kCI16 ? 4 : 8;
    const RxSrc rs{nullptr, (const char *)samples, (const char *)samples + ((size_t)(n_frames - 1) * (size_t)stride + (size_t)P.rx_len) * sbytes};
#define EMU_AQ(F, T) emu::launch(grid, block, sm, [&] { rx_acquire512w_kernel<F, T, TAPS>(P, samples, stride, n_frames, taps, fs, sync_less, rs); })
    if (fmt == kCI16) { if (use_tma) EMU_AQ(kCI16, true); else EMU_AQ(kCI16, false); }
    else { if (use_tma) EMU_AQ(kCF32, true); else EMU_AQ(kCF32, false); }
#undef EMU_AQ
}

// the one-warp-per-symbol demod kernel (rx512n.cuh), production instantiations
template <bool TAPS>
static void emu_demod512(const Params &P, const void *samples, int fmt, int use_tma, int n_frames, long long stride,
                         uint8_t *out, unsigned long long *amb, const RxTaps &taps, const FrameScal *fs, int sync_less) {
    const dim3 grid(n_frames), block(32 * P.num_symb);
    const size_t sm = rx_demod512_smem_bytes(P.num_symb);
    const size_t sbytes = fmt == kCI16 ? 4 : 8;
    const RxSrc rs{nullptr, (const char *)samples, (const char *)samples + ((size_t)(n_frames - 1) * (size_t)stride + (size_t)P.rx_len) * sbytes};
#define EMU_DM(F, T, MW, MD) emu::launch(grid, block, sm, [&] { rx_demod512_kernel<F, T, TAPS, MW, MD>(P, samples, stride, n_frames, out, amb, taps, sync_less, fs, rs); })
    // as launch_rx does: production instances specialised on QPSK / 16-QAM, everything else generic
#define EMU_DM_PICK(F, T) do { if (P.num_symb <= 8) { if (!TAPS && P.mod_type == 4) EMU_DM(F, T, 8, TAPS ? 0 : 4); else if (!TAPS && P.mod_type == 2) EMU_DM(F, T, 8, TAPS ? 0 : 2); else EMU_DM(F, T, 8, 0); } \
                               else EMU_DM(F, T, kMaxFusedSymb, 0); } while (0)
    if (fmt == kCI16) { if (use_tma) EMU_DM_PICK(kCI16, true); else EMU_DM_PICK(kCI16, false); }
    else { if (use_tma) EMU_DM_PICK(kCF32, true); else EMU_DM_PICK(kCF32, false); }
#undef EMU_DM_PICK
#undef EMU_DM
}

// the receive chain as launch_rx (cofdm_host.cu) runs it: acquire (one warp per frame) + demod (one warp per symbol)
template <bool TAPS>
static int emu_rx_chain(EmuHandle *h, const void *samples, int fmt, int use_tma, int n_frames, long long stride, int sync_less,
                        uint8_t *out, unsigned long long *amb, const RxTaps &taps) {
    if (!h->T.fused512_ok) return -1;
    const Params P = h->P;
    std::vector<FrameScal> fs(n_frames);
    if (!sync_less || taps.chan != nullptr) emu_acquire512w<TAPS>(P, samples, fmt, use_tma, n_frames, stride, taps, fs.data(), sync_less);
    emu_demod512<TAPS>(P, samples, fmt, use_tma, n_frames, stride, out, amb, taps, fs.data(), sync_less);
    if (taps.synced && taps.scal && !sync_less) emu::launch(dim3(n_frames), dim3(128), 0, [&] { rx_synced_fixup_kernel(P, n_frames, taps); });
    return 0;
}

extern "C" {

void *emu_create(const char *config_path) {
    try {
        auto *h = new EmuHandle{build_tables(parse_config_file(config_path)), {}};
        wire(h);
        return h;
    } catch (const std::exception &) { return nullptr; }
}
void emu_destroy(void *h) { delete (EmuHandle *)h; }
int emu_fused_ok(void *h) { return ((EmuHandle *)h)->T.fused512_ok ? 1 : 0; }

static int g_emu_pc_plain = 0;
void emu_set_pc_plain(int on) { g_emu_pc_plain = on; }
static int g_emu_tx_bulk = 1;
void emu_set_tx_bulk(int on) { g_emu_tx_bulk = on; }

int emu_rx_fused512_mode(void *hv, const void *samples, int fmt, int use_tma, int n_frames, long long stride, int sync_less,
                    uint8_t *out, unsigned long long *amb, float *scal, float2 *grid, float2 *chan,
                    float2 *constell, float2 *synced) {
    RxTaps taps{scal, grid, chan, constell, synced};
    return emu_rx_chain<true>((EmuHandle *)hv, samples, fmt, use_tma, n_frames, stride, sync_less, out, amb, taps);
}

int emu_rx_fused512(void *hv, const void *samples, int fmt, int use_tma, int n_frames, long long stride,
                    uint8_t *out, unsigned long long *amb, float *scal, float2 *grid, float2 *chan,
                    float2 *constell, float2 *synced) {
    return emu_rx_fused512_mode(hv, samples, fmt, use_tma, n_frames, stride, 0, out, amb, scal, grid, chan, constell, synced);
}

// the production instantiation (no taps; specialised on the modulation order where launch_rx does so)
int emu_rx_fused512_notaps(void *hv, const void *samples, int fmt, int use_tma, int n_frames, long long stride,
                           uint8_t *out, unsigned long long *amb) {
    RxTaps taps{};
    return emu_rx_chain<false>((EmuHandle *)hv, samples, fmt, use_tma, n_frames, stride, 0, out, amb, taps);
}

int emu_tx512(void *hv, const uint8_t *payload, int n_frames, void *frames, int fmt) {
    auto *h = (EmuHandle *)hv;
    if (!h->T.fused512_ok) return -1;
    const Params P = h->P;
    {
        // one warp per symbol (tx512w.cuh); g_emu_tx_bulk picks the bulk-store or the plain-store output stage
        const size_t smw = tx512w_smem_bytes(P.num_symb, P.t2sin_size + P.pf_size);
        const dim3 blk(tx512w_threads(P.num_symb)), grd(std::min(n_frames, 2));     // persistent CTAs: two of them walk over the frames
#define EMU_TXW(F, B) do { if (P.num_symb <= 8) emu::launch(grd, blk, smw, [&] { tx512w_kernel<F, B, 8>(P, payload, n_frames, frames); }); \
                           else emu::launch(grd, blk, smw, [&] { tx512w_kernel<F, B, kMaxFusedSymb>(P, payload, n_frames, frames); }); } while (0)
        if (fmt == kCI16) { if (g_emu_tx_bulk) EMU_TXW(kCI16, true); else EMU_TXW(kCI16, false); }
        else { if (g_emu_tx_bulk) EMU_TXW(kCF32, true); else EMU_TXW(kCF32, false); }
#undef EMU_TXW
        return 0;
    }
}

int emu_t2sin_metric(void *hv, const void *samples, int fmt, long long start, long long n_blocks, float *rel) {
    auto *h = (EmuHandle *)hv;
    const Params P = h->P;
    if (P.t2sin_size != 256) {
        const unsigned grid = (unsigned)((n_blocks + kT2AnyWarps - 1) / kT2AnyWarps);
        if (fmt == kCI16) emu::launch(dim3(grid), dim3(32 * kT2AnyWarps), t2sin_any_smem_bytes(P.t2sin_size), [&] { t2sin_metric_any_kernel<kCI16>(P, samples, start, n_blocks, rel); });
        else emu::launch(dim3(grid), dim3(32 * kT2AnyWarps), t2sin_any_smem_bytes(P.t2sin_size), [&] { t2sin_metric_any_kernel<kCF32>(P, samples, start, n_blocks, rel); });
        return 0;
    }
    if (n_blocks >= 64) {                            // as launch_t2 does: two blocks per warp
        const unsigned g2 = (unsigned)(((n_blocks + 1) / 2 + kT2PairWarps - 1) / kT2PairWarps);
        if (fmt == kCI16) emu::launch(dim3(g2), dim3(32 * kT2PairWarps), 0, [&] { t2sin_metric2_kernel<kCI16>(P, samples, start, n_blocks, rel); });
        else emu::launch(dim3(g2), dim3(32 * kT2PairWarps), 0, [&] { t2sin_metric2_kernel<kCF32>(P, samples, start, n_blocks, rel); });
        return 0;
    }
    const unsigned grid = (unsigned)((n_blocks + kT2WarpsPerCta - 1) / kT2WarpsPerCta);
    if (fmt == kCI16) emu::launch(dim3(grid), dim3(32 * kT2WarpsPerCta), 0, [&] { t2sin_metric_kernel<kCI16>(P, samples, start, n_blocks, rel); });
    else emu::launch(dim3(grid), dim3(32 * kT2WarpsPerCta), 0, [&] { t2sin_metric_kernel<kCF32>(P, samples, start, n_blocks, rel); });
    return 0;
}

int emu_preamble_corr(void *hv, const void *samples, int fmt, long long n_samples, const long long *starts,
                      int n_starts, float *cor, long long *first) {
    auto *h = (EmuHandle *)hv;
    const Params P = h->P;
    if ((P.pr_sin_len % 4) == 0 && (P.cor_size % 4) == 0 && !g_emu_pc_plain) {   // as launch_pc does
        const size_t sm4 = preamble_corr4_smem_bytes(P.cor_size, P.pr_sin_len);
        if (fmt == kCI16) emu::launch(dim3(n_starts), dim3(kPc4Threads), sm4, [&] { preamble_corr4_kernel<kCI16>(P, samples, n_samples, starts, n_starts, cor, first); });
        else emu::launch(dim3(n_starts), dim3(kPc4Threads), sm4, [&] { preamble_corr4_kernel<kCF32>(P, samples, n_samples, starts, n_starts, cor, first); });
        return 0;
    }
    const size_t sm = (size_t)(P.cor_size + 2 * P.pr_sin_len) * sizeof(float2);
    if (fmt == kCI16) emu::launch(dim3(n_starts), dim3(kPcThreads), sm, [&] { preamble_corr_kernel<kCI16>(P, samples, n_samples, starts, n_starts, cor, first); });
    else emu::launch(dim3(n_starts), dim3(kPcThreads), sm, [&] { preamble_corr_kernel<kCF32>(P, samples, n_samples, starts, n_starts, cor, first); });
    return 0;
}

// the device-side acquisition loop (stream.cuh): per-shard lists of absolute preamble positions
int emu_stream_scan(void *hv, const void *capture_i16, const long long *shard_first, const long long *shard_blocks, int n_shards,
                    long long *pos_out, int max_per_shard, int *count_out) {
    auto *h = (EmuHandle *)hv;
    const Params P = h->P;
    if (P.t2sin_size != 256 || (P.pr_sin_len % 4) || (P.cor_size % 4)) return -1;
    std::vector<StreamShard> sh(n_shards);
    for (int i = 0; i < n_shards; i++) sh[i] = StreamShard{shard_first[i], shard_blocks[i], shard_blocks[i]};   // (own = all: no early stop inside the overlap)
    emu::launch(dim3(n_shards), dim3(kScanThreads), stream_scan_smem_bytes(P.cor_size, P.pr_sin_len), [&] {
        stream_scan_kernel(P, (const unsigned *)capture_i16, sh.data(), n_shards, h->T.rx_buf_size, (long long)h->T.iterations,
                           pos_out, max_per_shard, count_out);
    });
    return 0;
}

int emu_mod(void *hv, int mod, const uint8_t *bytes, long long n_bytes, float2 *points, long long n_points) {
    auto *h = (EmuHandle *)hv;
    const float2 *table = h->T.constell[mod].data();
    emu::launch(dim3((unsigned)((n_points + 127) / 128)), dim3(128), 0, [&] { mod_kernel(table, mod, bytes, n_bytes, points, n_points); });
    return 0;
}

int emu_demod(void *, int mod, const float2 *points, long long n_points, uint8_t *bytes, long long n_bytes,
              unsigned long long *amb) {
    const long long groups = (n_points + 7) / 8;
    emu::launch(dim3((unsigned)((groups + 127) / 128)), dim3(128), 0, [&] { demod_kernel(mod, points, n_points, bytes, n_bytes, amb); });
    return 0;
}

int emu_generic_ok(void *h) { return ((EmuHandle *)h)->T.generic_ok ? 1 : 0; }

int emu_rx_generic_mode(void *hv, const void *samples, int fmt, int n_frames, long long stride, int sync_less, uint8_t *out,
                        unsigned long long *amb, float *scal, float2 *chan, float2 *constell);
int emu_rx_generic(void *hv, const void *samples, int fmt, int n_frames, long long stride, uint8_t *out,
                   unsigned long long *amb, float *scal, float2 *chan, float2 *constell) {
    return emu_rx_generic_mode(hv, samples, fmt, n_frames, stride, 0, out, amb, scal, chan, constell);
}
int emu_rx_generic_mode(void *hv, const void *samples, int fmt, int n_frames, long long stride, int sync_less, uint8_t *out,
                        unsigned long long *amb, float *scal, float2 *chan, float2 *constell) {
    auto *h = (EmuHandle *)hv;
    if (!h->T.generic_ok) return -1;
    const Params P = h->P;
    const size_t N = P.fft_size, L = P.ofdm_len, nsym = P.n_sym_rx;
    std::vector<GenFrame> gf(n_frames);
    std::vector<float2> spec((size_t)n_frames * nsym * N), pre((size_t)n_frames * P.pf_size);
    RxTaps taps{scal, nullptr, chan, constell, nullptr};
    const dim3 blk(kGenThreads);
    if (fmt == kCI16) {
        if (!sync_less) emu::launch(dim3(n_frames), blk, 2 * P.pf_size * sizeof(float2), [&] { gen_coarse_kernel<kCI16>(P, samples, stride, n_frames, gf.data()); });
        emu::launch(dim3(nsym, n_frames), blk, (L + N) * sizeof(float2), [&] { gen_symbol_kernel<kCI16>(P, samples, stride, n_frames, gf.data(), spec.data(), pre.data(), sync_less); });
    } else {
        if (!sync_less) emu::launch(dim3(n_frames), blk, 2 * P.pf_size * sizeof(float2), [&] { gen_coarse_kernel<kCF32>(P, samples, stride, n_frames, gf.data()); });
        emu::launch(dim3(nsym, n_frames), blk, (L + N) * sizeof(float2), [&] { gen_symbol_kernel<kCF32>(P, samples, stride, n_frames, gf.data(), spec.data(), pre.data(), sync_less); });
    }
    emu::launch(dim3(n_frames), blk, P.num_data_subc / 2 * sizeof(float) + 16, [&] { gen_chan_kernel<false>(P, n_frames, gf.data(), spec.data(), pre.data(), sync_less); });
    emu::launch(dim3(P.num_symb, n_frames), blk, 0, [&] { gen_demap_kernel(P, n_frames, gf.data(), spec.data(), out, amb, taps); });
    return 0;
}

static int g_emu_big_tx = 1, g_emu_big_lay = 1;
void emu_set_big_lay(int on) { g_emu_big_lay = on; }
int emu_big_lay(void *h) { return ((EmuHandle *)h)->T.p.big_lay; }
void emu_set_big_tx(int on) { g_emu_big_tx = on; }
int emu_big_ok(void *h) { return ((EmuHandle *)h)->T.big_ok ? 1 : 0; }

// the fft-4096 path (big.cuh) as launch_rx_big runs it; mode 0: acquisition by the any-size kernels + bridge, 1: big_acquire_kernel.
// One emulated block of num_symb * 256 threads stands for the cluster of big_demod_kernel.
int emu_rx_big(void *hv, const void *samples, int fmt, int use_tma, int n_frames, long long stride, int mode, uint8_t *out,
               unsigned long long *amb, float *scal, float2 *chan, float2 *constell) {
    auto *h = (EmuHandle *)hv;
    if (!h->T.big_ok) return -1;
    const Params P = h->P;
    const size_t N = P.fft_size, L = P.ofdm_len;
    RxTaps taps{scal, nullptr, chan, constell, nullptr};
    const bool want = scal || chan || constell;
    std::vector<FrameScal> fs(n_frames);
    if (mode == 0) {
        std::vector<GenFrame> gf(n_frames);
        std::vector<float2> spec((size_t)n_frames * N), pre((size_t)n_frames * L);
        const dim3 blk(kGenThreads);
        if (fmt == kCI16) {
            emu::launch(dim3(n_frames), blk, 2 * P.pf_size * sizeof(float2), [&] { gen_coarse_kernel<kCI16>(P, samples, stride, n_frames, gf.data()); });
            emu::launch(dim3(1, n_frames), blk, (L + N) * sizeof(float2), [&] { gen_symbol_kernel<kCI16, true>(P, samples, stride, n_frames, gf.data(), spec.data(), pre.data()); });
        } else {
            emu::launch(dim3(n_frames), blk, 2 * P.pf_size * sizeof(float2), [&] { gen_coarse_kernel<kCF32>(P, samples, stride, n_frames, gf.data()); });
            emu::launch(dim3(1, n_frames), blk, (L + N) * sizeof(float2), [&] { gen_symbol_kernel<kCF32, true>(P, samples, stride, n_frames, gf.data(), spec.data(), pre.data()); });
        }
        emu::launch(dim3(n_frames), blk, P.num_data_subc / 2 * sizeof(float) + 16, [&] { gen_chan_kernel<true>(P, n_frames, gf.data(), spec.data(), pre.data()); });
        emu::launch(dim3((n_frames + 127) / 128), dim3(128), 0, [&] { big_bridge_kernel(P, n_frames, gf.data(), fs.data(), taps); });
    } else {
        const dim3 g1(n_frames), b1(kBigThreads);
        const size_t sma = big_acquire_smem_bytes();
#define EMU_BACQ(F, T) do { if (want) emu::launch(g1, b1, sma, [&] { big_acquire_kernel<F, T, true>(P, samples, stride, n_frames, taps, fs.data()); }); \
                            else emu::launch(g1, b1, sma, [&] { big_acquire_kernel<F, T, false>(P, samples, stride, n_frames, taps, fs.data()); }); } while (0)
        if (fmt == kCI16) { if (use_tma) EMU_BACQ(kCI16, true); else EMU_BACQ(kCI16, false); }
        else { if (use_tma) EMU_BACQ(kCF32, true); else EMU_BACQ(kCF32, false); }
#undef EMU_BACQ
    }
    const dim3 grid(n_frames), blk((unsigned)P.num_symb * kBigThreads);
    const size_t sm = (size_t)P.num_symb * big_smem_bytes();
#define EMU_BIG(F, T) do { if (want) emu::launch(grid, blk, sm, [&] { big_demod_kernel<F, T, true, 0>(P, samples, stride, n_frames, out, amb, taps, fs.data()); }); \
                           else if (P.mod_type == 6 && P.big_lay && g_emu_big_lay) emu::launch(grid, blk, sm, [&] { big_demod_kernel<F, T, false, 6, true>(P, samples, stride, n_frames, out, amb, taps, fs.data()); }); \
                           else if (P.mod_type == 6) emu::launch(grid, blk, sm, [&] { big_demod_kernel<F, T, false, 6>(P, samples, stride, n_frames, out, amb, taps, fs.data()); }); \
                           else emu::launch(grid, blk, sm, [&] { big_demod_kernel<F, T, false, 0>(P, samples, stride, n_frames, out, amb, taps, fs.data()); }); } while (0)
    if (fmt == kCI16) { if (use_tma) EMU_BIG(kCI16, true); else EMU_BIG(kCI16, false); }
    else { if (use_tma) EMU_BIG(kCF32, true); else EMU_BIG(kCF32, false); }
#undef EMU_BIG
    return 0;
}

int emu_tx_generic(void *hv, const uint8_t *payload, int n_frames, void *frames, int fmt) {
    auto *h = (EmuHandle *)hv;
    if (!h->T.generic_ok) return -1;
    const Params P = h->P;
    if (h->T.big_ok && g_emu_big_tx) {
        const bool spec = P.mod_type == 6 && P.big_lay && g_emu_big_lay;
        const dim3 g(P.num_symb + 1, n_frames), b(kBigThreads);
        if (fmt == kCI16) { if (spec) emu::launch(g, b, big_tx_smem_bytes(), [&] { big_tx_kernel<kCI16, 6, true>(P, payload, n_frames, frames); });
                            else emu::launch(g, b, big_tx_smem_bytes(), [&] { big_tx_kernel<kCI16>(P, payload, n_frames, frames); }); }
        else { if (spec) emu::launch(g, b, big_tx_smem_bytes(), [&] { big_tx_kernel<kCF32, 6, true>(P, payload, n_frames, frames); });
               else emu::launch(g, b, big_tx_smem_bytes(), [&] { big_tx_kernel<kCF32>(P, payload, n_frames, frames); }); }
        return 0;
    }
    const size_t sm = 2 * (size_t)P.fft_size * sizeof(float2);
    if (fmt == kCI16) emu::launch(dim3(P.num_symb + 1, n_frames), dim3(kGenThreads), sm, [&] { gen_tx_kernel<kCI16>(P, payload, n_frames, frames); });
    else emu::launch(dim3(P.num_symb + 1, n_frames), dim3(kGenThreads), sm, [&] { gen_tx_kernel<kCF32>(P, payload, n_frames, frames); });
    return 0;
}

}  // extern "C"
