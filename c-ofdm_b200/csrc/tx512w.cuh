// tx512w.cuh -- the transmit chain for the fft-512 geometry: ONE WARP PER OFDM SYMBOL, persistent CTAs.
//
// Reference chain: FRAME_FORM::write + get / get_int16 (Frame.cpp:235-256) -> OFDM_FORM::write (:185-198: body after the CP slot,
// the last cp_size samples copied in front) -> FFT_FORM::write (:54-70: data points and pilots onto the grid, backward FFT,
// / sqrt(fft_size)) -> Modulation::mod (modulation.cpp:39-50: bits -> constellation table).
//
// A CTA of num_symb warps walks over frames (frame = blockIdx.x, + gridDim.x, ...); warp s builds message symbol s of each.
// Once per CTA: the frame-invariant sync tone + preamble (Frame.cpp:228-229) are laid out in shared memory in the wire format
// and the constellation table is copied there.  Per frame and warp:
//   * the symbol's 32 * modType payload bytes -> the warp's shared memory; the load for the NEXT frame is issued before this
//     frame's arithmetic starts, so its DRAM latency is covered by a whole symbol of work;
//   * the 16 grid points of lane l (bins l + 32 n1; only rows n1 = 0..4 and 11..15 can be used by the sub-carrier map) are
//     built branch-free from a per-lane descriptor table (byte offset + shift of the symbol's bits, or null / pilot);
//   * the backward transform as conj(FFT(conj G)) with warp_fft512 (radix 16 x 16 x 2, ONE shared-memory exchange private to
//     the warp); afterwards lane 2 k1 + g holds the samples n = k1 + 16 i + {0 | 384} (mn[i]) and {256 | 128} (ot[i]);
//   * scaled, converted to the wire format (cf32, or int16 truncated toward zero like Frame.cpp:252) and stored as FOUR
//     128-sample blocks, consecutive blocks 64 bytes further apart than their size: in one store instruction the even lanes
//     write block 0 or 2, the odd lanes block 3 or 1, and the 64-byte skew puts them on disjoint banks;
//   * the symbol leaves the SM as FIVE TMA bulk stores (cp.async.bulk.global.shared::cta): the cyclic prefix is block 3 sent a
//     second time (Frame.cpp:196-197), so it costs no shared-memory store at all.  Warp 0 also sends the constant part.
// Frame buffers that are not 16-byte aligned are served by plain coalesced stores from the same images.
// No CTA-wide barrier inside the frame loop: the warps of a frame share nothing.
#pragma once
#include "compat.cuh"
#include "params.h"
#include "fft512w.cuh"

namespace cofdmk {

// the warp's region: the four image blocks (aliasing the FFT exchange [0, 4672)), payload staging at [4672, 4672 + 272)
constexpr int kTxwRegion = 5120;
constexpr int kTxwPayOff = kFft512wBytes;
static_assert(kTxwPayOff + 32 * 8 + 16 <= kTxwRegion, "payload staging must fit behind the exchange region");
constexpr int kTxwSkew = 64;
COFDM_HD constexpr int txw_block_stride(int sample_bytes) { return 128 * sample_bytes + kTxwSkew; }
static_assert(4 * txw_block_stride(8) <= kTxwPayOff, "the image blocks must not reach the payload staging");
constexpr int kTxwSlots = 10;        // n1 = 0..4, 11..15: the grid rows a used bin can sit in (bins 1..132 and 380..511)
COFDM_HD constexpr int txw_n1(int slot) { return slot < 5 ? slot : slot + 6; }
// CTA-wide area behind the warps' regions: constellation table (256 x float2), the lanes' descriptors (32 x 2 x uint4), the lanes'
// pass-1 twiddle seeds (32 x float4: W512^lane, W512^{8 lane}), then the constant image
constexpr int kTxwConstellOff = 0, kTxwDescOff = 256 * 8, kTxwSeedOff = kTxwDescOff + 64 * 16;
constexpr int kTxwConstOff = kTxwSeedOff + 32 * 16;
COFDM_HD int tx512w_threads(int num_symb) { return 32 * num_symb; }
COFDM_HD size_t tx512w_smem_bytes(int num_symb, int n_const /*t2sin_size + pf_size*/) {
    return (size_t)num_symb * kTxwRegion + kTxwConstOff + (size_t)n_const * 8;
}

// grid point of one slot: null, pilot or the constellation point of MOD payload bits -- conjugated (the backward transform is
// evaluated as conj(FFT(conj G))).  d: 16-bit descriptor, [7:0] byte offset of the symbol's bits in the staged payload,
// [11:8] right shift (of the byte, or of the 16-bit window when MOD does not divide 8), [14] null, [15] pilot.  Branch-free.
template <int MOD>
COFDM_DEV float2 txw_point(const uint8_t *pl, const float2 *ctab, float pilot_ampl, unsigned d) {
    const unsigned b0 = d & 0xffu, sh = (d >> 8) & 15u;
    unsigned w = pl[b0];
    if ((8 % MOD) != 0) w = (w << 8) | pl[b0 + 1];                                   // a 6-bit symbol may straddle two bytes (staging is padded)
    const float2 c = ctab[(w >> sh) & ((1u << MOD) - 1u)];                           // Frame.cpp:59-62 + modulation.cpp:39-50
    const bool isdata = (d & 0xc000u) == 0u;
    const float alt = (d & 0x8000u) ? pilot_ampl : 0.f;                              // Frame.cpp:55-57
    return make_float2(isdata ? c.x : alt, isdata ? -c.y : 0.f);
}

template <int MOD>
COFDM_DEV void txw_points(const uint8_t *pl, const float2 *ctab, float pilot_ampl, const uint4 da, const unsigned db, float2 (&v)[16]) {
    const unsigned dw[5] = {da.x, da.y, da.z, da.w, db};
#pragma unroll
    for (int n1 = 5; n1 < 11; n1++) v[n1] = make_float2(0.f, 0.f);
#pragma unroll
    for (int sl = 0; sl < kTxwSlots; sl++) v[txw_n1(sl)] = txw_point<MOD>(pl, ctab, pilot_ampl, (dw[sl >> 1] >> (16 * (sl & 1))) & 0xffffu);
}

// one sample in the wire format at slot idx of an image
template <int FMT>
COFDM_DEV void txw_put(char *img, int idx, float2 y, float mult) {
    if (FMT == kCI16) {
        // Frame.cpp:252: int16(trunc(re*mult)), int16(trunc(im*mult))
        const unsigned lo = (unsigned)(unsigned short)(short)__float2int_rz(y.x * mult), hi = (unsigned)(unsigned short)(short)__float2int_rz(y.y * mult);
        reinterpret_cast<unsigned *>(img)[idx] = lo | (hi << 16);
    } else {
        reinterpret_cast<float2 *>(img)[idx] = y;
    }
}

#ifndef COFDM_TXW_MINB
#define COFDM_TXW_MINB 4
#endif
// BULK: the frame buffer is 16-byte aligned -> TMA bulk stores; otherwise plain stores from the same images
template <int FMT, bool BULK, int MAXW>
__global__ void __launch_bounds__(32 * MAXW, MAXW <= 8 ? COFDM_TXW_MINB : 1)
tx512w_kernel(const Params P, const uint8_t *__restrict__ payload, int n_frames, void *__restrict__ frames) {
    COFDM_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int ns = P.num_symb;
    constexpr int SB = (FMT == kCI16) ? 4 : 8;
    constexpr int BS = txw_block_stride(SB);
    char *region = reinterpret_cast<char *>(smem_raw) + (size_t)warp * kTxwRegion;
    char *shared_area = reinterpret_cast<char *>(smem_raw) + (size_t)ns * kTxwRegion;
    float2 *ctab = reinterpret_cast<float2 *>(shared_area + kTxwConstellOff);
    char *cimg = shared_area + kTxwConstOff;
    const int n_const = P.t2sin_size + P.pf_size;
    const int mod = P.mod_type, sym_bytes = 32 * mod;                   // num_data_subc = 256 points of `mod` bits
    const int nthreads = 32 * ns;

    // ---- once per CTA: constellation table and the constant part of every frame, in the wire format ----
    for (int i = tid; i < (1 << mod); i += nthreads) ctab[i] = __ldg(&P.constell[i]);
    uint4 *sdesc = reinterpret_cast<uint4 *>(shared_area + kTxwDescOff);
    float4 *sseed = reinterpret_cast<float4 *>(shared_area + kTxwSeedOff);
    // (lane's first four words at sdesc[lane], its fifth at word lane behind them: both reads are conflict-free)
    unsigned *sdesc5 = reinterpret_cast<unsigned *>(sdesc + 32);
    for (int i = tid; i < 32; i += nthreads) { sdesc[i] = __ldg(&P.tx_desc[2 * i]); sdesc5[i] = __ldg(&P.tx_desc[2 * i + 1]).x; }
    for (int i = tid; i < 32; i += nthreads) {
        const float2 a = __ldg(P.tw_fft + i), b = __ldg(P.tw_fft + 8 * i);
        sseed[i] = make_float4(a.x, a.y, b.x, b.y);
    }
    for (int i = tid; i < n_const; i += nthreads) {
        const float2 a = i < P.t2sin_size ? __ldg(&P.t2_tone[i]) : __ldg(&P.preamble_td[i - P.t2sin_size]);
        txw_put<FMT>(cimg, i, a, P.mult);
    }
    if (BULK) tma_store_fence();
    __syncthreads();

    uint8_t *pl = reinterpret_cast<uint8_t *>(region + kTxwPayOff);
    const bool pay_al = ((reinterpret_cast<uintptr_t>(payload) | (unsigned)P.bytes_per_frame) & 3) == 0;   // every symbol's bytes start on a word
    const int nwords = sym_bytes >> 2;                                  // 8 .. 64 words per symbol
    // the symbol's payload words of the frame to come (lane i holds words i and i + 32)
    unsigned w0 = 0u, w1 = 0u;
    int frame = blockIdx.x;
    if (pay_al && frame < n_frames) {
        const unsigned *src = reinterpret_cast<const unsigned *>(payload + (size_t)frame * P.bytes_per_frame + (size_t)warp * sym_bytes);
        if (lane < nwords) w0 = __ldg(src + lane);
        if (lane + 32 < nwords) w1 = __ldg(src + lane + 32);
    }
    if (lane < 4) reinterpret_cast<unsigned *>(pl + sym_bytes)[lane] = 0u;          // the byte a straddling 6-bit symbol reads past the end
    for (; frame < n_frames; frame += gridDim.x) {
        char *fout = reinterpret_cast<char *>(frames) + (size_t)frame * P.frame_len * SB;
        if (pay_al) {
            if (lane < nwords) reinterpret_cast<unsigned *>(pl)[lane] = w0;
            if (lane + 32 < nwords) reinterpret_cast<unsigned *>(pl)[lane + 32] = w1;
            const int nf = frame + gridDim.x;
            if (nf < n_frames) {
                const unsigned *src = reinterpret_cast<const unsigned *>(payload + (size_t)nf * P.bytes_per_frame + (size_t)warp * sym_bytes);
                if (lane < nwords) w0 = __ldg(src + lane);
                if (lane + 32 < nwords) w1 = __ldg(src + lane + 32);
            }
        } else {
            const uint8_t *src = payload + (size_t)frame * P.bytes_per_frame + (size_t)warp * sym_bytes;
            for (int i = lane; i < sym_bytes; i += 32) pl[i] = __ldg(src + i);
        }
        __syncwarp();
        float2 v[16];
        const uint4 da = sdesc[lane];
        const unsigned db = sdesc5[lane];
        switch (mod) {                              // uniform: the symbol width becomes a compile-time constant
            case 1: txw_points<1>(pl, ctab, P.pilot_ampl, da, db, v); break;
            case 2: txw_points<2>(pl, ctab, P.pilot_ampl, da, db, v); break;
            case 4: txw_points<4>(pl, ctab, P.pilot_ampl, da, db, v); break;
            case 6: txw_points<6>(pl, ctab, P.pilot_ampl, da, db, v); break;
            default: txw_points<8>(pl, ctab, P.pilot_ampl, da, db, v); break;
        }
        if (BULK) {
            if (lane == 0) tma_store_wait_read();                                   // the previous frame's blocks have left (they alias the exchange)
            __syncwarp();
        }
        float2 mn[8], ot[8];
        const float4 seed = sseed[lane];
        warp_fft512(v, make_float2(1.f, 0.f), reinterpret_cast<float2 *>(region), make_float2(seed.x, seed.y), make_float2(seed.z, seed.w), lane, mn, ot);
        // mn[i] = conj x[k1 + 16 i + (g ? 384 : 0)], ot[i] = conj x[k1 + 16 i + (g ? 128 : 256)] (unnormalised); the exchange region
        // is free again.  / sqrt(512) (Frame.cpp:66-68) and the conjugation in one packed multiply.
        const float sc = 0.04419417382415922028f;
        const float2 scj = make_float2(sc, -sc);
        const int k1 = lane >> 1, g = lane & 1;
        char *bm = region + (g ? 3 : 0) * BS, *bo = region + (g ? 1 : 2) * BS;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            txw_put<FMT>(bm, k1 + 16 * i, p_mul(mn[i], scj), P.mult);
            txw_put<FMT>(bo, k1 + 16 * i, p_mul(ot[i], scj), P.mult);
        }
        char *dst = fout + (size_t)(n_const + warp * 640) * SB;
        if (BULK) {
            tma_store_fence();
            __syncwarp();
            if (lane == 0) {
                tma_store_1d(dst, region + 3 * BS, 128u * SB);                       // the cyclic prefix: the last 128 samples (Frame.cpp:196-197)
#pragma unroll
                for (int b = 0; b < 4; b++) tma_store_1d(dst + (size_t)(128 + 128 * b) * SB, region + b * BS, 128u * SB);
                if (warp == 0) tma_store_1d(fout, cimg, (unsigned)n_const * SB);
                tma_store_commit();
            }
        } else {
            __syncwarp();
            for (int i = lane; i < 640; i += 32) {
                const int n = i < 128 ? i + 384 : i - 128;                          // body sample behind wire sample i
                const char *s = region + (n >> 7) * BS + (n & 127) * SB;
                if (FMT == kCI16) reinterpret_cast<unsigned *>(dst)[i] = *reinterpret_cast<const unsigned *>(s);
                else reinterpret_cast<float2 *>(dst)[i] = *reinterpret_cast<const float2 *>(s);
            }
            for (int i = tid; i < n_const; i += nthreads) {
                if (FMT == kCI16) reinterpret_cast<unsigned *>(fout)[i] = reinterpret_cast<const unsigned *>(cimg)[i];
                else reinterpret_cast<float2 *>(fout)[i] = reinterpret_cast<const float2 *>(cimg)[i];
            }
            __syncwarp();
        }
    }
    if (BULK && lane == 0) tma_store_wait_read();                                   // shared memory must outlive the bulk copies that read it
}

}  // namespace cofdmk
