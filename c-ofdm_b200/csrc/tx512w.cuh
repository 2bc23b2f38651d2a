// tx512w.cuh -- the transmit chain for the fft-512 geometry: ONE WARP PER OFDM SYMBOL.
//
// Reference chain: FRAME_FORM::write + get / get_int16 (Frame.cpp:235-256) -> OFDM_FORM::write (:185-198: body after the CP slot,
// the last cp_size samples copied in front) -> FFT_FORM::write (:54-70: data points and pilots onto the grid, backward FFT,
// / sqrt(fft_size)) -> Modulation::mod (modulation.cpp:39-50: bits -> constellation table).
//
// One CTA per frame.  Warp s (< num_symb) builds message symbol s; the last warp copies the frame-invariant sync tone + preamble.
// A symbol's warp: its 32 * modType payload bytes -> the warp's shared memory (one coalesced load); the 16 grid points of lane l
// (bins l + 32 n1; only n1 = 0..4 and 11..15 can be used by the sub-carrier map) built branch-free from a per-lane descriptor
// table; the backward transform as conj(FFT(conj G)) with warp_fft512 (radix 16 x 16 x 2, ONE shared-memory exchange private to
// the warp); the output sits in registers at sample n = k1 + 16 i + {0, 128, 256, 384}: it is scaled, converted to the wire
// format (cf32, or int16 truncated toward zero like Frame.cpp:252) and stored -- body and cyclic prefix -- as a LINEAR image of
// the symbol into the warp's region, which leaves the SM as ONE TMA bulk store (cp.async.bulk.global.shared::cta).  Frame
// buffers that are not 16-byte aligned are served by the same image and plain coalesced stores.
// No CTA-wide barrier anywhere: the warps of a frame share nothing.
#pragma once
#include "compat.cuh"
#include "params.h"
#include "fft512w.cuh"

namespace cofdmk {

// the warp's region: [0, 5120) symbol image (aliases the FFT exchange [0, 4672)), payload staging at [4672, 4672 + 272)
constexpr int kTxwRegion = 5120;
constexpr int kTxwPayOff = kFft512wBytes;
static_assert(kTxwPayOff + 32 * 8 + 16 <= kTxwRegion, "payload staging must fit behind the exchange region");
constexpr int kTxwSlots = 10;        // n1 = 0..4, 11..15: the grid rows a used bin can sit in (bins 1..132 and 380..511)
COFDM_HD constexpr int txw_n1(int slot) { return slot < 5 ? slot : slot + 6; }
COFDM_HD int tx512w_threads(int num_symb) { return 32 * (num_symb + 1); }
COFDM_HD size_t tx512w_smem_bytes(int num_symb) { return (size_t)num_symb * kTxwRegion; }

// grid point of one slot: null, pilot or the constellation point of MOD payload bits -- conjugated (the backward transform is
// evaluated as conj(FFT(conj G))).  d: 16-bit descriptor, data index (0..255) | 0x4000 null | 0x8000 pilot.  Branch-free.
template <int MOD>
COFDM_DEV float2 txw_point(const Params &P, const uint8_t *pl, unsigned d) {
    const unsigned di = d & 0xffu;
    const unsigned bit = di * MOD, b0 = bit >> 3;
    unsigned w;
    if ((8 % MOD) != 0) w = ((unsigned)pl[b0] << 8) | pl[b0 + 1];                    // a 6-bit symbol may straddle two bytes (staging is padded)
    else w = (unsigned)pl[b0] << 8;
    const unsigned sym = (w >> (16 - MOD - (bit & 7))) & ((1u << MOD) - 1u);
    const float2 c = __ldg(&P.constell[sym]);                                        // Frame.cpp:59-62 + modulation.cpp:39-50
    const bool isdata = (d & 0xc000u) == 0u;
    const float alt = (d & 0x8000u) ? P.pilot_ampl : 0.f;                            // Frame.cpp:55-57
    return make_float2(isdata ? c.x : alt, isdata ? -c.y : 0.f);
}

template <int MOD>
COFDM_DEV void txw_points(const Params &P, const uint8_t *pl, const uint4 da, const uint4 db, float2 (&v)[16]) {
    const unsigned dw[5] = {da.x, da.y, da.z, da.w, db.x};
#pragma unroll
    for (int n1 = 5; n1 < 11; n1++) v[n1] = make_float2(0.f, 0.f);
#pragma unroll
    for (int sl = 0; sl < kTxwSlots; sl++) v[txw_n1(sl)] = txw_point<MOD>(P, pl, (dw[sl >> 1] >> (16 * (sl & 1))) & 0xffffu);
}

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
// BULK: the frame buffer is 16-byte aligned -> one TMA bulk store per symbol; otherwise plain stores from the same image
template <int FMT, bool BULK, int MAXW>
__global__ void __launch_bounds__(32 * (MAXW + 1), MAXW <= 8 ? COFDM_TXW_MINB : 1)
tx512w_kernel(const Params P, const uint8_t *__restrict__ payload, int n_frames, void *__restrict__ frames) {
    COFDM_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int frame = blockIdx.x;
    if (frame >= n_frames) return;
    const int ns = P.num_symb;
    const size_t sample_bytes = (FMT == kCI16) ? 4 : 8;
    char *fout = reinterpret_cast<char *>(frames) + (size_t)frame * P.frame_len * sample_bytes;

    if (warp == ns) {
        // T2SIN tone + preamble are constants of the configuration (Frame.cpp:228-229)
        const bool wide = (reinterpret_cast<uintptr_t>(fout) & 15) == 0;
        const int n_const = P.t2sin_size + P.pf_size;
        for (int i = 2 * lane; i < n_const; i += 64) {
            const float2 a = i < P.t2sin_size ? __ldg(&P.t2_tone[i]) : __ldg(&P.preamble_td[i - P.t2sin_size]);
            const float2 b = i + 1 < P.t2sin_size ? __ldg(&P.t2_tone[i + 1]) : __ldg(&P.preamble_td[i + 1 - P.t2sin_size]);
            store_sample_pair<FMT>(fout, i, a, b, P.mult, wide);
        }
        return;
    }
    char *region = reinterpret_cast<char *>(smem_raw) + (size_t)warp * kTxwRegion;
    uint8_t *pl = reinterpret_cast<uint8_t *>(region + kTxwPayOff);
    const int mod = P.mod_type, sym_bytes = 32 * mod;                   // num_data_subc = 256 points of `mod` bits
    {
        const uint8_t *src = payload + (size_t)frame * P.bytes_per_frame + (size_t)warp * sym_bytes;
        if ((reinterpret_cast<uintptr_t>(src) & 3) == 0) {
            for (int i = lane; i < sym_bytes / 4; i += 32) reinterpret_cast<unsigned *>(pl)[i] = __ldg(reinterpret_cast<const unsigned *>(src) + i);
        } else {
            for (int i = lane; i < sym_bytes; i += 32) pl[i] = __ldg(src + i);
        }
        if (lane < 4) reinterpret_cast<unsigned *>(pl + sym_bytes)[lane] = 0u;      // the byte a straddling 6-bit symbol reads past the end
    }
    const uint4 da = __ldg(&P.tx_desc[2 * lane]), db = __ldg(&P.tx_desc[2 * lane + 1]);
    __syncwarp();
    float2 v[16];
    switch (mod) {                                  // uniform: the symbol width becomes a compile-time constant
        case 1: txw_points<1>(P, pl, da, db, v); break;
        case 2: txw_points<2>(P, pl, da, db, v); break;
        case 4: txw_points<4>(P, pl, da, db, v); break;
        case 6: txw_points<6>(P, pl, da, db, v); break;
        default: txw_points<8>(P, pl, da, db, v); break;
    }
    float2 mn[8], ot[8];
    warp_fft512(v, make_float2(1.f, 0.f), reinterpret_cast<float2 *>(region), P.tw_fft, lane, mn, ot);
    // mn[i] = conj x[k1 + 16 i + (g ? 384 : 0)], ot[i] = conj x[k1 + 16 i + (g ? 128 : 256)] (unnormalised); the region is free again.
    // / sqrt(512) (Frame.cpp:66-68) and the conjugation in one packed multiply; body after the CP slot (Frame.cpp:191-192), the
    // last 128 samples (mn[] of the odd lanes) also into the CP slot (:196-197)
    const float sc = 0.04419417382415922028f;
    const float2 scj = make_float2(sc, -sc);
    const int k1 = lane >> 1, g = lane & 1;
    const int bm = 128 + k1 + (g ? 384 : 0), bo = 128 + k1 + (g ? 128 : 256);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float2 ym = p_mul(mn[i], scj), yo = p_mul(ot[i], scj);
        txw_put<FMT>(region, bm + 16 * i, ym, P.mult);
        txw_put<FMT>(region, bo + 16 * i, yo, P.mult);
        if (g) txw_put<FMT>(region, k1 + 16 * i, ym, P.mult);
    }
    char *dst = fout + (size_t)(P.t2sin_size + P.pf_size + warp * 640) * sample_bytes;
    if (BULK) {
        tma_store_fence();
        __syncwarp();
        if (lane == 0) {
            tma_store_1d(dst, region, 640u * (unsigned)sample_bytes);
            tma_store_commit_and_wait_read();
        }
    } else {
        __syncwarp();
        if (FMT == kCI16) {
            for (int i = lane; i < 640; i += 32) reinterpret_cast<unsigned *>(dst)[i] = reinterpret_cast<const unsigned *>(region)[i];
        } else {
            for (int i = lane; i < 640; i += 32) reinterpret_cast<float2 *>(dst)[i] = reinterpret_cast<const float2 *>(region)[i];
        }
    }
}

}  // namespace cofdmk
