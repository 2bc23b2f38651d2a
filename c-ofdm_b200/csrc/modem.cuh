// modem.cuh -- QAM/BPSK map and hard demap, device side.
// Follows reference OFDM/modulation.cpp: natural-binary (not Gray) square QAM on [-1,1]^2 with the
// low mod/2 bits on I and the high bits on Q (:12-20), BPSK rotated by 5*pi/4 (:30-31), MSB-first
// bit grouping (:90-125), demap = clamp, scale, +0.5, truncate (:66-78), BPSK sign of re+im (:62-63).
#pragma once
#include "compat.cuh"

namespace cofdmk {

// extract `mod` bits starting at bit position `bitpos` (MSB-first numbering) of a byte stream of
// n_bytes bytes; bits beyond the end read as zero (bit_stream_converter's tail padding, :121-122)
COFDM_DEV int extract_bits(const uint8_t *bytes, int n_bytes, int bitpos, int mod) {
    const int b0 = bitpos >> 3, off = bitpos & 7;
    unsigned w = 0;
    if (b0 < n_bytes) w = (unsigned)bytes[b0] << 8;
    if ((8 % mod) != 0 && b0 + 1 < n_bytes) w |= (unsigned)bytes[b0 + 1];   // mod 1,2,4,8 never straddle a byte
    return (int)((w >> (16 - mod - off)) & ((1u << mod) - 1u));
}

// the same with the symbol width known at compile time (shifts and masks fold to constants)
template <int MOD>
COFDM_DEV int extract_bits_t(const uint8_t *bytes, int n_bytes, int bitpos) {
    const int b0 = bitpos >> 3, off = bitpos & 7;
    unsigned w = 0;
    if (b0 < n_bytes) w = (unsigned)bytes[b0] << 8;
    if ((8 % MOD) != 0 && b0 + 1 < n_bytes) w |= (unsigned)bytes[b0 + 1];
    return (int)((w >> (16 - MOD - off)) & ((1u << MOD) - 1u));
}

// run-time width, dispatched once to the compile-time variants (avoids an integer modulo per symbol)
COFDM_DEV int extract_bits_sw(const uint8_t *bytes, int n_bytes, int bitpos, int mod) {
    switch (mod) {
        case 1: return extract_bits_t<1>(bytes, n_bytes, bitpos);
        case 2: return extract_bits_t<2>(bytes, n_bytes, bitpos);
        case 4: return extract_bits_t<4>(bytes, n_bytes, bitpos);
        case 6: return extract_bits_t<6>(bytes, n_bytes, bitpos);
        case 8: return extract_bits_t<8>(bytes, n_bytes, bitpos);
        default: return extract_bits(bytes, n_bytes, bitpos, mod);
    }
}

// margin (in level units) inside which a hard decision is reported as boundary-ambiguous: an fp32
// pipeline and the reference's fp64 pipeline may legitimately land on different sides.
constexpr float kAmbigMargin = 2e-4f;

// per-modulation constants, computed once per thread
struct DemapK { int mod, qshift, lmax; float half; };
COFDM_DEV DemapK make_demapk(int mod) {
    DemapK k;
    k.mod = mod;
    k.qshift = mod >> 1;                                   // lq * L == lq << (mod/2)
    k.lmax = (1 << (mod >> 1)) - 1;                        // L - 1
    k.half = 0.5f * (float)k.lmax;                         // str_size_1 = 1/step = (L-1)/2
    return k;
}

// hard demap of one equalised point -> symbol value; `amb` is set when the point lies within
// kAmbigMargin of a decision boundary.  u = (clamp(v)+1)*half + 0.5 lies in [0.5, L-0.5], so its
// fractional part is near 0 or 1 exactly when v is near one of the L-1 interior boundaries.
COFDM_DEV int demap_point(float2 z, const DemapK &k, bool &amb) {
    if (k.mod == 1) {
        const float s = z.x + z.y;
        amb = fabsf(s) < kAmbigMargin;
        return s > 0.0f ? 1 : 0;
    }
    const float re = fminf(fmaxf(z.x, -1.0f), 1.0f);
    const float im = fminf(fmaxf(z.y, -1.0f), 1.0f);
    const float ui = fmaf(re + 1.0f, k.half, 0.5f);
    const float uq = fmaf(im + 1.0f, k.half, 0.5f);
    const int li = (int)ui, lq = (int)uq;                  // uint8_t(...) truncation
    const float fi = ui - (float)li, fq = uq - (float)lq;
    amb = fminf(fminf(fi, 1.0f - fi), fminf(fq, 1.0f - fq)) < kAmbigMargin;
    return (li | (lq << k.qshift)) & 0xff;
}
COFDM_DEV int demap_point(float2 z, int mod, bool &amb) { return demap_point(z, make_demapk(mod), amb); }

// the decision alone / the ambiguity test alone (the fused kernel only pays for the latter when asked to count)
COFDM_DEV int demap_fast(float2 z, const DemapK &k) {
    if (k.mod == 1) return z.x + z.y > 0.0f ? 1 : 0;
    // clamp to [-1,1], (v + 1) half + 0.5, truncate (modulation.cpp:60-75) == truncate v half + (half + 0.5) with the
    // level clamped to [0, 2 half] afterwards: one fma, one conversion and an integer clamp per component
    const int li = min(max(__float2int_rz(fmaf(z.x, k.half, k.half + 0.5f)), 0), k.lmax);
    const int lq = min(max(__float2int_rz(fmaf(z.y, k.half, k.half + 0.5f)), 0), k.lmax);
    return li | (lq << k.qshift);
}
COFDM_DEV bool demap_ambiguous(float2 z, const DemapK &k) {
    bool amb;
    demap_point(z, k, amb);
    return amb;
}

}  // namespace cofdmk
