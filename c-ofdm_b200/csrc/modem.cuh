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
    if (b0 + 1 < n_bytes) w |= (unsigned)bytes[b0 + 1];
    return (int)((w >> (16 - mod - off)) & ((1u << mod) - 1u));
}

// margin (in level units) inside which a hard decision is reported as boundary-ambiguous: an fp32
// pipeline and the reference's fp64 pipeline may legitimately land on different sides.
constexpr float kAmbigMargin = 2e-4f;

// hard demap of one equalised point -> symbol value; `amb` is set when the point lies within
// kAmbigMargin of a decision boundary.
COFDM_DEV int demap_point(float2 z, int mod, bool &amb) {
    if (mod == 1) {
        const float s = z.x + z.y;
        amb = fabsf(s) < kAmbigMargin;
        return s > 0.0f ? 1 : 0;
    }
    const int L = 1 << (mod >> 1);
    const float half = 0.5f * (float)(L - 1);              // str_size_1 = 1/step = (L-1)/2
    const float re = fminf(fmaxf(z.x, -1.0f), 1.0f);
    const float im = fminf(fmaxf(z.y, -1.0f), 1.0f);
    const float ui = (re + 1.0f) * half + 0.5f;
    const float uq = (im + 1.0f) * half + 0.5f;
    const int li = (int)ui, lq = (int)uq;                  // uint8_t(...) truncation
    const float ri = rintf(ui), rq = rintf(uq);
    amb = (fabsf(ui - ri) < kAmbigMargin && ri >= 1.0f && ri <= (float)(L - 1)) ||
          (fabsf(uq - rq) < kAmbigMargin && rq >= 1.0f && rq <= (float)(L - 1));
    return (li | (lq * L)) & 0xff;
}

}  // namespace cofdmk
