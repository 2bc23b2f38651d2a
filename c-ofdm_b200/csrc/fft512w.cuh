// fft512w.cuh -- FFT-512 by ONE warp, natural (re, im) register pairs, radix 8 x 8 x 8.
//
// Replaces FFTW plans F1/F2 (OFDM/Frame.cpp:16-24) on the receive side.  Every lane owns TWO radix-8
// butterflies per pass (16 complex values = 32 registers); all arithmetic is packed f32x2 on natural-layout
// complex numbers (compat.cuh: nadd / nmul / nadd_mj ...), so a butterfly with its 7 twiddles costs 41
// instructions.  The two exchanges go through one 512-slot float2 region of shared memory that belongs to the
// warp alone: no CTA or team barrier, only __syncwarp.  Every exchange access is 128 bits wide (two values of
// the lane's butterfly pair sit next to each other) and bank-conflict free by XOR swizzles
// (profiles/scripts/bank_check.py enumerates every access of every phase).
//
// Index algebra (forward transform, unnormalised like FFTW's):
//   n = 64 n1 + 8 n2 + n3,   k = k1 + 8 k2 + 64 k3   (all digits 0..7)
//   pass 1  A [k1; n2,n3] = sum_n1 x[n]          W8^{n1 k1}   then * W512^{(8 n2 + n3) k1}
//   pass 2  B [k1,k2; n3] = sum_n2 A'[k1;n2,n3]  W8^{n2 k2}   then * W64^{n3 k2}
//   pass 3  X [k]         = sum_n3 B'[k1,k2;n3]  W8^{n3 k3}
// Lane maps (slot a / slot b of the lane):
//   pass 1  lane l        : t = 8 n2 + n3 = 2l / 2l + 1          inputs x[t + 64 n1]   (adjacent samples: 128-bit loads)
//   pass 2  lane l' = 4 k1 + j   : (k1, n3 = 2j) / (k1, n3 = 2j + 1)
//   pass 3  lane l'' = k2 + 8 a  : (k1 = 2a, k2) / (k1 = 2a + 1, k2)   outputs X[c0 + 64 k3] / X[c0 + 1 + 64 k3], c0 = 2a + 8 k2
// Exchange layouts (float2 slots):
//   E1(k1, t)      = 64 k1 + (t ^ ((k1 & 1) << 3))
//   E2(k1, k2, n3) = 64 k1 + 8 (k2 ^ (k1 & 1)) + 2 ((n3 >> 1) ^ ((k2 >> 1) & 3)) + (n3 & 1)
#pragma once
#include "compat.cuh"

namespace cofdmk {

constexpr int kFft512wSlots = 512;      // float2 slots of the exchange region (16-byte aligned)

// multiply by W8^1 = (1 - j)/sqrt2 and W8^3 = (-1 - j)/sqrt2 (forward), natural layout
COFDM_DEV float2 nmul_w8_1(float2 a) { return nscale(p_add(a, make_float2(a.y, -a.x)), 0.70710678118654752440f); }
COFDM_DEV float2 nmul_w8_3(float2 a) { return nscale(p_add(make_float2(a.y, -a.x), make_float2(-a.x, -a.y)), 0.70710678118654752440f); }

// forward 8-point DFT in place: v[k] = sum_n v[n] W8^{nk}; the three multiplications by -j ride on the adds
COFDM_DEV void ndft8(float2 (&v)[8]) {
    const float2 a0 = nadd(v[0], v[4]), a4 = nsub(v[0], v[4]);
    const float2 a1 = nadd(v[1], v[5]), a5 = nmul_w8_1(nsub(v[1], v[5]));
    const float2 a2 = nadd(v[2], v[6]), d26 = nsub(v[2], v[6]);                 // a6 = -j d26
    const float2 a3 = nadd(v[3], v[7]), a7 = nmul_w8_3(nsub(v[3], v[7]));
    const float2 b0 = nadd(a0, a2), b2 = nsub(a0, a2), b1 = nadd(a1, a3), d13 = nsub(a1, a3);          // b3 = -j d13
    const float2 b4 = nadd_mj(a4, d26), b6 = nadd_pj(a4, d26), b5 = nadd(a5, a7), d57 = nsub(a5, a7);   // b7 = -j d57
    v[0] = nadd(b0, b1);      v[4] = nsub(b0, b1);
    v[2] = nadd_mj(b2, d13);  v[6] = nadd_pj(b2, d13);
    v[1] = nadd(b4, b5);      v[5] = nsub(b4, b5);
    v[3] = nadd_mj(b6, d57);  v[7] = nadd_pj(b6, d57);
}

COFDM_DEV int fft512w_e1(int k1, int t) { return 64 * k1 + (t ^ ((k1 & 1) << 3)); }
COFDM_DEV int fft512w_e2(int k1, int k2, int n3) {
    return 64 * k1 + 8 * (k2 ^ (k1 & 1)) + ((((n3 >> 1) ^ ((k2 >> 1) & 3)) << 1) | (n3 & 1));
}
// bins a lane holds after the transform: slot a = c0 + 64 k3, slot b = c0 + 1 + 64 k3
COFDM_DEV int fft512w_c0(int lane) { return 2 * (lane >> 3) + 8 * (lane & 7); }

// va[r] / vb[r] = x[2 lane + 64 r] / x[2 lane + 1 + 64 r] on entry (r = 0..7);
// on exit va[k3] / vb[k3] = X[c0 + 64 k3] / X[c0 + 1 + 64 k3].
// pa, pb: extra factors applied with the pass-1 twiddles (the per-sample CFO rotation of the rx chain contributes
// exp(-j 2 pi beta (128 + t) / 512) there); pass make_float2(1, 0) for a plain transform.
// w512 = exp(-j 2 pi k / 512) table (global, 16-byte aligned), tw2 = [8][8] exp(-j 2 pi n3 k2 / 64) (global or shared).
// E: the warp's exchange region; the caller guarantees (by __syncwarp) that nobody still reads it.
COFDM_DEV void warp_fft512(float2 (&va)[8], float2 (&vb)[8], float2 pa, float2 pb, float2 *E,
                           const float2 *__restrict__ w512, const float2 *__restrict__ tw2, int lane) {
    ndft8(va);
    ndft8(vb);
    {   // pass-1 twiddles by recurrence: T[k1] = p * V^k1, V = W512^t  (seven roundings at most: ~4e-7 relative)
        const float4 v4 = __ldg(reinterpret_cast<const float4 *>(w512) + lane);     // W512^{2 lane}, W512^{2 lane + 1}
        const float2 Va = make_float2(v4.x, v4.y), Vb = make_float2(v4.z, v4.w);
        float2 ta = pa, tb = pb;
        va[0] = nmul(va[0], ta);
        vb[0] = nmul(vb[0], tb);
#pragma unroll
        for (int k1 = 1; k1 < 8; k1++) {
            ta = nmul(ta, Va);
            tb = nmul(tb, Vb);
            va[k1] = nmul(va[k1], ta);
            vb[k1] = nmul(vb[k1], tb);
        }
    }
    float4 *E4 = reinterpret_cast<float4 *>(E);
    {
        const int t = 2 * lane;
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) E4[fft512w_e1(k1, t) >> 1] = make_float4(va[k1].x, va[k1].y, vb[k1].x, vb[k1].y);
    }
    __syncwarp();
    const int k1p = lane >> 2, j = lane & 3;
    {
#pragma unroll
        for (int n2 = 0; n2 < 8; n2++) {
            const float4 q = E4[fft512w_e1(k1p, 8 * n2 + 2 * j) >> 1];
            va[n2] = make_float2(q.x, q.y);
            vb[n2] = make_float2(q.z, q.w);
        }
    }
    __syncwarp();
    ndft8(va);
    ndft8(vb);
    {
        const float4 *t4 = reinterpret_cast<const float4 *>(tw2);
#pragma unroll
        for (int k2 = 1; k2 < 8; k2++) {
            const float4 w = t4[k2 * 4 + j];                                          // W64^{2j k2}, W64^{(2j+1) k2}
            va[k2] = nmul(va[k2], make_float2(w.x, w.y));
            vb[k2] = nmul(vb[k2], make_float2(w.z, w.w));
        }
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) E4[fft512w_e2(k1p, k2, 2 * j) >> 1] = make_float4(va[k2].x, va[k2].y, vb[k2].x, vb[k2].y);
    }
    __syncwarp();
    {
        const int k2 = lane & 7, a = lane >> 3;
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const float4 q = E4[fft512w_e2(2 * a, k2, 2 * jj) >> 1];
            va[2 * jj] = make_float2(q.x, q.y);
            va[2 * jj + 1] = make_float2(q.z, q.w);
            const float4 r = E4[fft512w_e2(2 * a + 1, k2, 2 * jj) >> 1];
            vb[2 * jj] = make_float2(r.x, r.y);
            vb[2 * jj + 1] = make_float2(r.z, r.w);
        }
    }
    __syncwarp();
    ndft8(va);
    ndft8(vb);
}

}  // namespace cofdmk
