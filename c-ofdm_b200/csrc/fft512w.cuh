// fft512w.cuh -- FFT-512 by ONE warp, natural (re, im) register pairs, radix 8 x 8 x 8.
//
// Replaces FFTW plans F1/F2 (OFDM/Frame.cpp:16-24) on the receive side.  Every lane owns TWO radix-8
// butterflies per pass (16 complex values = 32 registers); all arithmetic is packed f32x2 on natural-layout
// complex numbers (compat.cuh: nadd / nmul / nadd_mj ...), so a butterfly with its 7 twiddles costs 41
// instructions.  The two exchanges go through one 5152-byte region of shared memory that belongs to the
// warp alone: no CTA or team barrier, only __syncwarp.
//
// Index algebra (forward transform, unnormalised like FFTW's):
//   n = 64 n1 + 8 n2 + n3,   k = k1 + 8 k2 + 64 k3   (all digits 0..7)
//   pass 1  A [k1; n2,n3] = sum_n1 x[n]          W8^{n1 k1}   then * W512^{(8 n2 + n3) k1}
//   pass 2  B [k1,k2; n3] = sum_n2 A'[k1;n2,n3]  W8^{n2 k2}   then * W64^{n3 k2}
//   pass 3  X [k]         = sum_n3 B'[k1,k2;n3]  W8^{n3 k3}
// Lane maps (slot a / slot b of the lane):
//   pass 1  lane l        : t = 8 n2 + n3 = 2l / 2l + 1          inputs x[t + 64 n1]   (adjacent samples: 128-bit loads)
//   pass 2  lane l' = 4 k1 + j   : (k1, n3 = (j & 1) + 4 (j >> 1)) / (k1, n3 + 2)
//   pass 3  lane l'' = k2 + 8 a  : (k1 = 2a, k2) / (k1 = 2a + 1, k2)   outputs X[c0 + 64 k3] / X[c0 + 1 + 64 k3], c0 = 2a + 8 k2
// Exchanges: a lane WRITES its two slots with 64-bit stores into two planes (a 128-bit store would need the two values in
// four consecutive registers, which costs moves) and READS 128 bits = two neighbours of one plane.  Layouts in float2 slots,
// rows padded so that every access is base register + immediate and no phase of any access has a bank conflict
// (profiles/scripts/bank_check.py enumerates them all):
//   E1  plane A (t even) / B (t odd):   648 B' + 40 k1 + (t >> 1)            B' = 1 for plane B (plane B starts at slot 324)
//   E2  plane A (n3 in {0,1,4,5}) / B (n3 in {2,3,6,7}):   272 B' + 34 k2 + 4 k1 + (n3 & 1) + 2 (n3 >> 2)
#pragma once
#include "compat.cuh"

namespace cofdmk {

constexpr int kFft512wBytes = (324 + 320) * 8;   // bytes of the exchange region (16-byte aligned): 5152

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

// forward 5- and 10-point DFTs, natural layout (the 640-point coarse-CFO spectrum is 10 x 8 x 8)
COFDM_DEV void ndft5(float2 (&v)[5]) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;   // cos(2pi/5), cos(4pi/5)
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;    // sin(2pi/5), sin(4pi/5)
    const float2 t1 = nadd(v[1], v[4]), t2 = nadd(v[2], v[3]), t3 = nsub(v[1], v[4]), t4 = nsub(v[2], v[3]);
    const float2 x0 = nadd(v[0], nadd(t1, t2));
    const float2 m1 = p_fma(t2, p_bcast(c2), p_fma(t1, p_bcast(c1), v[0]));
    const float2 m2 = p_fma(t2, p_bcast(c1), p_fma(t1, p_bcast(c2), v[0]));
    const float2 q1 = p_fma(t4, p_bcast(s2), p_mul(t3, p_bcast(s1)));
    const float2 q2 = p_fma(t4, p_bcast(-s1), p_mul(t3, p_bcast(s2)));
    v[0] = x0;
    v[1] = nadd_mj(m1, q1);   // m1 - j q1
    v[4] = nadd_pj(m1, q1);
    v[2] = nadd_mj(m2, q2);
    v[3] = nadd_pj(m2, q2);
}
// X[k], X[k+5] = E[k] +- W10^k O[k]
COFDM_DEV void ndft10(float2 (&v)[10]) {
    float2 e[5] = {v[0], v[2], v[4], v[6], v[8]}, o[5] = {v[1], v[3], v[5], v[7], v[9]};
    ndft5(e);
    ndft5(o);
    o[1] = nmul(o[1], make_float2(0.80901699437494742410f, -0.58778525229247312917f));
    o[2] = nmul(o[2], make_float2(0.30901699437494742410f, -0.95105651629515357212f));
    o[3] = nmul(o[3], make_float2(-0.30901699437494742410f, -0.95105651629515357212f));
    o[4] = nmul(o[4], make_float2(-0.80901699437494742410f, -0.58778525229247312917f));
#pragma unroll
    for (int k = 0; k < 5; k++) { v[k] = nadd(e[k], o[k]); v[k + 5] = nsub(e[k], o[k]); }
}
// the powers w^1..w^7 of a unit phasor with at most three roundings each
COFDM_DEV void npowers7(float2 w1, float2 (&w)[8]) {
    w[0] = make_float2(1.f, 0.f);
    w[1] = w1;
    w[2] = nmul(w1, w1); w[3] = nmul(w[2], w1); w[4] = nmul(w[2], w[2]);
    w[5] = nmul(w[4], w1); w[6] = nmul(w[3], w[3]); w[7] = nmul(w[4], w[3]);
}

// bins a lane holds after the transform: slot a = c0 + 64 k3, slot b = c0 + 1 + 64 k3
COFDM_DEV int fft512w_c0(int lane) { return 2 * (lane >> 3) + 8 * (lane & 7); }

// va[r] / vb[r] = x[2 lane + 64 r] / x[2 lane + 1 + 64 r] on entry (r = 0..7);
// on exit va[k3] / vb[k3] = X[c0 + 64 k3] / X[c0 + 1 + 64 k3].
// pa, pb: extra factors applied with the pass-1 twiddles (the per-sample CFO rotation of the rx chain contributes
// exp(-j 2 pi beta (128 + t) / 512) there); pass make_float2(1, 0) for a plain transform.
// w512 = exp(-j 2 pi k / 512) table (global, 16-byte aligned).
// E: the warp's exchange region; the caller guarantees (by __syncwarp) that nobody still reads it.
COFDM_DEV void warp_fft512(float2 (&va)[8], float2 (&vb)[8], float2 pa, float2 pb, float2 *E,
                           const float2 *__restrict__ w512, int lane) {
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
    {
        float2 *w = E + lane;
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) { w[40 * k1] = va[k1]; w[324 + 40 * k1] = vb[k1]; }
    }
    __syncwarp();
    const int k1p = lane >> 2, j = lane & 3;
    {
        const float4 *r = reinterpret_cast<const float4 *>(E + 324 * (j & 1) + 40 * k1p + 2 * (j >> 1));
#pragma unroll
        for (int n2 = 0; n2 < 8; n2++) {
            const float4 q = r[2 * n2];                                               // t = 8 n2 + n3 and t + 2
            va[n2] = make_float2(q.x, q.y);
            vb[n2] = make_float2(q.z, q.w);
        }
    }
    __syncwarp();
    ndft8(va);
    ndft8(vb);
    {
        float2 *w = E + 4 * k1p + j;
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) { w[34 * k2] = va[k2]; w[272 + 34 * k2] = vb[k2]; }
    }
    __syncwarp();
    {
        const float4 *r = reinterpret_cast<const float4 *>(E + 34 * (lane & 7) + 8 * (lane >> 3));   // k2, k1 = 2a
#pragma unroll
        for (int b = 0; b < 2; b++) {                                                 // k1 = 2a + b: slot a / b
            float2 (&v)[8] = b ? vb : va;
            const float4 q0 = r[2 * b], q1 = r[2 * b + 1], q2 = r[136 + 2 * b], q3 = r[136 + 2 * b + 1];
            v[0] = make_float2(q0.x, q0.y); v[1] = make_float2(q0.z, q0.w);
            v[4] = make_float2(q1.x, q1.y); v[5] = make_float2(q1.z, q1.w);
            v[2] = make_float2(q2.x, q2.y); v[3] = make_float2(q2.z, q2.w);
            v[6] = make_float2(q3.x, q3.y); v[7] = make_float2(q3.z, q3.w);
        }
    }
    __syncwarp();
    {   // pass-2 twiddles W64^{n3 k2}, applied on the reading side where both slots share k2: the seven powers of
        // w = W64^{k2} come from one 8-byte load and six products (<= 3 roundings) instead of seven 16-byte table loads
        const float2 w1 = __ldg(w512 + 8 * (lane & 7));                                // W512^{8 k2} = W64^{k2}
        const float2 w2 = nmul(w1, w1), w3 = nmul(w2, w1), w4 = nmul(w2, w2);
        const float2 w5 = nmul(w4, w1), w6 = nmul(w3, w3), w7 = nmul(w4, w3);
        va[1] = nmul(va[1], w1); vb[1] = nmul(vb[1], w1);
        va[2] = nmul(va[2], w2); vb[2] = nmul(vb[2], w2);
        va[3] = nmul(va[3], w3); vb[3] = nmul(vb[3], w3);
        va[4] = nmul(va[4], w4); vb[4] = nmul(vb[4], w4);
        va[5] = nmul(va[5], w5); vb[5] = nmul(vb[5], w5);
        va[6] = nmul(va[6], w6); vb[6] = nmul(vb[6], w6);
        va[7] = nmul(va[7], w7); vb[7] = nmul(vb[7], w7);
    }
    ndft8(va);
    ndft8(vb);
}

}  // namespace cofdmk
