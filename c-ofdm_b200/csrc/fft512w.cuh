// fft512w.cuh -- FFT-512 by ONE warp, natural (re, im) register pairs, radix 16 x 16 x 2.
//
// Replaces FFTW plans F1/F2 (OFDM/Frame.cpp:16-24) on the receive side.  Every lane owns ONE radix-16 butterfly per
// pass (16 complex values = 32 registers); all arithmetic is packed f32x2 on natural-layout complex numbers (compat.cuh:
// nadd / nmul / nadd_mj ...).  ONE exchange through shared memory (a 16 x 16 transpose inside each half-warp, 4672 bytes
// that belong to the warp alone: no CTA or team barrier, only __syncwarp) and one exchange of 8 values between neighbour
// lanes by warp shuffle.  The kernels that use it are bound by the shared-memory pipe, so the second shared-memory
// exchange of a radix 8 x 8 x 8 plan was the thing to remove (profiles/README.md, round 2).
//
// Index algebra (forward transform, unnormalised like FFTW's):  n = 32 n1 + l,  l = 2 m + g = lane,  k = k1 + 16 k2 + 256 k3
//   pass 1  A [k1; l]     = sum_n1 x[n]            W16^{n1 k1}    then * W512^{l k1}
//   pass 2  Y_g[k1, k2]   = sum_m  A'[k1; 2m + g]  W16^{m k2}     (k2 = 0..15), then Y_1 *= W32^{k2}
//   pass 3  X [k]         = Y_0[k1, k2] + (-1)^k3 Y_1[k1, k2]
// Lane maps:  pass 1: lane l holds x[l + 32 n1];  passes 2, 3: lane 2 k1 + g.
// The lanes with g = 1 negate their pass-1 outputs when m is odd: their pass-2 results then come out rotated by 8
// (register i holds k2 = i + 8 mod 16), so EVERY lane keeps registers 0..7 and sends registers 8..15 to lane ^ 1.
// On exit:   mn[i] = X[k1 + 16 i + (g ? 384 : 0)],   ot[i] = X[k1 + 16 i + (g ? 128 : 256)],   i = 0..7
// -- for the receiver's sub-carrier map (bins 1..132 and 380..511) mn[] holds used bins only and ot[] almost none.
// Exchange layout (float2 slots):  296 g + 18 k1 + m   (64-bit stores, 128-bit loads of (m, m + 1); bank-conflict free:
// profiles/scripts/bank_check.py enumerates every access).
#pragma once
#include "compat.cuh"

namespace cofdmk {

constexpr int kFft512wBytes = (296 + 288) * 8;   // bytes of the exchange region (16-byte aligned): 4672

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

// forward 4-point DFT of (a, b, c, d) -> (o0, o1, o2, o3); CJ: c arrives WITHOUT its factor -j (folded into the adds)
template <bool CJ = false>
COFDM_DEV void ndft4(float2 a, float2 b, float2 c, float2 d, float2 &o0, float2 &o1, float2 &o2, float2 &o3) {
    const float2 e0 = CJ ? nadd_mj(a, c) : nadd(a, c), e1 = CJ ? nadd_pj(a, c) : nsub(a, c);
    const float2 f0 = nadd(b, d), f1 = nsub(b, d);                              // odd part; its -j rides on the adds
    o0 = nadd(e0, f0); o2 = nsub(e0, f0);
    o1 = nadd_mj(e1, f1); o3 = nadd_pj(e1, f1);
}

// forward 16-point DFT in place (4 x 4: n = 4 n1 + n2, k = k1 + 4 k2): 64 adds + 8 twiddle products = 80 instructions
COFDM_DEV void ndft16(float2 (&v)[16]) {
    const float2 w1 = make_float2(0.92387953251128675613f, -0.38268343236508977173f);    // W16^1
    const float2 w3 = make_float2(0.38268343236508977173f, -0.92387953251128675613f);    // W16^3
    const float2 w9 = make_float2(-0.92387953251128675613f, 0.38268343236508977173f);    // W16^9
    float2 u[4][4];
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) ndft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2], u[n2][0], u[n2][1], u[n2][2], u[n2][3]);
    // twiddles W16^{n2 k1}; u[2][2] * W16^4 = -j u[2][2] is folded into the second stage
    u[1][1] = nmul(u[1][1], w1); u[1][2] = nmul_w8_1(u[1][2]); u[1][3] = nmul(u[1][3], w3);
    u[2][1] = nmul_w8_1(u[2][1]);                              u[2][3] = nmul_w8_3(u[2][3]);
    u[3][1] = nmul(u[3][1], w3); u[3][2] = nmul_w8_3(u[3][2]); u[3][3] = nmul(u[3][3], w9);
    ndft4(u[0][0], u[1][0], u[2][0], u[3][0], v[0], v[4], v[8], v[12]);
    ndft4(u[0][1], u[1][1], u[2][1], u[3][1], v[1], v[5], v[9], v[13]);
    ndft4<true>(u[0][2], u[1][2], u[2][2], u[3][2], v[2], v[6], v[10], v[14]);
    ndft4(u[0][3], u[1][3], u[2][3], u[3][3], v[3], v[7], v[11], v[15]);
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

// position of bin k after warp_fft512: lane, array (mn / ot), register
COFDM_HD constexpr int f512_g(int k) { return (k >> 7) & 1; }
COFDM_HD constexpr bool f512_main(int k) { return k < 128 || k >= 384; }
COFDM_HD constexpr int f512_lane(int k) { return 2 * (k & 15) + f512_g(k); }
COFDM_HD constexpr int f512_i(int k) { return (k & 127) >> 4; }

// W32^q = exp(-j 2 pi q / 32), q = 1..15 except 8, as literals (compile-time operands of the multiplies)
#define COFDM_W32_LIST(X)                                                                                          \
    X(1, 0.98078528040323044913f, -0.19509032201612826785f) X(2, 0.92387953251128675613f, -0.38268343236508977173f)  \
    X(3, 0.83146961230254523708f, -0.55557023301960222474f) X(4, 0.70710678118654752440f, -0.70710678118654752440f)  \
    X(5, 0.55557023301960222474f, -0.83146961230254523708f) X(6, 0.38268343236508977173f, -0.92387953251128675613f)  \
    X(7, 0.19509032201612826785f, -0.98078528040323044913f) X(9, -0.19509032201612826785f, -0.98078528040323044913f) \
    X(10, -0.38268343236508977173f, -0.92387953251128675613f) X(11, -0.55557023301960222474f, -0.83146961230254523708f) \
    X(12, -0.70710678118654752440f, -0.70710678118654752440f) X(13, -0.83146961230254523708f, -0.55557023301960222474f) \
    X(14, -0.92387953251128675613f, -0.38268343236508977173f) X(15, -0.98078528040323044913f, -0.19509032201612826785f)

// v[n1] = x[lane + 32 n1] on entry (n1 = 0..15).  p: an extra factor applied with the pass-1 twiddles (the per-sample CFO rotation
// of the rx chain contributes exp(-j 2 pi beta (128 + lane) / 512) there); pass make_float2(1, 0) for a plain transform.
// w512 = exp(-j 2 pi k / 512) table (global).  E: the warp's exchange region; the caller guarantees (by __syncwarp) that nobody
// still reads it.  On exit mn[] / ot[] as described at the top of this file; v[] is consumed.
// (second form: the caller supplies V = W512^lane and V8 = W512^{8 lane} itself, e.g. loaded once for many transforms)
COFDM_DEV void warp_fft512(float2 (&v)[16], float2 p, float2 *E, const float2 V, const float2 V8, int lane,
                           float2 (&mn)[8], float2 (&ot)[8]) {
    const int g = lane & 1, m = lane >> 1;
    ndft16(v);
    {   // pass-1 twiddles by recurrence: T[k1] = p V^k1, V = W512^lane; restarted at k1 = 8 from the table (V^8 = W512^{8 lane}),
        // so no factor carries more than eight roundings
        if ((lane & 3) == 3) p = make_float2(-p.x, -p.y);                       // g = 1, m odd (see above)
        float2 t = p;
        v[0] = nmul(v[0], t);
#pragma unroll
        for (int k1 = 1; k1 < 8; k1++) { t = nmul(t, V); v[k1] = nmul(v[k1], t); }
        t = nmul(p, V8);
        v[8] = nmul(v[8], t);
#pragma unroll
        for (int k1 = 9; k1 < 16; k1++) { t = nmul(t, V); v[k1] = nmul(v[k1], t); }
    }
    {
        float2 *w = E + 296 * g + m;
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) w[18 * k1] = v[k1];
    }
    __syncwarp();
    {
        const float4 *r = reinterpret_cast<const float4 *>(E + 296 * g + 18 * m);   // as reader: k1 = lane >> 1
#pragma unroll
        for (int mm = 0; mm < 8; mm++) {
            const float4 q = r[mm];
            v[2 * mm] = make_float2(q.x, q.y);
            v[2 * mm + 1] = make_float2(q.z, q.w);
        }
    }
    __syncwarp();
    ndft16(v);                                                                      // v[i] = Y_0[k1, i] (g = 0) / Y_1[k1, i + 8 mod 16] (g = 1)
    if (g) {
        // Y_1 *= W32^{k2}: register i holds k2 = i + 8 mod 16
        v[0] = make_float2(v[0].y, -v[0].x);                                        // W32^8 = -j
#define COFDM_X(q, re, im) v[(q + 8) & 15] = nmul(v[(q + 8) & 15], make_float2(re, im));
        COFDM_W32_LIST(COFDM_X)
#undef COFDM_X
    }
    const float sg = g ? -1.0f : 1.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float2 recv = make_float2(__shfl_xor_sync(0xffffffffu, v[8 + i].x, 1), __shfl_xor_sync(0xffffffffu, v[8 + i].y, 1));
        mn[i] = p_fma(v[i], p_bcast(sg), recv);                                     // g = 0: Y0 + W Y1 (k3 = 0);  g = 1: Y0 - W Y1 (k3 = 1)
        ot[i] = p_fma(recv, p_bcast(-sg), v[i]);                                    // g = 0: Y0 - W Y1 (k3 = 1);  g = 1: Y0 + W Y1 (k3 = 0)
    }
}

COFDM_DEV void warp_fft512(float2 (&v)[16], float2 p, float2 *E, const float2 *__restrict__ w512, int lane,
                           float2 (&mn)[8], float2 (&ot)[8]) {
    warp_fft512(v, p, E, __ldg(w512 + lane), __ldg(w512 + 8 * lane), lane, mn, ot);
}

}  // namespace cofdmk
