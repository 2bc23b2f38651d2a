// fft.cuh -- hand-written FFT building blocks (no cuFFT).
//
//  * dft2/4/5/8/16: in-register butterflies, forward (W = e^{-j2pi/R}) or inverse (conjugate).
//  * stockham_pass<R>: generic mixed-radix autosort pass over a thread group, used for the 256-point sync-tone
//    detector and the any-size path.
//  (the fft-512 kernels use the one-warp radix 16 x 16 x 2 transform of fft512w.cuh)
//
// These replace the reference's FFTW plans F1-F5 (OFDM/Frame.cpp:16-24,108-112,147-150;
// OFDM/Frame.hpp:289-295).  All transforms are unnormalised like FFTW's.
#pragma once
#include "compat.cuh"

namespace cofdmk {

// multiply by -j (forward W4) or +j (inverse)
template <bool INV> COFDM_DEV float2 mul_w4(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }
// multiply by W8^1 = e^{-+j pi/4}
template <bool INV> COFDM_DEV float2 mul_w8_1(float2 a) {
    const float r = 0.70710678118654752440f;
    return INV ? make_float2((a.x - a.y) * r, (a.x + a.y) * r) : make_float2((a.x + a.y) * r, (a.y - a.x) * r);
}
// multiply by W8^3 = e^{-+j 3pi/4}
template <bool INV> COFDM_DEV float2 mul_w8_3(float2 a) {
    const float r = 0.70710678118654752440f;
    return INV ? make_float2((-a.x - a.y) * r, (a.x - a.y) * r) : make_float2((a.y - a.x) * r, (-a.x - a.y) * r);
}
template <bool INV> COFDM_DEV float2 twid(float2 w) { return INV ? make_float2(w.x, -w.y) : w; }

// the same three rotations on a packed complex pair (two problems per instruction)
template <bool INV> COFDM_DEV pc mul_w4(pc a) {
    pc r;
    if (INV) { r.re = p_neg(a.im); r.im = a.re; } else { r.re = a.im; r.im = p_neg(a.re); }
    return r;
}
template <bool INV> COFDM_DEV pc mul_w8_1(pc a) {
    const float2 r = p_bcast(0.70710678118654752440f);
    pc o;
    if (INV) { o.re = p_mul(p_sub(a.re, a.im), r); o.im = p_mul(p_add(a.re, a.im), r); }
    else { o.re = p_mul(p_add(a.re, a.im), r); o.im = p_mul(p_sub(a.im, a.re), r); }
    return o;
}
template <bool INV> COFDM_DEV pc mul_w8_3(pc a) {
    const float2 r = p_bcast(0.70710678118654752440f), nr = p_bcast(-0.70710678118654752440f);
    pc o;
    if (INV) { o.re = p_mul(p_add(a.re, a.im), nr); o.im = p_mul(p_sub(a.re, a.im), r); }
    else { o.re = p_mul(p_sub(a.im, a.re), r); o.im = p_mul(p_add(a.re, a.im), nr); }
    return o;
}

template <bool INV> COFDM_DEV void dft2(float2 *v) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}

// v[k] = sum_n v[n] W4^{nk}
template <bool INV, class T> COFDM_DEV void dft4(T *v) {
    T e0 = cadd(v[0], v[2]), e1 = csub(v[0], v[2]);
    T o0 = cadd(v[1], v[3]), o1 = mul_w4<INV>(csub(v[1], v[3]));
    v[0] = cadd(e0, o0);
    v[1] = cadd(e1, o1);
    v[2] = csub(e0, o0);
    v[3] = csub(e1, o1);
}

template <bool INV> COFDM_DEV void dft5(float2 *v) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;   // cos(2pi/5), cos(4pi/5)
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;    // sin(2pi/5), sin(4pi/5)
    float2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
    float2 t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
    float2 x0 = cadd(v[0], cadd(t1, t2));
    float2 m1 = make_float2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
    float2 m2 = make_float2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
    float2 q1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    float2 q2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    // forward: X1 = m1 - j q1, X4 = m1 + j q1, X2 = m2 - j q2, X3 = m2 + j q2 ; inverse swaps signs
    float2 jq1 = mul_w4<INV>(q1), jq2 = mul_w4<INV>(q2);   // (-j q) forward, (+j q) inverse
    v[0] = x0;
    v[1] = cadd(m1, jq1);
    v[4] = csub(m1, jq1);
    v[2] = cadd(m2, jq2);
    v[3] = csub(m2, jq2);
}

// v[k] = sum_n v[n] W8^{nk}: radix-2 split (even/odd outputs) followed by two 4-point DFTs
template <bool INV, class T> COFDM_DEV void dft8(T *v) {
    T a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
    T a1 = cadd(v[1], v[5]), a5 = mul_w8_1<INV>(csub(v[1], v[5]));
    T a2 = cadd(v[2], v[6]), a6 = mul_w4<INV>(csub(v[2], v[6]));
    T a3 = cadd(v[3], v[7]), a7 = mul_w8_3<INV>(csub(v[3], v[7]));
    T b0 = cadd(a0, a2), b2 = csub(a0, a2), b1 = cadd(a1, a3), b3 = mul_w4<INV>(csub(a1, a3));
    T b4 = cadd(a4, a6), b6 = csub(a4, a6), b5 = cadd(a5, a7), b7 = mul_w4<INV>(csub(a5, a7));
    v[0] = cadd(b0, b1);
    v[4] = csub(b0, b1);
    v[2] = cadd(b2, b3);
    v[6] = csub(b2, b3);
    v[1] = cadd(b4, b5);
    v[5] = csub(b4, b5);
    v[3] = cadd(b6, b7);
    v[7] = csub(b6, b7);
}

// 16 = 4 x 4: n = 4*n1 + n2, k = k1 + 4*k2
template <bool INV> COFDM_DEV void dft16(float2 *v) {
    // W16^m, m = 1,2,3,4,6,9 (forward values; conjugated for the inverse by twid<INV>)
    const float2 w1 = make_float2(0.92387953251128675613f, -0.38268343236508977173f);
    const float2 w2 = make_float2(0.70710678118654752440f, -0.70710678118654752440f);
    const float2 w3 = make_float2(0.38268343236508977173f, -0.92387953251128675613f);
    const float2 w6 = make_float2(-0.70710678118654752440f, -0.70710678118654752440f);
    const float2 w9 = make_float2(-0.92387953251128675613f, 0.38268343236508977173f);
    float2 u[4][4];
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) {
        float2 c[4] = {v[n2], v[4 + n2], v[8 + n2], v[12 + n2]};
        dft4<INV>(c);
#pragma unroll
        for (int k1 = 0; k1 < 4; k1++) u[n2][k1] = c[k1];
    }
    u[1][1] = cmul(u[1][1], twid<INV>(w1));
    u[1][2] = cmul(u[1][2], twid<INV>(w2));
    u[1][3] = cmul(u[1][3], twid<INV>(w3));
    u[2][1] = cmul(u[2][1], twid<INV>(w2));
    u[2][2] = mul_w4<INV>(u[2][2]);
    u[2][3] = cmul(u[2][3], twid<INV>(w6));
    u[3][1] = cmul(u[3][1], twid<INV>(w3));
    u[3][2] = cmul(u[3][2], twid<INV>(w6));
    u[3][3] = cmul(u[3][3], twid<INV>(w9));
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) {
        float2 c[4] = {u[0][k1], u[1][k1], u[2][k1], u[3][k1]};
        dft4<INV>(c);
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) v[k1 + 4 * k2] = c[k2];
    }
}

// 10 = 2 x 5: X[k], X[k+5] = E[k] +- W10^k O[k], E/O = 5-point DFTs of the even/odd inputs
template <bool INV> COFDM_DEV void dft10(float2 *v) {
    float2 e[5] = {v[0], v[2], v[4], v[6], v[8]}, o[5] = {v[1], v[3], v[5], v[7], v[9]};
    dft5<INV>(e);
    dft5<INV>(o);
    const float2 w1 = make_float2(0.80901699437494742410f, -0.58778525229247312917f);   // W10^1 (forward)
    const float2 w2 = make_float2(0.30901699437494742410f, -0.95105651629515357212f);
    const float2 w3 = make_float2(-0.30901699437494742410f, -0.95105651629515357212f);
    const float2 w4 = make_float2(-0.80901699437494742410f, -0.58778525229247312917f);
    o[1] = cmul(o[1], twid<INV>(w1));
    o[2] = cmul(o[2], twid<INV>(w2));
    o[3] = cmul(o[3], twid<INV>(w3));
    o[4] = cmul(o[4], twid<INV>(w4));
#pragma unroll
    for (int k = 0; k < 5; k++) {
        v[k] = cadd(e[k], o[k]);
        v[k + 5] = csub(e[k], o[k]);
    }
}

// packed (two problems per instruction) 5- and 10-point butterflies
template <bool INV> COFDM_DEV void dft5(pc *v) {
    const float2 c1 = p_bcast(0.30901699437494742410f), c2 = p_bcast(-0.80901699437494742410f);
    const float2 s1 = p_bcast(0.95105651629515357212f), s2 = p_bcast(0.58778525229247312917f);
    const pc t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
    const pc t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
    pc x0 = cadd(v[0], cadd(t1, t2)), m1, m2, q1, q2;
    m1.re = p_fma(c2, t2.re, p_fma(c1, t1.re, v[0].re)); m1.im = p_fma(c2, t2.im, p_fma(c1, t1.im, v[0].im));
    m2.re = p_fma(c1, t2.re, p_fma(c2, t1.re, v[0].re)); m2.im = p_fma(c1, t2.im, p_fma(c2, t1.im, v[0].im));
    q1.re = p_fma(s2, t4.re, p_mul(s1, t3.re));          q1.im = p_fma(s2, t4.im, p_mul(s1, t3.im));
    q2.re = p_fms(s2, t3.re, p_mul(s1, t4.re));          q2.im = p_fms(s2, t3.im, p_mul(s1, t4.im));
    const pc jq1 = mul_w4<INV>(q1), jq2 = mul_w4<INV>(q2);
    v[0] = x0;
    v[1] = cadd(m1, jq1);
    v[4] = csub(m1, jq1);
    v[2] = cadd(m2, jq2);
    v[3] = csub(m2, jq2);
}
template <bool INV> COFDM_DEV void dft10(pc *v) {
    pc e[5] = {v[0], v[2], v[4], v[6], v[8]}, o[5] = {v[1], v[3], v[5], v[7], v[9]};
    dft5<INV>(e);
    dft5<INV>(o);
    o[1] = cmul(o[1], twid<INV>(make_float2(0.80901699437494742410f, -0.58778525229247312917f)));
    o[2] = cmul(o[2], twid<INV>(make_float2(0.30901699437494742410f, -0.95105651629515357212f)));
    o[3] = cmul(o[3], twid<INV>(make_float2(-0.30901699437494742410f, -0.95105651629515357212f)));
    o[4] = cmul(o[4], twid<INV>(make_float2(-0.80901699437494742410f, -0.58778525229247312917f)));
#pragma unroll
    for (int k = 0; k < 5; k++) {
        v[k] = cadd(e[k], o[k]);
        v[k + 5] = csub(e[k], o[k]);
    }
}

template <int R, bool INV> COFDM_DEV void dftR(float2 *v) {
    if (R == 2) dft2<INV>(v);
    else if (R == 4) dft4<INV>(v);
    else if (R == 5) dft5<INV>(v);
    else if (R == 8) dft8<INV>(v);
    else if (R == 10) dft10<INV>(v);
    else dft16<INV>(v);
}

// One Stockham autosort pass (decimation in time) of an n-point transform, radix R, where `ns` is
// the product of the radices already applied.  Butterflies j = tid, tid+nthr, ... < n/R.
// tw[k] = exp(-j*2*pi*k/n) (forward table, conjugated on the fly for INV).  in != out.
// NOWRAP: the caller guarantees (R-1)*(ns-1)*tstep < n, so the twiddle index needs no reduction mod n.
// PADSH > 0: element i lives at slot i + (i >> PADSH).  With PADSH = 3 the radix-8 scatter of the first passes
// (stride 8, then stride 64 + 1) and the consecutive gathers are all free of bank conflicts; buffers need n + n/8 slots.
// padded slot of element i.  PADSH 0: none; 1..31: i + (i >> PADSH); 64: i + 8 * (i >> 6).  The 256-point detector uses none for
// its input (consecutive reads), 4 for the pass-1 output (scatter 8 j + q and consecutive reads both conflict-free) and 64
// for the pass-2 output (scatter (j - k) 8 + k + 8 q and consecutive reads both conflict-free): enumerated in profiles/README.md.
template <int PADSH> COFDM_DEV int pad_slot(int i) { return PADSH == 64 ? i + ((i >> 6) << 3) : (PADSH > 0 ? i + (i >> PADSH) : i); }
template <int R, bool INV, bool NOWRAP = false, int PADSH = 0, int PADOUT = PADSH>
COFDM_DEV void stockham_pass(const float2 *in, float2 *out, int n, int ns, const float2 *tw, int tid, int nthr) {
    const int m = n / R;
    const int tstep = n / (ns * R);
    const bool pow2 = (n & (n - 1)) == 0 && (ns & (ns - 1)) == 0;   // masks instead of integer division (uniform)
    for (int j = tid; j < m; j += nthr) {
        const int k = pow2 ? (j & (ns - 1)) : j % ns;
        float2 v[R];
#pragma unroll
        for (int q = 0; q < R; q++) {
            v[q] = in[pad_slot<PADSH>(j + q * m)];
            if (q > 0 && ns > 1) {
                const int ti = q * k * tstep;
                v[q] = cmul(v[q], twid<INV>(__ldg(&tw[NOWRAP ? ti : (pow2 ? (ti & (n - 1)) : ti % n)])));
            }
        }
        dftR<R, INV>(v);
        const int o = (j - k) * R + k;              // (j / ns) * ns * R + k
#pragma unroll
        for (int q = 0; q < R; q++) out[pad_slot<PADOUT>(o + q * ns)] = v[q];
    }
}

// The same pass for TWO transforms at once (packed f32x2; operands in separate re / im planes of float2 pairs).
template <int R, bool INV, bool NOWRAP = false, int PADSH = 0, int PADOUT = PADSH>
COFDM_DEV void stockham_pass_pc(const float2 *in_re, const float2 *in_im, float2 *out_re, float2 *out_im, int n, int ns,
                                const float2 *tw, int tid, int nthr) {
    static_assert(R == 8 || R == 4, "packed passes exist for radix 4 and 8");
    const int m = n / R;
    const int tstep = n / (ns * R);
    for (int j = tid; j < m; j += nthr) {
        const int k = j % ns;
        pc v[R];
#pragma unroll
        for (int q = 0; q < R; q++) {
            v[q].re = in_re[pad_slot<PADSH>(j + q * m)];
            v[q].im = in_im[pad_slot<PADSH>(j + q * m)];
            if (q > 0 && ns > 1) {
                const int ti = q * k * tstep;
                v[q] = cmul(v[q], twid<INV>(__ldg(&tw[NOWRAP ? ti : ti % n])));
            }
        }
        if (R == 8) dft8<INV>(v); else dft4<INV>(v);
        const int o = (j - k) * R + k;
#pragma unroll
        for (int q = 0; q < R; q++) { out_re[pad_slot<PADOUT>(o + q * ns)] = v[q].re; out_im[pad_slot<PADOUT>(o + q * ns)] = v[q].im; }
    }
}

}  // namespace cofdmk
