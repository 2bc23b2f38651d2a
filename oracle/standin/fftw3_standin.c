/* fftw3_standin.c -- double-precision mixed-radix FFT behind the FFTW3 plan API subset
 * declared in ./fftw3.h.  TEST INFRASTRUCTURE ONLY (oracle / CPU baseline); never linked
 * into the product library.
 *
 * Algorithm: Stockham autosort, decimation in time, radices 4, 2, 5, 3 and a generic
 * O(r^2) butterfly for any other prime factor; twiddles W_n^k tabulated once per plan in
 * long double -> double.  Output is the plain unnormalised DFT
 *     X[k] = sum_j x[j] * exp(sign * 2*pi*i * j*k / n),
 * which is what every FFTW3 call site of the reference expects.
 */
#include "fftw3.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cpx;

struct standin_plan_s {
    int n, howmany, istride, idist, ostride, odist, sign;
    cpx *in, *out;
    int nfac, fac[32];
    cpx *tw;      /* tw[k] = exp(sign*2*pi*i*k/n), k < n */
    cpx *wa, *wb; /* work buffers, n each */
};

static void factorize(struct standin_plan_s *p) {
    int n = p->n; p->nfac = 0;
    while (n % 4 == 0) { p->fac[p->nfac++] = 4; n /= 4; }
    while (n % 2 == 0) { p->fac[p->nfac++] = 2; n /= 2; }
    for (int f = 3; f * f <= n; f += 2)
        while (n % f == 0) { p->fac[p->nfac++] = f; n /= f; }
    if (n > 1) p->fac[p->nfac++] = n;
}

/* Twiddle tables are cached per (n, sign) for the life of the process, so that re-planning the same
 * size (the reference plans a 640-point transform on every pilot_freq_sinh call, Frame.hpp:289-298)
 * costs what a wisdom hit costs in FFTW rather than n long-double sin/cos evaluations. */
static cpx *cached_twiddles(int n, int sign) {
    static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    static struct { int n, sign; cpx *tw; } cache[64];
    static int ncache = 0;
    pthread_mutex_lock(&mu);
    for (int i = 0; i < ncache; i++)
        if (cache[i].n == n && cache[i].sign == sign) { cpx *t = cache[i].tw; pthread_mutex_unlock(&mu); return t; }
    cpx *tw = (cpx *)malloc(sizeof(cpx) * (size_t)n);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int k = 0; k < n; k++) {
        long double a = two_pi * (long double)k / (long double)n;
        tw[k].re = (double)cosl(a);
        tw[k].im = (double)(sign < 0 ? -sinl(a) : sinl(a));
    }
    if (ncache < 64) { cache[ncache].n = n; cache[ncache].sign = sign; cache[ncache].tw = tw; ncache++; }
    pthread_mutex_unlock(&mu);
    return tw;
}

static struct standin_plan_s *make_plan(int n, int howmany, cpx *in, int istride, int idist,
                                        cpx *out, int ostride, int odist, int sign) {
    struct standin_plan_s *p = (struct standin_plan_s *)calloc(1, sizeof *p);
    p->n = n; p->howmany = howmany; p->in = in; p->out = out;
    p->istride = istride; p->idist = idist; p->ostride = ostride; p->odist = odist; p->sign = sign;
    if (n <= 0) return p;
    factorize(p);
    p->tw = cached_twiddles(n, sign);
    p->wa = (cpx *)malloc(sizeof(cpx) * (size_t)n);
    p->wb = (cpx *)malloc(sizeof(cpx) * (size_t)n);
    return p;
}

fftw_plan fftw_plan_many_dft(int rank, const int *n, int howmany,
                             fftw_complex *in, const int *inembed, int istride, int idist,
                             fftw_complex *out, const int *onembed, int ostride, int odist,
                             int sign, unsigned flags) {
    (void)inembed; (void)onembed; (void)flags;
    if (rank != 1) return NULL;
    return make_plan(n[0], howmany, (cpx *)in, istride, idist, (cpx *)out, ostride, odist, sign);
}

fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags) {
    (void)flags;
    return make_plan(n, 1, (cpx *)in, 1, n, (cpx *)out, 1, n, sign);
}

void fftw_destroy_plan(fftw_plan p) {
    if (!p) return;
    free(p->wa); free(p->wb); free(p);   /* p->tw belongs to the cache */
}

static inline cpx cmul(cpx a, cpx b) { cpx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }

/* one Stockham pass: radix r, ns = product of the radices already applied.
 * in/out are contiguous length-n arrays.  j indexes the n/r butterflies. */
static void pass(const struct standin_plan_s *p, const cpx *x, cpx *y, int r, int ns) {
    const int n = p->n, m = n / r, sg = p->sign;
    const int tstep = n / (ns * r);          /* W_{ns*r}^k = tw[k*tstep] */
    for (int j = 0; j < m; j++) {
        const int k = j % ns;
        const int o = (j / ns) * ns * r + k;
        if (r == 2) {
            cpx a = x[j], b = cmul(x[j + m], p->tw[k * tstep]);
            y[o].re = a.re + b.re; y[o].im = a.im + b.im;
            y[o + ns].re = a.re - b.re; y[o + ns].im = a.im - b.im;
        } else if (r == 4) {
            cpx a = x[j];
            cpx b = cmul(x[j + m], p->tw[k * tstep]);
            cpx c = cmul(x[j + 2 * m], p->tw[2 * k * tstep]);
            cpx d = cmul(x[j + 3 * m], p->tw[3 * k * tstep]);
            cpx s0 = { a.re + c.re, a.im + c.im }, s1 = { a.re - c.re, a.im - c.im };
            cpx s2 = { b.re + d.re, b.im + d.im }, s3 = { b.re - d.re, b.im - d.im };
            /* multiply s3 by sign*i : forward (-1) -> -i*s3 = (im, -re) */
            cpx t3; if (sg < 0) { t3.re = s3.im; t3.im = -s3.re; } else { t3.re = -s3.im; t3.im = s3.re; }
            y[o].re = s0.re + s2.re;          y[o].im = s0.im + s2.im;
            y[o + ns].re = s1.re + t3.re;     y[o + ns].im = s1.im + t3.im;
            y[o + 2 * ns].re = s0.re - s2.re; y[o + 2 * ns].im = s0.im - s2.im;
            y[o + 3 * ns].re = s1.re - t3.re; y[o + 3 * ns].im = s1.im - t3.im;
        } else {
            cpx v[64];
            cpx *vv = v, *heap = NULL;
            if (r > 64) vv = heap = (cpx *)malloc(sizeof(cpx) * (size_t)r);
            for (int q = 0; q < r; q++) vv[q] = cmul(x[j + q * m], p->tw[(long)q * k * tstep % n]);
            const int rstep = n / r;         /* W_r^k = tw[k*rstep] */
            for (int kk = 0; kk < r; kk++) {
                double sr = 0.0, si = 0.0;
                for (int q = 0; q < r; q++) {
                    cpx w = p->tw[(long)(q * kk % r) * rstep];
                    sr += vv[q].re * w.re - vv[q].im * w.im;
                    si += vv[q].re * w.im + vv[q].im * w.re;
                }
                y[o + kk * ns].re = sr; y[o + kk * ns].im = si;
            }
            free(heap);
        }
    }
}

void fftw_execute(const fftw_plan p) {
    if (!p || p->n <= 0) return;
    const int n = p->n;
    for (int h = 0; h < p->howmany; h++) {
        const cpx *src = p->in + (size_t)h * p->idist;
        cpx *dst = p->out + (size_t)h * p->odist;
        cpx *a = p->wa, *b = p->wb;
        for (int i = 0; i < n; i++) a[i] = src[(size_t)i * p->istride];
        int ns = 1;
        for (int f = 0; f < p->nfac; f++) {
            pass(p, a, b, p->fac[f], ns);
            ns *= p->fac[f];
            cpx *t = a; a = b; b = t;
        }
        for (int i = 0; i < n; i++) dst[(size_t)i * p->ostride] = a[i];
    }
}
