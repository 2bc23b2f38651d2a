/* cofdm_oracle.c -- plain-C, double-precision RESTATEMENT of the C-OFDM baseband hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the CPU oracle the CUDA path is checked against; it is
 * never linked into, loaded by, or reachable from the product library (c-ofdm_b200/).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against
 *   (1) the reference's own recorded artefacts (data/source.bin, data.bin -> t2_sin_corr.bin,
 *       phases.bin, constell.bin, data.txt; committed as tests/golden/ref_capture.npz), and
 *   (2) outputs of the unmodified reference sources compiled in the build container
 *       (oracle/_ref/libcofdm_ref.so, see oracle/ref_shim.cpp), committed as
 *       tests/golden/ref_vectors.npz by tests/golden/make_golden.py,
 * and, when oracle/_ref is present, directly against it on fresh random inputs.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Arithmetic mirrors the reference: std::complex<double> -> C99 double complex (same libgcc
 * __muldc3/__divdc3 semantics), std::abs -> cabs, std::arg -> carg, std::exp -> cexp.
 * The DFTs go through the FFTW3-subset API of oracle/standin (our own FFT; FFTW3 itself is a
 * system package that is absent from this image -- any correct double DFT agrees to ~1e-15).
 */
#define _GNU_SOURCE
#include "oracle_api.h"
#include "standin/fftw3.h"

#include <complex.h>
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef double complex cpx;

/* ---------------------------------------------------------------------------------------------
 * config/parser.cpp:4-33  parse_config: "key = long" lines, '#' comments, lines without '='
 * skipped, all whitespace inside key/value removed; a key that is absent reads as 0
 * (std::unordered_map::operator[] at every use site).
 * ------------------------------------------------------------------------------------------- */
#define MAX_KEYS 128
typedef struct { char key[64]; long val; } kv_t;
typedef struct { kv_t kv[MAX_KEYS]; int n; } config_t;

static long cfg_get(const config_t *c, const char *key) {
    for (int i = 0; i < c->n; i++)
        if (strcmp(c->kv[i].key, key) == 0) return c->kv[i].val;
    return 0;
}

static int cfg_parse(const char *path, config_t *c, char *err, size_t errlen) {
    FILE *f = fopen(path, "r");
    if (!f) { snprintf(err, errlen, "Cannot open config file"); return -1; }   /* parser.cpp:6 */
    char line[1024];
    c->n = 0;
    while (fgets(line, sizeof line, f)) {
        char *s = line;
        while (*s && isspace((unsigned char)*s)) s++;                          /* parser.cpp:13-14 */
        size_t len = strlen(s);
        while (len && isspace((unsigned char)s[len - 1])) s[--len] = 0;        /* parser.cpp:15-16 */
        if (!*s || *s == '#') continue;                                        /* parser.cpp:18 */
        char *eq = strchr(s, '=');
        if (!eq) continue;                                                     /* parser.cpp:20-21 */
        char key[64], val[64];
        size_t k = 0, v = 0;
        for (char *p = s; p < eq; p++) if (!isspace((unsigned char)*p) && k < 63) key[k++] = *p;
        for (char *p = eq + 1; *p; p++) if (!isspace((unsigned char)*p) && v < 63) val[v++] = *p;
        key[k] = 0; val[v] = 0;
        char *end;
        long x = strtol(val, &end, 10);                                        /* parser.cpp:30 std::stol */
        if (end == val) { snprintf(err, errlen, "stol: no conversion for key %s", key); fclose(f); return -1; }
        int i;
        for (i = 0; i < c->n; i++) if (strcmp(c->kv[i].key, key) == 0) break;
        if (i == c->n) { if (c->n == MAX_KEYS) break; strcpy(c->kv[c->n++].key, key); }
        c->kv[i].val = x;
    }
    fclose(f);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * std::mt19937 + std::uniform_int_distribution<int>(0,255) as used at OFDM/Frame.cpp:269-272.
 * libstdc++ (GCC >= 11, the reference objects were built by GCC 11.4.0) maps a 32-bit URNG onto a
 * range of 256 with Lemire's method: (uint64(rng()) * 256) >> 32 == rng() >> 24, and its rejection
 * threshold (2^32 mod 256) is 0, so no draw is ever rejected.  Pinned by data/source.bin.
 * ------------------------------------------------------------------------------------------- */
typedef struct { uint32_t mt[624]; int idx; } mt19937_t;
static void mt_seed(mt19937_t *m, uint32_t seed) {
    m->mt[0] = seed;
    for (int i = 1; i < 624; i++) m->mt[i] = 1812433253u * (m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) + (uint32_t)i;
    m->idx = 624;
}
static uint32_t mt_next(mt19937_t *m) {
    if (m->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (m->mt[i] & 0x80000000u) | (m->mt[(i + 1) % 624] & 0x7fffffffu);
            m->mt[i] = m->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        m->idx = 0;
    }
    uint32_t y = m->mt[m->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}

/* ---------------------------------------------------------------------------------------------
 * OFDM/modulation.cpp
 * ------------------------------------------------------------------------------------------- */
/* modulation.cpp:4-9  psk */
static cpx psk(uint8_t input, double angle, int deg) {
    double step = M_PI * 2 / (double)deg;
    return cexp(I * (step * (double)input + angle));
}
/* modulation.cpp:12-20  qam: low deg/2 bits -> I, high bits -> Q, natural binary, peak-normalised */
static cpx qam(uint8_t input, int deg) {
    if ((deg % 2) || deg > 8) return 0.0;
    uint8_t num = (uint8_t)(1u << (deg / 2));
    return CMPLX(2.0 / (num - 1) * (double)(input % num) - 1.0,
                 2.0 / (num - 1) * (double)(input >> (deg / 2)) - 1.0);
}
/* modulation.cpp:23-36  Modulation::Modulation (constellation table) */
static void constell_table(int mod, cpx *table) {
    int n = 1 << mod;
    for (int i = 0; i < n; i++)
        table[i] = (mod == 1) ? psk((uint8_t)i, M_PI_4 * 5, 2) : qam((uint8_t)i, mod);
}
/* modulation.cpp:90-125  Modulation::bit_stream_converter: regroup MSB-first, zero-pad the tail */
static size_t bsc_len(size_t out_bs, size_t in_bs, size_t len) {
    return (len * in_bs) / out_bs + ((len * in_bs) % out_bs > 0);
}
static size_t bit_stream_converter(size_t out_bs, size_t in_bs, const uint8_t *in, size_t len, uint8_t *out) {
    size_t out_len = bsc_len(out_bs, in_bs, len);
    memset(out, 0, out_len);
    size_t oi = 0, ii = 0;
    uint8_t mask = (uint8_t)(1 << (in_bs - 1));
    for (size_t i = 0, j = 0; i < in_bs * len; i++) {
        if (j == out_bs) { j = 0; oi++; }
        j++;
        out[oi] <<= 1;
        if ((mask & in[ii]) > 0) out[oi]++;
        mask >>= 1;
        if (mask == 0) { mask = (uint8_t)(1 << (in_bs - 1)); ii++; }
    }
    if ((len * in_bs) % out_bs > 0) out[out_len - 1] <<= out_bs - (len * in_bs) % out_bs;
    return out_len;
}
/* modulation.cpp:39-50  Modulation::mod */
static size_t modulate(int mod, const cpx *table, const uint8_t *bytes, size_t n_bytes, cpx *out) {
    size_t n = bsc_len((size_t)mod, 8, n_bytes);
    uint8_t *sym = (uint8_t *)malloc(n + 1);
    bit_stream_converter((size_t)mod, 8, bytes, n_bytes, sym);
    for (size_t i = 0; i < n; i++) out[i] = table[sym[i]];
    free(sym);
    return n;
}
/* modulation.cpp:53-87  Modulation::demod (QAM branch clamps the caller's points in place) */
static size_t demodulate(int mod, cpx *in, size_t len, uint8_t *bytes) {
    uint8_t *buf = (uint8_t *)malloc(len + 1);
    if (mod == 1) {
        for (size_t i = 0; i < len; i++) buf[i] = (uint8_t)(creal(in[i]) + cimag(in[i]) > 0);
    } else {
        uint8_t str_size = (uint8_t)(1u << (mod / 2));                 /* :23 */
        double step = 2.0 / (str_size - 1);                            /* :25 */
        double str_size_1 = 1.0 / step;                                /* :26 */
        for (size_t i = 0; i < len; i++) {                             /* :68-73 std::clamp */
            double re = creal(in[i]), im = cimag(in[i]);
            re = re < -1.0 ? -1.0 : (1.0 < re ? 1.0 : re);
            im = im < -1.0 ? -1.0 : (1.0 < im ? 1.0 : im);
            in[i] = CMPLX(re, im);
        }
        for (size_t i = 0; i < len; i++)                               /* :75-78 */
            buf[i] = (uint8_t)((uint8_t)((creal(in[i]) + 1.0) * str_size_1 + 0.5) |
                               (uint8_t)((cimag(in[i]) + 1.0) * str_size_1 + 0.5) * str_size);
    }
    size_t n = bit_stream_converter(8, (size_t)mod, buf, len, bytes);
    free(buf);
    return n;
}

/* ---------------------------------------------------------------------------------------------
 * OFDM/Frame.hpp:27-55, Frame.cpp:4-96   FFT_FORM
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int fft_size, num_data_subc, num_pilot_subc, num_symb, segment_step, segment_size;
    cpx *FFT_buf;          /* num_symb * fft_size */
    int *segment, *pilot;  /* offsets into FFT_buf (the reference stores pointers) */
    int n_pilot;           /* num_pilot_subc * num_symb */
    fftw_plan backward_plan, forward_plan;
    cpx *restored_buf;     /* num_data_subc * num_symb */
    double norm_factor, pilot_ampl;
} fft_form;

/* Frame.cpp:4-45 */
static void fft_form_init(fft_form *t, int fft_size, int nd, int np, int num_symb, double pilot_ampl) {
    t->fft_size = fft_size; t->num_data_subc = nd; t->num_pilot_subc = np; t->num_symb = num_symb;
    t->segment_step = nd / np + 1;                         /* :9  */
    t->segment_size = t->segment_step - 1;                 /* :10 */
    t->FFT_buf = (cpx *)calloc((size_t)num_symb * fft_size + 1, sizeof(cpx));
    t->n_pilot = np * num_symb;
    t->segment = (int *)calloc((size_t)t->n_pilot + 1, sizeof(int));
    t->pilot = (int *)calloc((size_t)t->n_pilot + 1, sizeof(int));
    t->backward_plan = fftw_plan_many_dft(1, &t->fft_size, num_symb, (fftw_complex *)t->FFT_buf, NULL, 1, fft_size,
                                          (fftw_complex *)t->FFT_buf, NULL, 1, fft_size, FFTW_BACKWARD, FFTW_MEASURE);
    t->forward_plan = fftw_plan_many_dft(1, &t->fft_size, num_symb, (fftw_complex *)t->FFT_buf, NULL, 1, fft_size,
                                         (fftw_complex *)t->FFT_buf, NULL, 1, fft_size, FFTW_FORWARD, FFTW_MEASURE);
    t->restored_buf = (cpx *)calloc((size_t)nd * num_symb + 1, sizeof(cpx));
    t->norm_factor = sqrt((double)fft_size);               /* :28 */
    t->pilot_ampl = pilot_ampl;
    int np2 = np / 2;                                      /* :31 */
    for (int i = 0, pi = 0, di = 0; i < num_symb; i++, pi += np, di += fft_size) {   /* :33-43 */
        int j = 0;
        for (int pos = 1 + t->segment_size; j < np2; j++, pos += t->segment_step) {
            t->pilot[pi + j] = di + pos;
            t->segment[pi + j] = di + pos - t->segment_size;
        }
        for (int pos = fft_size - t->segment_step * np2; j < np; j++, pos += t->segment_step) {
            t->pilot[pi + j] = di + pos;
            t->segment[pi + j] = di + pos + 1;
        }
    }
}
static void fft_form_free(fft_form *t) {
    fftw_destroy_plan(t->backward_plan); fftw_destroy_plan(t->forward_plan);
    free(t->FFT_buf); free(t->segment); free(t->pilot); free(t->restored_buf);
}
/* Frame.cpp:54-70  FFT_FORM::write */
static void fft_form_write(fft_form *t, const cpx *input) {
    size_t n = (size_t)t->num_symb * t->fft_size;
    for (size_t i = 0; i < n; i++) t->FFT_buf[i] = 0;                        /* :55 */
    for (int i = 0; i < t->n_pilot; i++) t->FFT_buf[t->pilot[i]] = t->pilot_ampl;   /* :56-57 */
    for (int i = 0; i < t->n_pilot; i++, input += t->segment_size)           /* :59-62 */
        memcpy(t->FFT_buf + t->segment[i], input, (size_t)t->segment_size * sizeof(cpx));
    fftw_execute(t->backward_plan);                                           /* :64 */
    for (size_t i = 0; i < n; i++) t->FFT_buf[i] /= t->norm_factor;           /* :66-68 */
}
/* Frame.cpp:73-96  FFT_FORM::read */
static cpx *fft_form_read(fft_form *t) {
    size_t n = (size_t)t->num_symb * t->fft_size;
    fftw_execute(t->forward_plan);                                            /* :74 */
    double phys_pilot_ampl = 0.0;
    for (int i = 0; i < t->n_pilot; i++) phys_pilot_ampl += cabs(t->FFT_buf[t->pilot[i]]);   /* :77-78 */
    phys_pilot_ampl /= t->n_pilot * t->pilot_ampl;                            /* :80 */
    for (size_t i = 0; i < n; i++) t->FFT_buf[i] /= phys_pilot_ampl;          /* :82-84 */
    cpx *out = t->restored_buf;
    for (int i = 0; i < t->n_pilot; i++, out += t->segment_size) {            /* :87-93 */
        memcpy(out, t->FFT_buf + t->segment[i], (size_t)t->segment_size * sizeof(cpx));
        cpx coef = t->FFT_buf[t->pilot[i]] / t->FFT_buf[t->pilot[i % t->num_pilot_subc]];
        for (int j = 0; j < t->segment_size; j++) out[j] /= coef;
    }
    return t->restored_buf;
}

/* ---------------------------------------------------------------------------------------------
 * OFDM/Frame.hpp:202-355, Frame.cpp:157-208   OFDM_FORM
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int fft_size, num_data_subc, num_pilot_subc, cp_size, num_symb, mod, ofdm_len, size, usefull_size;
    cpx *base;        /* output[0]; output[i] = base + i*ofdm_len  (Frame.cpp:178-182) */
    fft_form fft_task;
    cpx table[256];   /* Mod.constell */
} ofdm_form;

/* Frame.cpp:157-176 */
static void ofdm_form_init(ofdm_form *o, const config_t *c, int data, int with_preamble) {
    o->fft_size = (int)cfg_get(c, "fft_size");
    o->num_data_subc = (int)cfg_get(c, "num_data_subc");
    o->num_pilot_subc = (int)cfg_get(c, "num_pilot_subc");
    o->cp_size = (int)cfg_get(c, "cp_size");
    o->num_symb = (int)(with_preamble ? cfg_get(c, "num_symb") + cfg_get(c, "num_pr_symb")
                                      : (data ? cfg_get(c, "num_symb") : cfg_get(c, "num_pr_symb")));   /* :164 */
    o->mod = data ? (int)cfg_get(c, "modType") : 1;                                                       /* :167 */
    o->ofdm_len = o->fft_size + o->cp_size;
    o->size = o->ofdm_len * o->num_symb;
    o->usefull_size = o->num_data_subc * o->num_symb;
    o->base = NULL;
    fft_form_init(&o->fft_task, o->fft_size, o->num_data_subc, o->num_pilot_subc, o->num_symb,
                  (double)cfg_get(c, "pilot_ampl") / 1000);                                               /* :172 */
    constell_table(o->mod, o->table);
}
/* Frame.cpp:185-198  OFDM_FORM::write */
static void ofdm_form_write(ofdm_form *o, const uint8_t *bytes, size_t n_bytes) {
    size_t n_pts = bsc_len((size_t)o->mod, 8, n_bytes);
    size_t need = (size_t)o->fft_task.n_pilot * o->fft_task.segment_size;
    cpx *pts = (cpx *)calloc((n_pts > need ? n_pts : need) + 1, sizeof(cpx));
    modulate(o->mod, o->table, bytes, n_bytes, pts);                                     /* :187 */
    fft_form_write(&o->fft_task, pts);                                                   /* :189 */
    free(pts);
    for (int i = 0, j = 0; i < o->num_symb; i++, j += o->fft_size)                        /* :191-192 */
        memcpy(o->base + (size_t)i * o->ofdm_len + o->cp_size, o->fft_task.FFT_buf + j, (size_t)o->fft_size * sizeof(cpx));
    for (int i = 0; i < o->num_symb; i++)                                                 /* :196-197 */
        memcpy(o->base + (size_t)i * o->ofdm_len, o->base + (size_t)i * o->ofdm_len + o->fft_size,
               (size_t)o->cp_size * sizeof(cpx));
}
/* Frame.hpp:276-282  OFDM_FORM::fft (CP strip + FFT_FORM::read) */
static cpx *ofdm_form_fft(ofdm_form *o) {
    for (int i = 0, j = 0; i < o->num_symb; i++, j += o->fft_size)
        memcpy(o->fft_task.FFT_buf + j, o->base + (size_t)i * o->ofdm_len + o->cp_size, (size_t)o->fft_size * sizeof(cpx));
    return fft_form_read(&o->fft_task);
}
/* Frame.hpp:238-263  OFDM_FORM::cp_freq_sinh */
static void ofdm_form_cp_freq_sinh(ofdm_form *o) {
    cpx *x = o->base;
    cpx shift = 1.0;
    for (int i = 0; i < o->size; i += o->ofdm_len) {
        cpx phase = 0.0, curent_shift = 1.0;
        for (int j = 0; j < o->ofdm_len; ++j) x[i + j] *= shift;                           /* :247-249 */
        for (int j = 0; j < o->cp_size; j++) phase += conj(x[i + j]) * x[i + j + o->fft_size];   /* :251-253 */
        cpx step = cexp(-I * (carg(phase) / o->fft_size));                                 /* :254 */
        for (int j = 0; j < o->ofdm_len; ++j) { x[i + j] *= curent_shift; curent_shift *= step; }   /* :256-259 */
        shift *= curent_shift;                                                             /* :261 */
    }
}
/* Frame.hpp:265-274  OFDM_FORM::pr_phase_sinh */
static void ofdm_form_pr_phase_sinh(ofdm_form *o, const cpx *pr, int pr_size) {
    cpx phase = 0.0;
    for (int i = 0; i < pr_size; i++) phase += conj(pr[i]) * o->base[i];
    phase = cexp(-I * carg(phase));
    for (int i = 0; i < o->size; i++) o->base[i] *= phase;
}
/* Frame.hpp:285-337  OFDM_FORM::pilot_freq_sinh.  The out-of-bounds store at :322 writes a value
 * nobody reads and is not restated (the in-range clamp it was meant to be is a no-op whenever the
 * last border is <= size, which holds for every config the reference itself decodes). */
static double ofdm_form_pilot_freq_sinh(ofdm_form *o) {
    int size = o->size;
    cpx *spec = (cpx *)calloc((size_t)size + 1, sizeof(cpx));
    double *amplitude = (double *)calloc((size_t)size + 1, sizeof(double));
    fftw_plan plan = fftw_plan_dft_1d(size, (fftw_complex *)o->base, (fftw_complex *)spec, FFTW_FORWARD, FFTW_ESTIMATE);
    fftw_execute(plan);
    fftw_destroy_plan(plan);
    int half = size / 2;
    for (int i = 0; i < half; i++) {                                                       /* :300-309 */
        amplitude[i] = cabs(spec[i + half]);
        amplitude[i + half] = cabs(spec[i]);
    }
    double rel_bw = (double)(o->num_data_subc + o->num_pilot_subc) / (o->fft_size);        /* :311 */
    double rel_pilot_w = rel_bw / o->num_pilot_subc;                                       /* :312 */
    int pilot_w = (int)(size * rel_pilot_w);                                               /* :313 */
    int nb = o->num_pilot_subc + 2;
    int *borders = (int *)calloc((size_t)nb, sizeof(int));
    for (int i = 0, j = (int)((1.0 - rel_bw - rel_pilot_w) / 2.0 * size); i < nb; i++) {   /* :316-319 */
        borders[i] = j;
        j += pilot_w;
    }
    if (borders[0] < 0) borders[0] = 0;                                                    /* :321 */
    double shift = 0;
    for (int i = 0; i < o->num_pilot_subc + 1; i++) {                                      /* :325-331 */
        if (i == o->num_pilot_subc / 2) continue;
        int best = borders[i];
        for (int k = borders[i]; k < borders[i + 1]; k++)      /* std::max_element: first maximum */
            if (amplitude[k] > amplitude[best]) best = k;
        shift += best;
    }
    shift /= o->num_pilot_subc;                                                            /* :332 */
    shift -= size / 2;                                                                     /* :333 */
    shift /= size;                                                                         /* :334 */
    free(spec); free(amplitude); free(borders);
    return shift;
}
/* Frame.hpp:340-348  OFDM_FORM::freq_shift */
static void ofdm_form_freq_shift(ofdm_form *o, double shift) {
    cpx step = cexp(CMPLX(0, -2 * M_PI * shift));
    cpx phase = 1.0;
    for (int i = 0; i < o->size; i++) { o->base[i] *= phase; phase *= step; }
}

/* ---------------------------------------------------------------------------------------------
 * The handle = FRAME_FORM (Frame.hpp:442-519, Frame.cpp:213-256) + T2SIN_FORM + PREAMBLE_FORM
 * ------------------------------------------------------------------------------------------- */
/* one FRAME_FORM */
typedef struct frame_form {
    config_t config;
    /* T2SIN_FORM (Frame.hpp:58-199, Frame.cpp:99-154) */
    int t2_size, t2_f1, t2_f2, t2_smooth;
    double t2_level;
    double *detect_mask;
    cpx *detect_buf;
    fftw_plan detect_plan;
    /* PREAMBLE_FORM (Frame.hpp:358-439, Frame.cpp:259-378) */
    ofdm_form preamble;
    double pr_level;
    int pr_sin_len;
    uint8_t *preamble_bytes; int n_preamble_bytes;
    cpx *mod_preamble, *ofdm_preamble, *conjected_sinh_part, *chan_est;
    int cor_size;
    /* message / message_with_preamble */
    ofdm_form message, message_with_preamble;
    /* FRAME_FORM buffers */
    cpx *buf; int16_t *int16_buf;
    cpx *from_sdr_buf; int16_t *from_sdr_int16_buf; long from_sdr_size;
    int usefull_size, output_size;
} frame_form;

/* like reference main.cpp:23-24 the handle owns a tx FRAME_FORM and an rx FRAME_FORM */
struct oc_handle { frame_form *tx_frame, *rx_frame; };

static __thread char g_err[256];

const char *oc_kind(void) { return "port"; }
const char *oc_last_error(void) { return g_err; }

static frame_form *frame_form_create(const char *config_path) {
    frame_form *h = (frame_form *)calloc(1, sizeof *h);
    if (cfg_parse(config_path, &h->config, g_err, sizeof g_err)) { free(h); return NULL; }
    const config_t *c = &h->config;
    /* T2SIN_FORM::T2SIN_FORM  Frame.cpp:99-136 */
    h->t2_size = (int)cfg_get(c, "T2sin_size");
    h->t2_f1 = (int)cfg_get(c, "T2_sin_f1");
    h->t2_f2 = (int)cfg_get(c, "T2_sin_f2");
    h->t2_smooth = (int)cfg_get(c, "smooth");
    h->t2_level = (double)cfg_get(c, "T2_sin_level") / 1000;                   /* :105 */
    h->detect_mask = (double *)calloc((size_t)h->t2_size + 1, sizeof(double));
    h->detect_buf = (cpx *)calloc((size_t)h->t2_size + 1, sizeof(cpx));
    h->detect_plan = fftw_plan_many_dft(1, &h->t2_size, 1, (fftw_complex *)h->detect_buf, NULL, 1, h->t2_size,
                                        (fftw_complex *)h->detect_buf, NULL, 1, h->t2_size, FFTW_FORWARD, FFTW_MEASURE);
    {
        int a1 = h->t2_f1 - h->t2_smooth; if (a1 < 0) a1 = 0;                  /* :120-123 */
        int b1 = h->t2_f1 + h->t2_smooth; if (b1 > h->t2_size - 1) b1 = h->t2_size - 1;
        int a2 = h->t2_f2 - h->t2_smooth; if (a2 < 0) a2 = 0;
        int b2 = h->t2_f2 + h->t2_smooth; if (b2 > h->t2_size - 1) b2 = h->t2_size - 1;
        for (int i = a1; i <= b1; i++) h->detect_mask[i] += 1.0;               /* :127-133 */
        for (int i = a2; i <= b2; i++) h->detect_mask[i] += 1.0;
    }
    /* PREAMBLE_FORM::PREAMBLE_FORM  Frame.cpp:259-273 */
    ofdm_form_init(&h->preamble, c, 0, 0);
    h->pr_level = (double)cfg_get(c, "pr_level") / 1000;
    h->pr_sin_len = (int)cfg_get(c, "pr_sin_len");
    h->n_preamble_bytes = h->preamble.usefull_size * h->preamble.mod / 8;
    h->preamble_bytes = (uint8_t *)calloc((size_t)h->n_preamble_bytes + 1, 1);
    h->cor_size = (int)cfg_get(c, "T2sin_size") * 2 + h->pr_sin_len;           /* :266 */
    h->chan_est = (cpx *)calloc((size_t)h->preamble.num_data_subc + 1, sizeof(cpx));
    {
        mt19937_t rng;
        mt_seed(&rng, (uint32_t)cfg_get(c, "pr_seed"));                        /* :269-272 */
        for (int i = 0; i < h->n_preamble_bytes; i++) h->preamble_bytes[i] = (uint8_t)(mt_next(&rng) >> 24);
    }
    ofdm_form_init(&h->message, c, 1, 0);                                      /* Frame.cpp:217 */
    ofdm_form_init(&h->message_with_preamble, c, 1, 1);                        /* Frame.cpp:218 */
    /* FRAME_FORM::FRAME_FORM  Frame.cpp:219-231 */
    h->output_size = h->t2_size + h->preamble.size + h->message.size;
    h->buf = (cpx *)calloc((size_t)h->output_size + 1, sizeof(cpx));
    h->int16_buf = (int16_t *)calloc((size_t)h->output_size * 2 + 2, sizeof(int16_t));
    h->from_sdr_size = (long)h->output_size * (cfg_get(c, "rx_buf_size") + 1);
    h->from_sdr_buf = (cpx *)calloc((size_t)h->from_sdr_size + 1, sizeof(cpx));
    h->from_sdr_int16_buf = (int16_t *)calloc((size_t)h->from_sdr_size * 2 + 2, sizeof(int16_t));
    h->usefull_size = h->message.usefull_size * h->message.mod / 8;            /* :223 */
    /* T2SIN_FORM::set  Frame.cpp:139-154: delta at f1 and f2 (0.5 each), unnormalised backward FFT in place */
    if (h->t2_size) {
        h->buf[h->t2_f1] = 0.5;
        h->buf[h->t2_f2] = 0.5;
        fftw_plan bp = fftw_plan_many_dft(1, &h->t2_size, 1, (fftw_complex *)h->buf, NULL, 1, h->t2_size,
                                          (fftw_complex *)h->buf, NULL, 1, h->t2_size, FFTW_BACKWARD, FFTW_ESTIMATE);
        fftw_execute(bp);
        fftw_destroy_plan(bp);
    }
    /* PREAMBLE_FORM::set  Frame.cpp:276-294 */
    h->preamble.base = h->buf + h->t2_size;
    ofdm_form_write(&h->preamble, h->preamble_bytes, (size_t)h->n_preamble_bytes);          /* :281 */
    h->ofdm_preamble = (cpx *)calloc((size_t)h->preamble.size + 1, sizeof(cpx));
    memcpy(h->ofdm_preamble, h->preamble.base, (size_t)h->preamble.size * sizeof(cpx));     /* :282 */
    h->mod_preamble = (cpx *)calloc((size_t)h->preamble.usefull_size + 1, sizeof(cpx));
    modulate(1, h->preamble.table, h->preamble_bytes, (size_t)h->n_preamble_bytes, h->mod_preamble);   /* :283 */
    h->conjected_sinh_part = (cpx *)calloc((size_t)h->pr_sin_len + 1, sizeof(cpx));
    {
        double norm = 0.0;
        for (int i = 0; i < h->pr_sin_len; i++) {                                           /* :286-289 */
            h->conjected_sinh_part[i] = conj(h->ofdm_preamble[i]);
            norm += cabs(h->conjected_sinh_part[i] * h->conjected_sinh_part[i]);
        }
        norm = sqrt(norm);
        for (int i = 0; i < h->pr_sin_len; i++) h->conjected_sinh_part[i] /= (cpx)norm;     /* :291-293 */
    }
    h->message.base = h->buf + h->t2_size + h->preamble.size;                                /* Frame.cpp:230 */
    h->message_with_preamble.base = h->buf + h->t2_size;                                     /* Frame.cpp:231 */
    return h;
}

static void frame_form_destroy(frame_form *h) {
    if (!h) return;
    fftw_destroy_plan(h->detect_plan);
    fft_form_free(&h->preamble.fft_task); fft_form_free(&h->message.fft_task); fft_form_free(&h->message_with_preamble.fft_task);
    free(h->detect_mask); free(h->detect_buf); free(h->preamble_bytes); free(h->chan_est);
    free(h->mod_preamble); free(h->ofdm_preamble); free(h->conjected_sinh_part);
    free(h->buf); free(h->int16_buf); free(h->from_sdr_buf); free(h->from_sdr_int16_buf);
    free(h);
}

oc_handle *oc_create(const char *config_path) {
    frame_form *tx = frame_form_create(config_path);
    if (!tx) return NULL;
    frame_form *rx = frame_form_create(config_path);
    if (!rx) { frame_form_destroy(tx); return NULL; }
    oc_handle *h = (oc_handle *)calloc(1, sizeof *h);
    h->tx_frame = tx; h->rx_frame = rx;
    return h;
}
void oc_destroy(oc_handle *h) {
    if (!h) return;
    frame_form_destroy(h->tx_frame); frame_form_destroy(h->rx_frame); free(h);
}

void oc_get_sizes(const oc_handle *hh, oc_sizes *o) {
    const frame_form *h = hh->rx_frame;
    const config_t *c = &h->config;
    memset(o, 0, sizeof *o);
    o->fft_size = h->message.fft_size; o->num_data_subc = h->message.num_data_subc;
    o->num_pilot_subc = h->message.num_pilot_subc; o->cp_size = h->message.cp_size;
    o->num_symb = h->message.num_symb; o->num_pr_symb = h->preamble.num_symb;
    o->pr_sin_len = h->pr_sin_len; o->pr_seed = (int)cfg_get(c, "pr_seed");
    o->t2sin_size = h->t2_size; o->t2_f1 = h->t2_f1; o->t2_f2 = h->t2_f2; o->smooth = h->t2_smooth;
    o->mod_type = h->message.mod; o->ofdm_len = h->message.ofdm_len;
    o->preamble_size = h->preamble.size; o->message_size = h->message.size;
    o->output_size = h->output_size; o->usefull_size = h->usefull_size;
    o->constell_size = h->message.usefull_size;
    o->mult = (int)cfg_get(c, "mult"); o->rx_buf_size = (int)cfg_get(c, "rx_buf_size");
    o->iterations = (int)cfg_get(c, "iterations"); o->cor_size = h->cor_size;
    o->t2_level = h->t2_level; o->pr_level = h->pr_level; o->pilot_ampl = h->message.fft_task.pilot_ampl;
}

static void put(double *dst, const cpx *src, size_t n) { if (dst) memcpy(dst, src, n * sizeof(cpx)); }

void oc_get_constants(oc_handle *hh, double *t2sin_tone, double *t2_mask, uint8_t *preamble_bytes,
                      double *ofdm_preamble, double *mod_preamble, double *matched, double *constell) {
    frame_form *h = hh->tx_frame;
    /* the tone and the preamble are frame invariant: rebuild them as the constructor did */
    if (t2sin_tone) {
        cpx *tmp = (cpx *)calloc((size_t)h->t2_size + 1, sizeof(cpx));
        tmp[h->t2_f1] = 0.5; tmp[h->t2_f2] = 0.5;
        fftw_plan bp = fftw_plan_dft_1d(h->t2_size, (fftw_complex *)tmp, (fftw_complex *)tmp, FFTW_BACKWARD, FFTW_ESTIMATE);
        fftw_execute(bp); fftw_destroy_plan(bp);
        put(t2sin_tone, tmp, (size_t)h->t2_size);
        free(tmp);
    }
    if (t2_mask) memcpy(t2_mask, h->detect_mask, (size_t)h->t2_size * sizeof(double));
    if (preamble_bytes) memcpy(preamble_bytes, h->preamble_bytes, (size_t)h->n_preamble_bytes);
    put(ofdm_preamble, h->ofdm_preamble, (size_t)h->preamble.size);
    put(mod_preamble, h->mod_preamble, (size_t)h->preamble.usefull_size);
    put(matched, h->conjected_sinh_part, (size_t)h->pr_sin_len);
    put(constell, h->message.table, (size_t)1 << h->message.mod);
}

int oc_bit_stream_converter(int out_bits, int in_bits, const uint8_t *in, int n_in, uint8_t *out) {
    return (int)bit_stream_converter((size_t)out_bits, (size_t)in_bits, in, (size_t)n_in, out);
}
int oc_mod(int mod_type, const uint8_t *bytes, int n_bytes, double *points) {
    cpx table[256];
    constell_table(mod_type, table);
    return (int)modulate(mod_type, table, bytes, (size_t)n_bytes, (cpx *)points);
}
int oc_demod(int mod_type, double *points_inout, int n_points, uint8_t *bytes) {
    return (int)demodulate(mod_type, (cpx *)points_inout, (size_t)n_points, bytes);
}

/* FRAME_FORM::write / get / get_int16  Frame.cpp:235-256 */
void oc_tx(oc_handle *hh, const uint8_t *bytes, double *frame, int16_t *frame_i16) {
    frame_form *h = hh->tx_frame;
    ofdm_form_write(&h->message, bytes, (size_t)h->usefull_size);               /* :236 */
    put(frame, h->buf, (size_t)h->output_size);                                 /* :244-246 */
    if (frame_i16) {
        double mult = (double)cfg_get(&h->config, "mult");
        for (int i = 0; i < h->output_size; i++) {                              /* :251-253 trunc toward zero */
            cpx v = h->buf[i] * (cpx)mult;
            h->int16_buf[2 * i] = (int16_t)creal(v);
            h->int16_buf[2 * i + 1] = (int16_t)cimag(v);
        }
        memcpy(frame_i16, h->int16_buf, (size_t)h->output_size * 2 * sizeof(int16_t));
    }
}

/* one block of T2SIN_FORM::corr / find_t2sin  Frame.hpp:112-143 / 164-193.
 * returns 1 and *rel when the block has a usable ratio, 0 when it is skipped (`continue`). */
static int t2sin_block(frame_form *h, const cpx *sig, double *rel) {
    double total_energy = 0.0, sin_energy = 0.0;
    memcpy(h->detect_buf, sig, (size_t)h->t2_size * sizeof(cpx));
    fftw_execute(h->detect_plan);
    for (int j = 0; j < h->t2_size; j++) {
        double re = creal(h->detect_buf[j]), im = cimag(h->detect_buf[j]);
        double subc_energy = re * re + im * im;
        total_energy += subc_energy;
        sin_energy += h->detect_mask[j] * subc_energy;
    }
    if (total_energy == 0) return 0;
    *rel = sin_energy / total_energy;
    if (isnan(*rel)) return 0;
    return 1;
}
/* T2SIN_FORM::corr  Frame.hpp:96-147 */
int oc_t2sin_corr(oc_handle *hh, const double *sig, long n, double *out) {
    frame_form *h = hh->rx_frame;
    int cycles = (int)(n / h->t2_size);
    const cpx *p = (const cpx *)sig;
    for (int i = 0; i < cycles; i++, p += h->t2_size) {
        double rel;
        out[i] = 0.0;
        if (t2sin_block(h, p, &rel) && rel > h->t2_level) out[i] = rel;
    }
    return cycles;
}
/* T2SIN_FORM::find_t2sin  Frame.hpp:150-197 */
static int find_t2sin(frame_form *h, const cpx *sig, long n, int start) {
    int cycles = (int)((n - start) / h->t2_size);
    const cpx *p = sig + start;
    for (int i = 0; i < cycles; i++, p += h->t2_size) {
        double rel;
        if (t2sin_block(h, p, &rel) && rel > h->t2_level) return i * h->t2_size + start;
    }
    return -1;
}
int oc_find_t2sin(oc_handle *hh, const double *sig, long n, int start) { return find_t2sin(hh->rx_frame, (const cpx *)sig, n, start); }

/* PREAMBLE_FORM::find_corr (store == 1, Frame.cpp:297-335) and find_preamble (store == 0,
 * Frame.cpp:338-378) share one loop in the reference apart from store-vs-return. */
static int preamble_scan(frame_form *h, const cpx *input, int start, double *cor) {
    double norm = 0, re, im;
    const cpx *p = input + start;
    for (int i = 0; i < h->pr_sin_len; i++, p++) { re = creal(*p); im = cimag(*p); norm += re * re + im * im; }
    p = input + start;
    if (cor) for (int i = 0; i < h->cor_size; i++) cor[i] = 0.0;
    for (int i = 0; i < h->cor_size; i++, p++) {
        cpx energy = 0.0;
        if (norm > 1.0) {
            for (int j = 0; j < h->pr_sin_len; j++) energy += p[j] * h->conjected_sinh_part[j];
            if (cor) { cor[i] = cabs(energy); cor[i] /= sqrt(norm); }
            else if (cabs(energy) / sqrt(norm) > h->pr_level) return i + start;
        }
        re = creal(p[h->pr_sin_len]); im = cimag(p[h->pr_sin_len]); norm += re * re + im * im;
        re = creal(p[0]); im = cimag(p[0]); norm -= re * re + im * im;
    }
    return -10;
}
void oc_find_corr(oc_handle *hh, const double *sig, long n, int start, double *cor) {
    frame_form *h = hh->rx_frame;
    (void)n; preamble_scan(h, (const cpx *)sig, start, cor);
}
int oc_find_preamble(oc_handle *hh, const double *sig, long n, int start) {
    frame_form *h = hh->rx_frame;
    (void)n; return preamble_scan(h, (const cpx *)sig, start, NULL);
}

/* PREAMBLE_FORM::chan_char_lq  Frame.hpp:389-434 (sums are used where means belong -- kept) */
static cpx *chan_char_lq(frame_form *h) {
    cpx *pr = ofdm_form_fft(&h->preamble);
    int nce = h->preamble.num_data_subc, nph = nce / 2;
    double mean_x = 0.0, mean_y = 0.0, mean_xy = 0.0, mean_x2 = 0.0, a, b;
    for (int i = 0; i < nce; i++) h->chan_est[i] = 0.0;
    double *phase = (double *)calloc((size_t)nph + 1, sizeof(double));
    for (int i = 0; i < nph; i++) phase[i] = carg(pr[i] / h->mod_preamble[i]);          /* :403-405 */
    for (int i = 1; i < nph; i++) {                                                      /* :407-414 */
        double d = phase[i] - phase[i - 1];
        if (d > M_PI) phase[i] -= 2 * M_PI;
        else if (d < -M_PI) phase[i] += 2 * M_PI;
    }
    for (int i = 0; i < nph; i++) {                                                      /* :416-421 */
        mean_xy += phase[i] * i;
        mean_x2 += i * i;
        mean_x += i;
        mean_y += phase[i];
    }
    b = (mean_xy - mean_x * mean_y) / (mean_x2 - mean_x * mean_x);                       /* :422 */
    a = mean_y - b * mean_x;                                                             /* :423 */
    for (int i = 0; i < nce / 2; i++) h->chan_est[i] = cexp(CMPLX(0, b * i + a));        /* :425-427 */
    for (int i = nce / 2; i < nce; i++)                                                  /* :428-430 */
        h->chan_est[i] = cexp(CMPLX(0, -b * (double)nce / 2 + (double)(i - nce / 2) * b + a));
    free(phase);
    return h->chan_est;
}

/* main.cpp:60-80 == rx.cpp:200-220 on the samples sitting in buf[t2_size ...] */
static void demod_chain(frame_form *h, double *scal, double *synced, double *grid, double *chan,
                        double *constell_out, uint8_t *bytes) {
    double fs = ofdm_form_pilot_freq_sinh(&h->preamble);                                 /* main.cpp:60 */
    ofdm_form_freq_shift(&h->message_with_preamble, fs);                                 /* :61 */
    ofdm_form_cp_freq_sinh(&h->message_with_preamble);                                   /* :62 */
    ofdm_form_pr_phase_sinh(&h->message_with_preamble, h->ofdm_preamble, h->preamble.size);   /* :63 */
    put(synced, h->buf + h->t2_size, (size_t)h->message_with_preamble.size);
    cpx *chan_char = chan_char_lq(h);                                                    /* :66 */
    cpx *restored = ofdm_form_fft(&h->message);                                          /* :67 */
    put(grid, h->message.fft_task.FFT_buf, (size_t)h->message.num_symb * h->message.fft_size);
    int n = h->message.usefull_size, nce = h->preamble.num_data_subc;
    cpx *constell = (cpx *)malloc(((size_t)n + 1) * sizeof(cpx));
    for (int i = 0; i < n; i++) constell[i] = restored[i] / chan_char[i % nce];          /* :69-71 */
    put(chan, chan_char, (size_t)nce);
    put(constell_out, constell, (size_t)n);
    if (scal) {
        scal[0] = fs;
        scal[1] = carg(chan_char[0]);
        scal[2] = carg(chan_char[1] / chan_char[0]);
        scal[3] = 0.0;
    }
    if (bytes) demodulate(h->message.mod, constell, (size_t)n, bytes);                   /* :80 */
    free(constell);
}

void oc_rx_aligned(oc_handle *hh, const double *rx_samples, double *scal, double *synced, double *grid,
                   double *chan, double *constell, uint8_t *bytes) {
    frame_form *h = hh->rx_frame;
    memcpy(h->buf + h->t2_size, rx_samples, (size_t)h->message_with_preamble.size * sizeof(cpx));   /* main.cpp:55-58 */
    demod_chain(h, scal, synced, grid, chan, constell, bytes);
}

/* FRAME_FORM::read  Frame.cpp:239-242 -> OFDM_FORM::read Frame.cpp:201-208 */
void oc_read(oc_handle *hh, const double *frame, double *restored, uint8_t *bytes) {
    frame_form *h = hh->rx_frame;
    memcpy(h->buf, frame, sizeof(cpx) * (size_t)h->output_size);
    cpx *r = ofdm_form_fft(&h->message);
    put(restored, r, (size_t)h->message.usefull_size);
    demodulate(h->message.mod, r, (size_t)h->message.usefull_size, bytes);
}

/* PREAMBLE_FORM::chan_char  Frame.hpp:375-385 */
void oc_chan_char(oc_handle *hh, const double *rx_samples, double *chan) {
    frame_form *h = hh->rx_frame;
    memcpy(h->buf + h->t2_size, rx_samples, (size_t)h->preamble.size * sizeof(cpx));
    cpx *pr = ofdm_form_fft(&h->preamble);
    int nd = h->preamble.num_data_subc, ns = h->preamble.num_symb;
    for (int i = 0; i < nd; i++) h->chan_est[i] = 0.0;
    for (int i = 0; i < nd * ns; i++) h->chan_est[i % nd] += pr[i] / h->mod_preamble[i];
    for (int i = 0; i < nd; i++) h->chan_est[i] /= (cpx)(double)ns;
    put(chan, h->chan_est, (size_t)nd);
}

/* FRAME_FORM::form_int16_to_double  Frame.hpp:472-481 */
static void form_int16_to_double(frame_form *h) {
    long len = h->from_sdr_size * 2;
    double *d = (double *)h->from_sdr_buf;
    for (long i = 0; i < len; ++i) d[i] = (double)h->from_sdr_int16_buf[i];
}

/* rx.cpp:101-235: acquisition state machine over an in-memory capture (see oracle_api.h) */
int oc_rx_stream(oc_handle *hh, const int16_t *capture, long n_samples, int max_frames,
                 long *pr_begin_abs, uint8_t *bytes) {
    frame_form *h = hh->rx_frame;
    const long block = (long)h->output_size * cfg_get(&h->config, "rx_buf_size");   /* sdr.hpp:141 */
    const long n_blocks = n_samples / block;
    long next_block = 0, cur_block = -1;
    memset(h->from_sdr_int16_buf, 0, (size_t)h->from_sdr_size * 2 * sizeof(int16_t));
#define BUF_UPDATE()  /* rx.cpp:73-91 */ \
    (next_block >= n_blocks ? 0 : (memcpy(h->from_sdr_int16_buf + 2 * (size_t)h->output_size, \
        capture + 2 * next_block * block, (size_t)block * 2 * sizeof(int16_t)), \
        cur_block = next_block++, form_int16_to_double(h), 1))
#define CARRY()       /* rx.cpp:149-153 / 182-186 */ \
    memcpy(h->from_sdr_int16_buf, h->from_sdr_int16_buf + 2 * (size_t)threshold, (size_t)h->output_size * 2 * sizeof(int16_t))
    if (!BUF_UPDATE()) return 0;                                          /* rx.cpp:103-112 */
    int pos = 0;
    const int threshold = (int)(h->from_sdr_size - h->output_size);      /* rx.cpp:116 */
    const int cycles = (int)cfg_get(&h->config, "iterations");           /* rx.cpp:124 */
    int found = 0;
    for (int i = 0; i < cycles && found < max_frames; i++) {              /* rx.cpp:126 */
        pos = find_t2sin(h, h->from_sdr_buf, h->from_sdr_size, pos);      /* :133 */
        if (pos == -1) {                                                  /* :137-145 */
            pos = h->output_size;
            if (!BUF_UPDATE()) break;
            continue;
        }
        if (pos >= threshold) {                                           /* :147-156 */
            pos -= threshold;
            CARRY();
            if (!BUF_UPDATE()) break;
        }
        int preamble_begin = preamble_scan(h, h->from_sdr_buf, pos, NULL) + 1;   /* :158 */
        if (preamble_begin < -2) { pos += h->message.size; continue; }    /* :160-166 */
        pos = preamble_begin;                                             /* :168 */
        if (pos == -1) {                                                  /* :170-178 */
            pos = h->output_size;
            if (!BUF_UPDATE()) break;
            continue;
        }
        if (pos >= threshold + h->t2_size) {                              /* :180-189 */
            pos -= threshold;
            CARRY();
            if (!BUF_UPDATE()) break;
        }
        memcpy(h->buf + h->t2_size, h->from_sdr_buf + pos,
               (size_t)(h->output_size - h->t2_size) * sizeof(cpx));      /* :192-196 */
        if (pr_begin_abs) pr_begin_abs[found] = cur_block * block + (long)pos - h->output_size;
        pos += h->message.size;                                           /* :198 */
        demod_chain(h, NULL, NULL, NULL, NULL, NULL, bytes ? bytes + (size_t)found * h->usefull_size : NULL);
        found++;
    }
#undef BUF_UPDATE
#undef CARRY
    return found;
}

/* bench.py CPU baseline: tx -> int16 -> double -> aligned rx, n_frames times (see oracle_api.h) */
long oc_txrx_loop(oc_handle *hh, const uint8_t *payloads, int n_frames, uint8_t *bytes_out) {
    frame_form *t = hh->tx_frame, *r = hh->rx_frame;
    const int us = t->usefull_size, n = t->output_size;
    double mult = (double)cfg_get(&t->config, "mult");
    long bad = 0;
    uint8_t *tmp = (uint8_t *)malloc((size_t)us + 1);
    for (int f = 0; f < n_frames; f++) {
        const uint8_t *pay = payloads + (size_t)f * us;
        ofdm_form_write(&t->message, pay, (size_t)us);                    /* Frame.cpp:235-237 */
        for (int i = 0; i < n; i++) {                                      /* Frame.cpp:249-256 */
            cpx v = t->buf[i] * (cpx)mult;
            t->int16_buf[2 * i] = (int16_t)creal(v);
            t->int16_buf[2 * i + 1] = (int16_t)cimag(v);
        }
        double *d = (double *)r->buf;                                      /* Frame.hpp:472-481 */
        for (int i = 0; i < 2 * n; i++) d[i] = (double)t->int16_buf[i];
        uint8_t *out = bytes_out ? bytes_out + (size_t)f * us : tmp;
        demod_chain(r, NULL, NULL, NULL, NULL, NULL, out);                 /* main.cpp:60-80 */
        for (int i = 0; i < us; i++) bad += out[i] != pay[i];
    }
    free(tmp);
    return bad;
}
