// ref_shim.cpp -- C ABI (oracle_api.h) over the UNMODIFIED reference classes.  TEST INFRASTRUCTURE ONLY.
//
// Built by oracle/Makefile into oracle/_ref/libcofdm_ref.so from the reference sources where they
// lie (/root/reference/OFDM/Frame.cpp, OFDM/modulation.cpp, config/parser.cpp) plus this file and
// the stand-in FFT (oracle/standin).  No reference source is copied or edited.
//
// Two accommodations, both outside the reference sources:
//  * `-include cstdint` on the command line (GCC 13 no longer leaks uint8_t into modulation.hpp:23).
//  * the reference writes one int out of bounds on every pilot_freq_sinh() call
//    (OFDM/Frame.hpp:322: `borders[num_data_subc+1]` on a num_pilot_subc+2 element vector).  The
//    value written is never read, but the store lands (num_data_subc+1)*4 bytes past a 40-byte heap
//    block.  Instead of patching the source, this library replaces operator new so that every
//    allocation made by the reference code carries that much slack; the stray store then falls
//    into the block's own padding and the algorithm runs exactly as written.
//
// The call order of oc_rx_aligned is reference main.cpp:60-80 (== rx.cpp:200-220); oc_rx_stream
// replays rx.cpp:101-235 with an in-memory capture in place of SDR::recv (the apps themselves
// cannot be compiled: mac/mac_frame.hpp is missing from the tree, iio.h / Python.h are absent).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "OFDM/Frame.hpp"   // from /root/reference via -I
#include "oracle_api.h"

// ---- padded allocator (see header comment) -------------------------------------------------------
static size_t g_pad = 4096;
void *operator new(std::size_t n) {
    void *p = std::malloc(n + g_pad);
    if (!p) throw std::bad_alloc();
    return p;
}
void *operator new[](std::size_t n) { return operator new(n); }
void operator delete(void *p) noexcept { std::free(p); }
void operator delete[](void *p) noexcept { std::free(p); }
void operator delete(void *p, std::size_t) noexcept { std::free(p); }
void operator delete[](void *p, std::size_t) noexcept { std::free(p); }

struct oc_handle {
    FRAME_FORM tx_frame;
    FRAME_FORM rx_frame;
    explicit oc_handle(const std::string &path) : tx_frame(path), rx_frame(path) {}
};

static thread_local std::string g_err;

extern "C" {

const char *oc_kind(void) { return "reference"; }
const char *oc_last_error(void) { return g_err.c_str(); }

oc_handle *oc_create(const char *config_path) {
    try {
        ConfigMap cfg = parse_config(config_path);
        size_t need = (size_t)(cfg["num_data_subc"] + 4) * sizeof(int);
        if (need > g_pad) g_pad = (need + 4095) & ~(size_t)4095;
        return new oc_handle(config_path);
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

void oc_destroy(oc_handle *h) { delete h; }

void oc_get_sizes(const oc_handle *hc, oc_sizes *o) {
    oc_handle *h = const_cast<oc_handle *>(hc);
    FRAME_FORM &f = h->rx_frame;
    std::memset(o, 0, sizeof *o);
    o->fft_size = f.message.fft_size;
    o->num_data_subc = f.message.num_data_subc;
    o->num_pilot_subc = f.message.num_pilot_subc;
    o->cp_size = f.message.cp_size;
    o->num_symb = f.message.num_symb;
    o->num_pr_symb = f.preamble.num_symb;
    o->pr_sin_len = f.preamble.pr_sin_len;
    o->pr_seed = f.preamble.pr_seed;
    o->t2sin_size = f.t2sin.size;
    o->t2_f1 = f.t2sin.f1;
    o->t2_f2 = f.t2sin.f2;
    o->smooth = f.t2sin.smooth;
    o->mod_type = (int)f.message.modType;
    o->ofdm_len = f.message.ofdm_len;
    o->preamble_size = f.preamble.size;
    o->message_size = f.message.size;
    o->output_size = f.output_size;
    o->usefull_size = f.usefull_size;
    o->constell_size = f.message.usefull_size;
    o->mult = (int)f.config["mult"];
    o->rx_buf_size = (int)f.config["rx_buf_size"];
    o->iterations = (int)f.config["iterations"];
    o->cor_size = (int)f.preamble.cor.size();
    o->t2_level = f.t2sin.level;
    o->pr_level = f.preamble.level;
    o->pilot_ampl = f.message.fft_task.pilot_ampl;
}

static void put(double *dst, const complex_double *src, size_t n) {
    if (dst) std::memcpy(dst, src, n * sizeof(complex_double));
}

void oc_get_constants(oc_handle *h, double *t2sin_tone, double *t2_mask, uint8_t *preamble_bytes,
                      double *ofdm_preamble, double *mod_preamble, double *matched, double *constell) {
    FRAME_FORM &f = h->tx_frame;   // tx_frame.buf still holds the constructor-built tone + preamble
    put(t2sin_tone, f.buf.data(), f.t2sin.size);
    if (t2_mask) std::memcpy(t2_mask, f.t2sin.detect_mask.data(), f.t2sin.size * sizeof(double));
    if (preamble_bytes) std::memcpy(preamble_bytes, f.preamble.preamble.data(), f.preamble.preamble.size());
    put(ofdm_preamble, f.preamble.ofdm_preamble.data(), f.preamble.ofdm_preamble.size());
    put(mod_preamble, f.preamble.mod_preamble.data(), f.preamble.mod_preamble.size());
    put(matched, f.preamble.conjected_sinh_part.data(), f.preamble.conjected_sinh_part.size());
    put(constell, f.message.Mod.constell.data(), f.message.Mod.constell.size());
}

int oc_bit_stream_converter(int out_bits, int in_bits, const uint8_t *in, int n_in, uint8_t *out) {
    Modulation m(qam4);
    std::vector<uint8_t> v(in, in + n_in);
    auto r = m.bit_stream_converter(out_bits, in_bits, v);
    std::memcpy(out, r.data(), r.size());
    return (int)r.size();
}

int oc_mod(int mod_type, const uint8_t *bytes, int n_bytes, double *points) {
    Modulation m(static_cast<::mod_type>(mod_type));
    std::vector<uint8_t> v(bytes, bytes + n_bytes);
    auto r = m.mod(v);
    put(points, r.data(), r.size());
    return (int)r.size();
}

int oc_demod(int mod_type, double *points_inout, int n_points, uint8_t *bytes) {
    Modulation m(static_cast<::mod_type>(mod_type));
    complex_vector v(n_points);
    std::memcpy((void *)v.data(), points_inout, n_points * sizeof(complex_double));
    auto r = m.demod(v);
    std::memcpy(points_inout, v.data(), n_points * sizeof(complex_double));   // demod clamps its argument
    std::memcpy(bytes, r.data(), r.size());
    return (int)r.size();
}

void oc_tx(oc_handle *h, const uint8_t *bytes, double *frame, int16_t *frame_i16) {
    FRAME_FORM &f = h->tx_frame;
    bit_vector v(bytes, bytes + f.usefull_size);
    f.write(v);                                   // tx.cpp:35 / main.cpp:39
    auto mod_data = f.get();                      // main.cpp:41
    put(frame, mod_data.data(), mod_data.size());
    if (frame_i16) {
        auto q = f.get_int16();                   // main.cpp:42
        std::memcpy(frame_i16, q.data(), q.size() * sizeof(std::complex<int16_t>));
    }
}

static complex_vector as_vec(const double *sig, long n) {
    complex_vector v((size_t)n);
    std::memcpy((void *)v.data(), sig, (size_t)n * sizeof(complex_double));
    return v;
}

int oc_t2sin_corr(oc_handle *h, const double *sig, long n, double *out) {
    auto v = as_vec(sig, n);
    auto c = h->rx_frame.t2sin.corr(v);           // main.cpp:50
    std::memcpy(out, c.data(), c.size() * sizeof(double));
    return (int)c.size();
}

int oc_find_t2sin(oc_handle *h, const double *sig, long n, int start) {
    auto v = as_vec(sig, n);
    return h->rx_frame.t2sin.find_t2sin(v, start);   // main.cpp:51
}

void oc_find_corr(oc_handle *h, const double *sig, long n, int start, double *cor) {
    auto v = as_vec(sig, n);
    h->rx_frame.preamble.find_corr(v, start);
    std::memcpy(cor, h->rx_frame.preamble.cor.data(), h->rx_frame.preamble.cor.size() * sizeof(double));
}

int oc_find_preamble(oc_handle *h, const double *sig, long n, int start) {
    auto v = as_vec(sig, n);
    return h->rx_frame.preamble.find_preamble(v, start);   // main.cpp:53 (callers add 1)
}

// main.cpp:60-80 on the samples already sitting in rx_frame.buf[t2sin.size ...]
static void demod_chain(FRAME_FORM &f, double *scal, double *synced, double *grid, double *chan,
                        double *constell_out, uint8_t *bytes) {
    double freq_shift = f.preamble.pilot_freq_sinh();                                         // main.cpp:60
    double shift_copy = freq_shift;
    f.message_with_preamble.freq_shift(freq_shift);                                           // :61
    f.message_with_preamble.cp_freq_sinh();                                                   // :62
    f.message_with_preamble.pr_phase_sinh(f.preamble.ofdm_preamble.data(), f.preamble.size);  // :63
    put(synced, f.buf.data() + f.t2sin.size, f.message_with_preamble.size);
    auto chan_char = f.preamble.chan_char_lq();                                               // :66
    auto constell = f.message.fft();                                                          // :67
    put(grid, f.message.fft_task.FFT_buf.data(), f.message.fft_task.FFT_buf.size());
    for (size_t i = 0; i < constell.size(); i++)                                              // :69-71
        constell[i] /= chan_char[i % chan_char.size()];
    put(chan, chan_char.data(), chan_char.size());
    put(constell_out, constell.data(), constell.size());
    if (scal) {
        scal[0] = shift_copy;
        scal[1] = std::arg(chan_char[0]);
        scal[2] = std::arg(chan_char[1] / chan_char[0]);
        scal[3] = 0.0;
    }
    if (bytes) {
        auto res = f.message.Mod.demod(constell);                                             // :80
        std::memcpy(bytes, res.data(), res.size());
    }
}

void oc_rx_aligned(oc_handle *h, const double *rx_samples, double *scal, double *synced, double *grid,
                   double *chan, double *constell, uint8_t *bytes) {
    FRAME_FORM &f = h->rx_frame;
    std::memcpy((void *)(f.buf.data() + f.t2sin.size), rx_samples,
                (size_t)f.message_with_preamble.size * sizeof(complex_double));               // main.cpp:55-58
    demod_chain(f, scal, synced, grid, chan, constell, bytes);
}

void oc_read(oc_handle *h, const double *frame, double *restored, uint8_t *bytes) {
    FRAME_FORM &f = h->rx_frame;
    if (restored) {
        std::memcpy((void *)f.buf.data(), frame, sizeof(complex_double) * f.buf.size());
        auto r = f.message.fft();
        put(restored, r.data(), r.size());
    }
    auto res = f.read(const_cast<double *>(frame));                                           // Frame.cpp:239-242
    std::memcpy(bytes, res.data(), res.size());
}

void oc_chan_char(oc_handle *h, const double *rx_samples, double *chan) {
    FRAME_FORM &f = h->rx_frame;
    std::memcpy((void *)(f.buf.data() + f.t2sin.size), rx_samples,
                (size_t)f.preamble.size * sizeof(complex_double));
    auto c = f.preamble.chan_char();
    put(chan, c.data(), c.size());
}

int oc_rx_stream(oc_handle *h, const int16_t *capture, long n_samples, int max_frames,
                 long *pr_begin_abs, uint8_t *bytes) {
    FRAME_FORM &f = h->rx_frame;
    const long block = (long)f.output_size * f.config["rx_buf_size"];   // SDR::rx_buf_size, sdr.hpp:141
    const long n_blocks = n_samples / block;
    long next_block = 0;
    long cur_block = -1;                                                 // block sitting at ring[output_size...]
    std::fill(f.from_sdr_int16_buf.begin(), f.from_sdr_int16_buf.end(), std::complex<int16_t>(0, 0));

    auto buf_update = [&]() -> bool {                                    // rx.cpp:73-91
        if (next_block >= n_blocks) return false;
        std::memcpy((void *)(f.from_sdr_int16_buf.data() + f.output_size), capture + 2 * next_block * block,
                    (size_t)block * sizeof(std::complex<int16_t>));
        cur_block = next_block++;
        f.form_int16_to_double();
        return true;
    };
    auto carry = [&](int threshold) {                                    // rx.cpp:149-153 / 182-186
        std::memcpy((void *)f.from_sdr_int16_buf.data(), f.from_sdr_int16_buf.data() + threshold,
                    (size_t)f.output_size * sizeof(std::complex<int16_t>));
    };

    if (!buf_update()) return 0;                                         // rx.cpp:103-112
    int pos = 0;
    const int threshold = (int)f.from_sdr_buf.size() - f.output_size;   // rx.cpp:116
    const int cycles = (int)f.config["iterations"];                      // rx.cpp:124
    int found = 0;

    for (int i = 0; i < cycles && found < max_frames; i++) {             // rx.cpp:126
        pos = f.t2sin.find_t2sin(f.from_sdr_buf, pos);                   // :133
        if (pos == -1) {                                                 // :137-145
            pos = f.output_size;
            if (!buf_update()) break;
            continue;
        }
        if (pos >= threshold) {                                          // :147-156
            pos -= threshold;
            carry(threshold);
            if (!buf_update()) break;
        }
        int preamble_begin = f.preamble.find_preamble(f.from_sdr_buf, pos) + 1;   // :158
        if (preamble_begin < -2) {                                       // :160-166
            pos += f.message.size;
            continue;
        }
        pos = preamble_begin;                                            // :168
        if (pos == -1) {                                                 // :170-178
            pos = f.output_size;
            if (!buf_update()) break;
            continue;
        }
        if (pos >= threshold + f.t2sin.size) {                           // :180-189
            pos -= threshold;
            carry(threshold);
            if (!buf_update()) break;
        }
        std::memcpy((void *)(f.buf.data() + f.t2sin.size), f.from_sdr_buf.data() + pos,
                    (size_t)(f.output_size - f.t2sin.size) * sizeof(complex_double));   // :192-196
        if (pr_begin_abs) pr_begin_abs[found] = cur_block * block + (long)pos - f.output_size;
        pos += f.message.size;                                           // :198
        demod_chain(f, nullptr, nullptr, nullptr, nullptr, nullptr,
                    bytes ? bytes + (size_t)found * f.usefull_size : nullptr);   // :200-220
        found++;
    }
    return found;
}

long oc_txrx_loop(oc_handle *h, const uint8_t *payloads, int n_frames, uint8_t *bytes_out) {
    FRAME_FORM &t = h->tx_frame, &r = h->rx_frame;
    const int us = t.usefull_size;
    long bad = 0;
    std::vector<uint8_t> tmp(us);
    bit_vector v(us);
    for (int f = 0; f < n_frames; f++) {
        const uint8_t *pay = payloads + (size_t)f * us;
        std::memcpy(v.data(), pay, us);
        t.write(v);                                                       // tx.cpp:35
        t.get_int16();                                                    // tx.cpp:37 (fills int16_buf)
        int16_t *q = (int16_t *)t.int16_buf.data();
        double *d = (double *)r.buf.data();                               // as form_int16_to_double, Frame.hpp:472-481
        for (int i = 0; i < 2 * t.output_size; i++) d[i] = (double)q[i];
        uint8_t *out = bytes_out ? bytes_out + (size_t)f * us : tmp.data();
        demod_chain(r, nullptr, nullptr, nullptr, nullptr, nullptr, out); // main.cpp:60-80
        for (int i = 0; i < us; i++) bad += out[i] != pay[i];
    }
    return bad;
}

}  // extern "C"
