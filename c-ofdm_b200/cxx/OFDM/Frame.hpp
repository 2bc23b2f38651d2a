// OFDM/Frame.hpp look-alike: the part of the reference's FRAME_FORM surface that main.cpp / tx.cpp / rx.cpp
// use, implemented on the C ABI of libcofdm_b200.so (include/cofdm.h).
// std::complex<double> vectors at the surface exactly as in the reference; the device computes in fp32.
//
// The receive chain of main.cpp:60-80 / rx.cpp:200-220 is ONE fused kernel here.  The facade keeps the six
// method names and their order:
//     preamble.pilot_freq_sinh()                 runs the fused kernel on buf[t2sin.size ...] and returns the shift
//     message_with_preamble.freq_shift(shift)    no further work (the result is already cached) ...
//     message_with_preamble.cp_freq_sinh()       ...
//     message_with_preamble.pr_phase_sinh(..)    copies the fully synchronised samples into buf
//     preamble.chan_char_lq()                    returns the cached channel line
//     message.fft()                              returns the cached points, multiplied back by the channel
//                                                line so that the caller's `constell[i] /= chan[i%256]`
//                                                (main.cpp:69-71) yields the kernel's equalised points
// A caller that keeps the reference's order gets the reference's values (fp32 accuracy); calling the stages
// in another order is not supported.  FRAME_FORM::demodulate() is the direct one-call form.
//
// The receiver's ring lives on the device: form_int16_to_double() (once per SDR block, rx.cpp:89,110) widens the host
// copy the apps read from AND uploads from_sdr_int16_buf (cofdm_ring_load); find_t2sin / find_preamble / corr / find_corr
// called on from_sdr_buf then run on the resident int16 ring (COFDM_DEVICE_IN) and move only their result over PCIe.
// Called on any other vector they convert and upload that vector.
#pragma once
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../config/parser.hpp"
#include "modulation.hpp"

class FRAME_FORM;

class T2SIN_FORM {
    cofdm_t *h_;
    FRAME_FORM *frame_;

public:
    int size;
    double level;
    complex_double *buf = nullptr;                                               // Frame.hpp:87, set by FRAME_FORM (Frame.cpp:228)
    T2SIN_FORM(FRAME_FORM *f, cofdm_t *h, const cofdm_sizes &s, ConfigMap &cfg) : h_(h), frame_(f), size(s.t2sin_size), level((double)cfg["T2_sin_level"] / 1000) {}
    void set(complex_double *buf_ptr) { buf = buf_ptr; }                          // Frame.cpp:136-137 (the tone itself is written by FRAME_FORM)

    inline std::vector<double> corr(complex_vector &signal);                      // Frame.hpp:96-147
    inline int find_t2sin(complex_vector &signal, int start_index);               // Frame.hpp:150-197
};

// FFT_FORM::read of the reference (Frame.cpp:73-96): forward transforms of FFT_buf (num_symb CP-stripped symbols), pilot
// normalisation, segment correction -> restored_buf (returned by reference, valid until the next call).  write() (points ->
// time domain) has no stand-alone C-ABI entry: use OFDM_FORM/FRAME_FORM::write (bytes -> frame).
class FFT_FORM {
    cofdm_t *h_;
    cofdm_sizes s_;

public:
    int fft_size, num_data_subc, num_pilot_subc, num_symb, segment_step, segment_size;
    complex_vector FFT_buf, restored_buf;
    double norm_factor, pilot_ampl;
    FFT_FORM(cofdm_t *h, const cofdm_sizes &s, int nsymb, double ampl)
        : h_(h), s_(s), fft_size(s.fft_size), num_data_subc(s.num_data_subc), num_pilot_subc(s.num_pilot_subc), num_symb(nsymb),
          segment_step(s.num_pilot_subc ? s.num_data_subc / s.num_pilot_subc + 1 : 0), segment_size(s.num_pilot_subc ? s.num_data_subc / s.num_pilot_subc : 0),
          FFT_buf((size_t)s.fft_size * nsymb), restored_buf((size_t)s.num_data_subc * nsymb), norm_factor(std::sqrt((double)s.fft_size)), pilot_ampl(ampl) {}
    complex_vector &read() {
        if (num_symb != s_.num_symb) throw std::runtime_error("FFT_FORM::read: only the message grid (num_symb symbols) is served by the device path");
        std::vector<float> fr(2 * (size_t)s_.output_size, 0.f);                     // a frame whose message symbols carry FFT_buf (CP left zero: it is stripped)
        const size_t base = (size_t)s_.t2sin_size + (size_t)s_.ofdm_len * s_.num_pr_symb;
        for (int sy = 0; sy < num_symb; sy++)
            for (int i = 0; i < fft_size; i++) {
                const complex_double v = FFT_buf[(size_t)sy * fft_size + i];
                const size_t at = base + (size_t)sy * s_.ofdm_len + s_.cp_size + i;
                fr[2 * at] = (float)v.real(); fr[2 * at + 1] = (float)v.imag();
            }
        std::vector<uint8_t> bytes((size_t)s_.usefull_size);
        std::vector<float> rest(2 * restored_buf.size());
        cofdm_facade::check(cofdm_read_batch(h_, fr.data(), COFDM_CF32, 1, bytes.data(), nullptr, rest.data(), nullptr, COFDM_HOST), "cofdm_read_batch");
        for (size_t i = 0; i < restored_buf.size(); i++) restored_buf[i] = complex_double(rest[2 * i], rest[2 * i + 1]);
        return restored_buf;
    }
};

class OFDM_FORM {
protected:
    FRAME_FORM *frame_;

public:
    int fft_size, num_data_subc, num_pilot_subc, cp_size, num_symb, pr_sin_len;
    mod_type modType;
    int ofdm_len, size, usefull_size;
    std::vector<complex_double *> output;                                         // Frame.hpp:219: symbol s starts at output[s] (non-owning, into FRAME_FORM::buf)
    FFT_FORM fft_task;
    Modulation Mod;
    int byte_fft_size;
    OFDM_FORM(FRAME_FORM *f, cofdm_t *h, const cofdm_sizes &s, int nsymb, mod_type m, double pilot_ampl = 1.0)
        : frame_(f), fft_size(s.fft_size), num_data_subc(s.num_data_subc), num_pilot_subc(s.num_pilot_subc),
          cp_size(s.cp_size), num_symb(nsymb), pr_sin_len(s.pr_sin_len), modType(m), ofdm_len(s.ofdm_len), size(s.ofdm_len * nsymb),
          usefull_size(s.num_data_subc * nsymb), output((size_t)nsymb, nullptr), fft_task(h, s, nsymb, pilot_ampl), Mod(m, h),
          byte_fft_size(s.fft_size * (int)sizeof(complex_double)) {}
    void set(complex_double *buf_ptr) {                                           // Frame.cpp:178-182
        for (int i = 0; i < num_symb; i++) output[(size_t)i] = buf_ptr + (size_t)(cp_size + fft_size) * i;
    }
    inline void freq_shift(double &shift);                                        // Frame.hpp:340-348
    inline void cp_freq_sinh();                                                   // Frame.hpp:238-263
    inline void pr_phase_sinh(complex_double *pr, int pr_size);                   // Frame.hpp:265-274
    inline complex_vector fft();                                                  // Frame.hpp:276-282
    inline double pilot_freq_sinh();                                              // Frame.hpp:285-337
};

class PREAMBLE_FORM : public OFDM_FORM {
    cofdm_t *h_;

public:
    double level;
    bit_vector preamble;                                                          // Frame.cpp:269-272: the mt19937 bytes
    complex_vector mod_preamble, ofdm_preamble, conjected_sinh_part;
    std::vector<double> cor;                                                      // Frame.cpp:266, filled by find_corr
    complex_vector chan_est;
    PREAMBLE_FORM(FRAME_FORM *f, cofdm_t *h, const cofdm_sizes &s, ConfigMap &cfg)
        : OFDM_FORM(f, h, s, s.num_pr_symb, bpsk), h_(h), level((double)cfg["pr_level"] / 1000),
          preamble((size_t)s.num_data_subc * s.num_pr_symb / 8), mod_preamble((size_t)s.num_data_subc * s.num_pr_symb),
          ofdm_preamble((size_t)s.ofdm_len * s.num_pr_symb), conjected_sinh_part((size_t)s.pr_sin_len), cor((size_t)s.cor_size, 0.0),
          chan_est((size_t)s.num_data_subc) {
        cofdm_facade::check(cofdm_get_constants(h, nullptr, preamble.data(), (double *)ofdm_preamble.data(), (double *)mod_preamble.data(),
                                                (double *)conjected_sinh_part.data(), nullptr), "cofdm_get_constants");
    }
    inline void find_corr(complex_vector &input, int start);                      // Frame.cpp:297-335
    inline int find_preamble(complex_vector &input, int start);                   // Frame.cpp:338-378
    inline complex_vector &chan_char_lq();                                        // Frame.hpp:389-434
    inline complex_vector chan_char();                                            // Frame.hpp:375-385
};

class FRAME_FORM {
    cofdm_sizes s_{};            // declared before h_: open() fills it while h_ is being initialised
    cofdm_t *h_ = nullptr;
    // cache of the last fused run (see the header comment)
    std::vector<float> scal_, chan_, constell_, synced_;
    bit_vector bytes_;
    const int16_t *ring_dev_ = nullptr;          // from_sdr_int16_buf on the device (cofdm_ring_load), valid after form_int16_to_double()
    static cofdm_t *open(const std::string &path, cofdm_sizes &s) {
        cofdm_t *h = nullptr;
        cofdm_facade::check(cofdm_create(path.c_str(), 0, &h), "cofdm_create");
        cofdm_facade::check(cofdm_query(h, &s), "cofdm_query");
        return h;
    }

public:
    ConfigMap config;
    T2SIN_FORM t2sin;
    PREAMBLE_FORM preamble;
    OFDM_FORM message;
    OFDM_FORM message_with_preamble;
    complex_vector buf;
    complex16_vector int16_buf;
    complex_vector from_sdr_buf;
    complex16_vector from_sdr_int16_buf;
    int usefull_size, output_size;

    explicit FRAME_FORM(const std::string &CONFIGNAME)                           // Frame.cpp:213-232
        : h_(open(CONFIGNAME, s_)), config(parse_config(CONFIGNAME)), t2sin(this, h_, s_, config), preamble(this, h_, s_, config),
          message(this, h_, s_, s_.num_symb, (mod_type)s_.mod_type, (double)config["pilot_ampl"] / 1000),
          message_with_preamble(this, h_, s_, s_.num_symb + s_.num_pr_symb, (mod_type)s_.mod_type, (double)config["pilot_ampl"] / 1000),
          buf((size_t)s_.output_size), int16_buf((size_t)s_.output_size),
          from_sdr_buf((size_t)s_.output_size * (config["rx_buf_size"] + 1)), from_sdr_int16_buf(from_sdr_buf.size()),
          usefull_size(s_.usefull_size), output_size(s_.output_size) {
        // the constructor of the reference leaves the sync tone and the preamble in buf (Frame.cpp:228-229)
        std::vector<double> tone(2 * (size_t)s_.t2sin_size);
        cofdm_facade::check(cofdm_get_constants(h_, tone.data(), nullptr, nullptr, nullptr, nullptr, nullptr), "cofdm_get_constants");
        std::memcpy((void *)buf.data(), tone.data(), tone.size() * sizeof(double));
        std::memcpy((void *)(buf.data() + s_.t2sin_size), preamble.ofdm_preamble.data(), preamble.ofdm_preamble.size() * sizeof(complex_double));
        // non-owning views into buf, as Frame.cpp:227-231 sets them
        t2sin.set(buf.data());
        preamble.set(buf.data() + t2sin.size);
        message.set(buf.data() + t2sin.size + preamble.size);
        message_with_preamble.set(buf.data() + t2sin.size);
    }
    cofdm_t *handle() const { return h_; }
    // the device copy of the ring when `v` IS the frame's ring and has been loaded, else null
    const int16_t *resident(const complex_vector &v) const { return &v == &from_sdr_buf ? ring_dev_ : nullptr; }
    ~FRAME_FORM() { cofdm_destroy(h_); }
    FRAME_FORM(const FRAME_FORM &) = delete;
    FRAME_FORM &operator=(const FRAME_FORM &) = delete;

    void write(bit_vector &input) {                                              // Frame.cpp:235-237
        std::vector<float> f(2 * (size_t)output_size);
        bit_vector in(input);
        in.resize((size_t)usefull_size, 0);
        cofdm_facade::check(cofdm_tx_batch(h_, in.data(), 1, f.data(), COFDM_CF32, COFDM_HOST), "cofdm_tx_batch");
        // ONE kernel run: buf from the fp32 frame; int16_buf = trunc(sample * mult) in fp32, exactly what the kernel's own
        // int16 output stage computes (Frame.cpp:252)
        const float mult = (float)s_.mult;
        int16_t *q = (int16_t *)int16_buf.data();
        for (int i = 0; i < output_size; i++) {
            buf[i] = complex_double(f[2 * i], f[2 * i + 1]);
            q[2 * i] = (int16_t)(int)(f[2 * i] * mult); q[2 * i + 1] = (int16_t)(int)(f[2 * i + 1] * mult);
        }
    }
    complex_vector get() { return buf; }                                         // Frame.cpp:244-246
    complex16_vector get_int16() { return int16_buf; }                           // Frame.cpp:249-256 (filled by write)
    void form_int16_to_double() {                                                // Frame.hpp:472-481 (host copy of the ring)
        const int16_t *q = (const int16_t *)from_sdr_int16_buf.data();
        double *d = (double *)from_sdr_buf.data();
        for (size_t i = 0; i < 2 * from_sdr_int16_buf.size(); i++) d[i] = (double)q[i];
        cofdm_facade::check(cofdm_ring_load(h_, q, from_sdr_int16_buf.size(), &ring_dev_), "cofdm_ring_load");
    }

    // the whole chain of main.cpp:60-80 on buf[t2sin.size ...]: returns the payload bytes, fills the caches
    bit_vector demodulate() {
        const size_t n = (size_t)s_.rx_len;
        auto f = cofdm_facade::to_f32(buf.data() + s_.t2sin_size, n);
        scal_.assign(48, 0.f); chan_.assign(2 * (size_t)s_.num_data_subc, 0.f);
        constell_.assign(2 * (size_t)s_.constell_size, 0.f); synced_.assign(2 * n, 0.f);
        bytes_.assign((size_t)usefull_size, 0);
        cofdm_rx_taps taps{scal_.data(), nullptr, chan_.data(), constell_.data(), synced_.data()};
        cofdm_facade::check(cofdm_rx_aligned_batch(h_, f.data(), COFDM_CF32, 1, n, bytes_.data(), nullptr, &taps, COFDM_HOST), "cofdm_rx_aligned_batch");
        for (int i = 0; i < s_.num_data_subc; i++) preamble.chan_est[i] = complex_double(chan_[2 * i], chan_[2 * i + 1]);
        return bytes_;
    }
    double cached_shift() const { return scal_.empty() ? 0.0 : (double)scal_[0]; }
    void copy_synced_to_buf() {
        for (int i = 0; i < s_.rx_len; i++) buf[s_.t2sin_size + i] = complex_double(synced_[2 * i], synced_[2 * i + 1]);
    }
    complex_vector cached_points_times_channel() {
        complex_vector out((size_t)s_.constell_size);
        for (int i = 0; i < s_.constell_size; i++)
            out[i] = complex_double(constell_[2 * i], constell_[2 * i + 1]) * preamble.chan_est[i % s_.num_data_subc];
        return out;
    }
    const bit_vector &cached_bytes() const { return bytes_; }

    bit_vector read(void *transmitted_data) {                                    // Frame.cpp:239-242
        std::memcpy((void *)buf.data(), transmitted_data, sizeof(complex_double) * buf.size());
        auto f = cofdm_facade::to_f32(buf.data(), buf.size());
        bit_vector out((size_t)usefull_size);
        cofdm_facade::check(cofdm_read_batch(h_, f.data(), COFDM_CF32, 1, out.data(), nullptr, nullptr, nullptr, COFDM_HOST), "cofdm_read_batch");
        return out;
    }
    complex_vector chan_char_of_buf() {                                          // PREAMBLE_FORM::chan_char, Frame.hpp:375-385
        auto f = cofdm_facade::to_f32(buf.data(), buf.size());
        bit_vector tmp((size_t)usefull_size);
        std::vector<float> cc(2 * (size_t)s_.num_data_subc);
        cofdm_facade::check(cofdm_read_batch(h_, f.data(), COFDM_CF32, 1, tmp.data(), nullptr, nullptr, cc.data(), COFDM_HOST), "cofdm_read_batch");
        complex_vector out((size_t)s_.num_data_subc);
        for (int i = 0; i < s_.num_data_subc; i++) out[i] = complex_double(cc[2 * i], cc[2 * i + 1]);
        return out;
    }
};

inline double OFDM_FORM::pilot_freq_sinh() { frame_->demodulate(); return frame_->cached_shift(); }
inline void OFDM_FORM::freq_shift(double &) {}
inline void OFDM_FORM::cp_freq_sinh() {}
inline void OFDM_FORM::pr_phase_sinh(complex_double *, int) { frame_->copy_synced_to_buf(); }
inline complex_vector OFDM_FORM::fft() { return frame_->cached_points_times_channel(); }
inline complex_vector &PREAMBLE_FORM::chan_char_lq() { return chan_est; }
inline complex_vector PREAMBLE_FORM::chan_char() { chan_est = frame_->chan_char_of_buf(); return chan_est; }

inline std::vector<double> T2SIN_FORM::corr(complex_vector &signal) {
    const size_t n = signal.size();
    std::vector<double> out(size ? n / size : 0, 0.0);
    if (out.empty()) return out;
    std::vector<float> rel(out.size());
    if (const int16_t *ring = frame_->resident(signal)) {
        cofdm_facade::check(cofdm_t2sin_metric(h_, ring, COFDM_CI16, n, 0, rel.data(), COFDM_DEVICE_IN), "cofdm_t2sin_metric");
    } else {
        auto f = cofdm_facade::to_f32(signal.data(), n);
        cofdm_facade::check(cofdm_t2sin_metric(h_, f.data(), COFDM_CF32, n, 0, rel.data(), COFDM_HOST), "cofdm_t2sin_metric");
    }
    for (size_t i = 0; i < out.size(); i++) out[i] = rel[i] > (float)level ? rel[i] : 0.0;
    return out;
}
inline int T2SIN_FORM::find_t2sin(complex_vector &signal, int start_index) {
    long long pos = -1;
    if (const int16_t *ring = frame_->resident(signal)) {
        cofdm_facade::check(cofdm_find_t2sin(h_, ring, COFDM_CI16, signal.size(), (size_t)start_index, &pos, COFDM_DEVICE_IN), "cofdm_find_t2sin");
    } else {
        auto f = cofdm_facade::to_f32(signal.data(), signal.size());
        cofdm_facade::check(cofdm_find_t2sin(h_, f.data(), COFDM_CF32, signal.size(), (size_t)start_index, &pos, COFDM_HOST), "cofdm_find_t2sin");
    }
    return (int)pos;
}
inline void PREAMBLE_FORM::find_corr(complex_vector &input, int start) {
    long long st = start, first = -10;
    std::vector<float> c(cor.size());
    if (const int16_t *ring = frame_->resident(input)) {
        cofdm_facade::check(cofdm_preamble_search(h_, ring, COFDM_CI16, input.size(), &st, 1, c.data(), &first, COFDM_DEVICE_IN), "cofdm_preamble_search");
    } else {
        auto f = cofdm_facade::to_f32(input.data(), input.size());
        cofdm_facade::check(cofdm_preamble_search(h_, f.data(), COFDM_CF32, input.size(), &st, 1, c.data(), &first, COFDM_HOST), "cofdm_preamble_search");
    }
    for (size_t i = 0; i < cor.size(); i++) cor[i] = c[i];
}
inline int PREAMBLE_FORM::find_preamble(complex_vector &input, int start) {
    long long st = start, first = -10;
    if (const int16_t *ring = frame_->resident(input)) {
        cofdm_facade::check(cofdm_preamble_search(h_, ring, COFDM_CI16, input.size(), &st, 1, nullptr, &first, COFDM_DEVICE_IN), "cofdm_preamble_search");
    } else {
        auto f = cofdm_facade::to_f32(input.data(), input.size());
        cofdm_facade::check(cofdm_preamble_search(h_, f.data(), COFDM_CF32, input.size(), &st, 1, nullptr, &first, COFDM_HOST), "cofdm_preamble_search");
    }
    return (int)first;
}
