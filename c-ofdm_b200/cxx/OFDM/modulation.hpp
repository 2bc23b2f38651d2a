// OFDM/modulation.hpp look-alike: the reference's Modulation interface (OFDM/modulation.hpp:11-47) on top of
// the C ABI (include/cofdm.h).  Same names, argument meaning and return types; complex<double> at the
// surface, fp32 on the device.  Differences, both documented in INTEGRATION.md:
//   * demod() does not clamp its argument in place (the reference does, modulation.cpp:68-73; no caller reads it back);
//   * a Modulation built on its own (not inside a FRAME_FORM) borrows a process-wide handle created from
//     $COFDM_CONFIG or "config/config.txt" (the path the reference apps hard-wire).
#pragma once
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "cofdm.h"

enum mod_type { bpsk = 1, qam4 = 2, qam16 = 4, qam64 = 6, qam256 = 8 };

using complex_double = std::complex<double>;
using complex_vector = std::vector<std::complex<double>>;
using complex16_vector = std::vector<std::complex<int16_t>>;
using bit_vector = std::vector<uint8_t>;

namespace cofdm_facade {
inline void check(int rc, const char *what) {
    if (rc != COFDM_OK) throw std::runtime_error(std::string(what) + ": " + cofdm_last_error());
}
inline cofdm_t *default_handle() {
    static cofdm_t *h = [] {
        const char *p = std::getenv("COFDM_CONFIG");
        cofdm_t *x = nullptr;
        check(cofdm_create(p ? p : "config/config.txt", 0, &x), "cofdm_create");
        return x;
    }();
    return h;
}
inline std::vector<float> to_f32(const complex_double *p, size_t n) {
    std::vector<float> v(2 * n);
    for (size_t i = 0; i < n; i++) { v[2 * i] = (float)p[i].real(); v[2 * i + 1] = (float)p[i].imag(); }
    return v;
}
}  // namespace cofdm_facade

class Modulation {
    cofdm_t *h_;

public:
    mod_type modulation;
    std::vector<complex_double> constell;
    size_t mod_index;

    explicit Modulation(mod_type mod, cofdm_t *h = nullptr)
        : h_(h ? h : cofdm_facade::default_handle()), modulation(mod), constell(size_t(1) << mod), mod_index(mod) {
        std::vector<uint8_t> all(constell.size());
        for (size_t i = 0; i < all.size(); i++) all[i] = (uint8_t)i;
        // table = mod() of every symbol value, one value per `mod` bits, packed MSB first
        auto packed = bit_stream_converter(8, mod_index, all);
        auto pts = this->mod(packed);
        for (size_t i = 0; i < constell.size(); i++) constell[i] = pts[i];
    }

    complex_vector mod(std::vector<uint8_t> &bin_input) {                       // modulation.cpp:39-50
        const size_t n = (bin_input.size() * 8 + mod_index - 1) / mod_index;
        std::vector<float> pts(2 * n);
        if (n) cofdm_facade::check(cofdm_mod(h_, (int)modulation, bin_input.data(), bin_input.size(), pts.data(), COFDM_HOST), "cofdm_mod");
        complex_vector out(n);
        for (size_t i = 0; i < n; i++) out[i] = complex_double(pts[2 * i], pts[2 * i + 1]);
        return out;
    }

    std::vector<uint8_t> demod(complex_vector &input) {                         // modulation.cpp:53-87
        auto f = cofdm_facade::to_f32(input.data(), input.size());
        std::vector<uint8_t> out((input.size() * mod_index + 7) / 8);
        if (!input.empty()) cofdm_facade::check(cofdm_demod(h_, (int)modulation, f.data(), input.size(), out.data(), nullptr, COFDM_HOST), "cofdm_demod");
        return out;
    }

    // host-side integer regrouping, MSB first, zero padded tail (modulation.cpp:90-125)
    std::vector<uint8_t> bit_stream_converter(size_t output_block_size, size_t input_block_size, std::vector<uint8_t> &input) {
        const size_t nbits = input.size() * input_block_size;
        std::vector<uint8_t> output(nbits / output_block_size + (nbits % output_block_size > 0), 0);
        for (size_t i = 0; i < nbits; i++) {
            const unsigned bit = (input[i / input_block_size] >> (input_block_size - 1 - i % input_block_size)) & 1u;
            output[i / output_block_size] |= (uint8_t)(bit << (output_block_size - 1 - i % output_block_size));
        }
        return output;
    }
};
