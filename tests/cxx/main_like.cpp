// main_like.cpp -- the receive half of the reference's main.cpp:48-104, written against the FRAME_FORM
// look-alike (c-ofdm_b200/cxx/OFDM/Frame.hpp).  Reads an int16 capture + the expected payload, runs the
// reference's call sequence, prints byte accuracy.  Built and run by tests/test_gpu_parity.py (gpu).
#include <cstdio>
#include <fstream>
#include <iostream>
#include <iterator>

#include "OFDM/Frame.hpp"

int main(int argc, char **argv) {
    if (argc < 4) { std::cerr << "usage: main_like config capture_i16.bin payload.bin\n"; return 2; }
    FRAME_FORM rx_frame(argv[1]);
    std::ifstream cf(argv[2], std::ios::binary), pf(argv[3], std::ios::binary);
    std::vector<char> cap((std::istreambuf_iterator<char>(cf)), {}), pay((std::istreambuf_iterator<char>(pf)), {});
    const size_t n = std::min(cap.size() / 4, rx_frame.from_sdr_int16_buf.size());
    std::memcpy((void *)rx_frame.from_sdr_int16_buf.data(), cap.data(), n * 4);

    rx_frame.form_int16_to_double();                                                                   // main.cpp:48
    auto t2_sin_corr = rx_frame.t2sin.corr(rx_frame.from_sdr_buf);                                     // :50
    auto t2_sin_begin = rx_frame.t2sin.find_t2sin(rx_frame.from_sdr_buf, 0);                           // :51
    auto pr_begin = rx_frame.preamble.find_preamble(rx_frame.from_sdr_buf, t2_sin_begin) + 1;          // :53
    std::copy(rx_frame.from_sdr_buf.begin() + pr_begin - rx_frame.t2sin.size,
              rx_frame.from_sdr_buf.begin() + pr_begin - rx_frame.t2sin.size + rx_frame.output_size, rx_frame.buf.begin());   // :55-58
    auto freq_shift = rx_frame.preamble.pilot_freq_sinh();                                             // :60
    rx_frame.message_with_preamble.freq_shift(freq_shift);                                             // :61
    rx_frame.message_with_preamble.cp_freq_sinh();                                                     // :62
    rx_frame.message_with_preamble.pr_phase_sinh(rx_frame.preamble.ofdm_preamble.data(), rx_frame.preamble.size);   // :63
    auto chan_char = rx_frame.preamble.chan_char_lq();                                                 // :66
    auto constell = rx_frame.message.fft();                                                            // :67
    for (size_t i = 0; i < constell.size(); i++) constell[i] /= chan_char[i % chan_char.size()];       // :69-71
    auto res = rx_frame.message.Mod.demod(constell);                                                   // :80

    size_t ok = 0, hits = 0;
    for (auto v : t2_sin_corr) hits += v > 0;
    for (size_t i = 0; i < res.size() && i < pay.size(); i++) ok += res[i] == (uint8_t)pay[i];
    std::printf("t2_hits %zu t2_sin_begin %d pr_begin %d shift %.10f bytes_ok %zu of %zu\n", hits, t2_sin_begin, pr_begin, freq_shift, ok, res.size());
    return ok == res.size() && res.size() == pay.size() ? 0 : 1;
}
