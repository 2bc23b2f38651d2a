// facade_latency.cpp -- per-frame latency of the drop-in FRAME_FORM look-alike on the receive body of rx.cpp:126-232
// (find_t2sin -> find_preamble -> copy -> six stage calls -> equalise -> demod), ring resident on the device.
// The reference's own figure for the same body is 238 us per frame (LOG.txt, FFTW3 on its author's CPU).
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iterator>

#include "OFDM/Frame.hpp"

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: facade_latency config n_iter\n"); return 2; }
    const int iters = std::atoi(argv[2]);
    FRAME_FORM tx_frame(argv[1]), rx_frame(argv[1]);
    bit_vector pay((size_t)tx_frame.usefull_size);
    for (size_t i = 0; i < pay.size(); i++) pay[i] = (uint8_t)(i * 37 + 11);
    tx_frame.write(pay);
    auto tx_data = tx_frame.get_int16();
    // one SDR block: frames back to back with 500-sample gaps
    size_t at = 300, placed = 0;
    while (at + tx_data.size() + 1000 < rx_frame.from_sdr_int16_buf.size()) {
        std::copy(tx_data.begin(), tx_data.end(), rx_frame.from_sdr_int16_buf.begin() + (long)at);
        at += tx_data.size() + 500; placed++;
    }
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    rx_frame.form_int16_to_double();                                                   // once per SDR block (rx.cpp:89)
    const double t_block = std::chrono::duration<double, std::micro>(clk::now() - t0).count();
    double us_t2 = 0, us_pr = 0, us_chain = 0;
    size_t ok = 0, frames = 0;
    for (int it = 0; it < iters; it++) {
        int pos = 0;
        for (size_t f = 0; f < placed; f++) {
            auto a = clk::now();
            pos = rx_frame.t2sin.find_t2sin(rx_frame.from_sdr_buf, pos);               // rx.cpp:133
            auto b = clk::now();
            if (pos < 0) break;
            int pr_begin = rx_frame.preamble.find_preamble(rx_frame.from_sdr_buf, pos) + 1;   // rx.cpp:161
            auto c = clk::now();
            if (pr_begin < 0) break;
            pos = pr_begin;
            std::memcpy((void *)(rx_frame.buf.data() + rx_frame.t2sin.size), (const void *)(rx_frame.from_sdr_buf.data() + pos),
                        (size_t)(rx_frame.output_size - rx_frame.t2sin.size) * sizeof(complex_double));   // rx.cpp:192-196
            pos += rx_frame.message.size;
            auto shift = rx_frame.preamble.pilot_freq_sinh();
            rx_frame.message_with_preamble.freq_shift(shift);
            rx_frame.message_with_preamble.cp_freq_sinh();
            rx_frame.message_with_preamble.pr_phase_sinh(rx_frame.preamble.ofdm_preamble.data(), rx_frame.preamble.size);
            auto chan_char = rx_frame.preamble.chan_char_lq();
            auto constell = rx_frame.message.fft();
            for (size_t j = 0; j < constell.size(); j++) constell[j] /= chan_char[j % chan_char.size()];
            auto res = rx_frame.message.Mod.demod(constell);
            auto d = clk::now();
            us_t2 += std::chrono::duration<double, std::micro>(b - a).count();
            us_pr += std::chrono::duration<double, std::micro>(c - b).count();
            us_chain += std::chrono::duration<double, std::micro>(d - c).count();
            ok += res == pay;
            frames++;
        }
    }
    std::printf("{\"frames\": %zu, \"payload_ok\": %zu, \"us_per_frame\": %.1f, \"find_t2sin_us\": %.1f, \"find_preamble_us\": %.1f, "
                "\"chain_us\": %.1f, \"form_int16_to_double_us_per_block\": %.1f, \"frames_per_block\": %zu, \"reference_us_per_frame_authors_log\": 238}\n",
                frames, ok, frames ? (us_t2 + us_pr + us_chain) / frames : 0.0, frames ? us_t2 / frames : 0.0, frames ? us_pr / frames : 0.0,
                frames ? us_chain / frames : 0.0, t_block, placed);
    return ok == frames && frames > 0 ? 0 : 1;
}
