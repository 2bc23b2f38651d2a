// txrx_like.cpp -- the transmit loop of the reference's tx.cpp:26-40 and the per-frame receive body of rx.cpp:200-232,
// written against the drop-in headers (c-ofdm_b200/cxx: OFDM/Frame.hpp, mac/mac_frame.hpp) with the SDR replaced by a
// memory copy: file -> MAC -> FRAME_FORM::write -> get_int16 -> [air] -> form_int16_to_double -> find_t2sin ->
// find_preamble -> the six rx stage calls -> demod -> MAC::read -> file.  Built and run by tests/test_gpu_parity.py (gpu).
#include <cstdio>
#include <fstream>
#include <iostream>
#include <iterator>

#include "OFDM/Frame.hpp"
#include "mac/mac_frame.hpp"

int main(int argc, char **argv) {
    if (argc < 4) { std::cerr << "usage: txrx_like config in_file out_file\n"; return 2; }
    FRAME_FORM tx_frame(argv[1]);                                            // tx.cpp:23
    FRAME_FORM rx_frame(argv[1]);                                            // rx.cpp:51
    MAC mac(1, 0, tx_frame.usefull_size), rmac(1, 0, rx_frame.usefull_size);  // tx.cpp:26, rx.cpp:52
    bit_vector origin_mes(mac.payload);                                      // tx.cpp:29
    FILE *file = std::fopen(argv[2], "rb"), *res_file = std::fopen(argv[3], "wb");
    if (!file || !res_file) return 2;
    size_t frames = 0, bad_cs = 0, bad_seq = 0, got;
    while ((got = std::fread(origin_mes.data(), 1, origin_mes.size(), file))) {        // tx.cpp:32
        for (size_t i = got; i < origin_mes.size(); i++) origin_mes[i] = 0;
        auto tx_mac_frame = mac.write(origin_mes, 0);                                  // tx.cpp:34
        tx_frame.write(tx_mac_frame);                                                  // tx.cpp:35
        auto tx_data = tx_frame.get_int16();                                           // tx.cpp:37
        // [air] the frame lands 300 samples into the receiver's ring (rx.cpp:83,106 fill from_sdr_int16_buf)
        std::fill(rx_frame.from_sdr_int16_buf.begin(), rx_frame.from_sdr_int16_buf.end(), std::complex<int16_t>(0, 0));
        std::copy(tx_data.begin(), tx_data.end(), rx_frame.from_sdr_int16_buf.begin() + 300);
        rx_frame.form_int16_to_double();                                               // rx.cpp:129
        auto t2 = rx_frame.t2sin.find_t2sin(rx_frame.from_sdr_buf, 0);                 // rx.cpp:133
        if (t2 < 0) { std::printf("frame %zu: no sync tone\n", frames); return 1; }
        auto pr_begin = rx_frame.preamble.find_preamble(rx_frame.from_sdr_buf, t2) + 1;   // rx.cpp:161
        if (pr_begin <= 0) { std::printf("frame %zu: no preamble\n", frames); return 1; }
        std::copy(rx_frame.from_sdr_buf.begin() + pr_begin - rx_frame.t2sin.size,
                  rx_frame.from_sdr_buf.begin() + pr_begin - rx_frame.t2sin.size + rx_frame.output_size, rx_frame.buf.begin());   // rx.cpp:195-198
        auto freq_shift = rx_frame.preamble.pilot_freq_sinh();                         // rx.cpp:202
        rx_frame.message_with_preamble.freq_shift(freq_shift);                         // :203
        rx_frame.message_with_preamble.cp_freq_sinh();                                 // :206
        rx_frame.message_with_preamble.pr_phase_sinh(rx_frame.preamble.ofdm_preamble.data(), rx_frame.preamble.size);   // :207
        auto chan_char = rx_frame.preamble.chan_char_lq();                             // :211
        auto constell = rx_frame.message.fft();                                        // :212
        for (size_t j = 0; j < constell.size(); j++) constell[j] /= chan_char[j % chan_char.size()];   // :214-216
        auto res_ofdm = rx_frame.message.Mod.demod(constell);                          // :220
        auto res = rmac.read(res_ofdm);                                                // :221
        bad_cs += !rmac.checksum_ok();
        bad_seq += rmac.input_seq_num != (uint16_t)frames;                             // rx.cpp:225 logs SEQ
        std::fwrite(res.data(), 1, std::min(res.size(), got), res_file);               // rx.cpp:231
        frames++;
    }
    std::fclose(file); std::fclose(res_file);
    std::printf("frames %zu bad_cs %zu bad_seq %zu from %u to %u\n", frames, bad_cs, bad_seq, (unsigned)rmac.input_tx_id, (unsigned)rmac.input_rx_id);
    return bad_cs || bad_seq ? 1 : 0;
}
