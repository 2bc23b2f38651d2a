// mac/mac_frame.hpp -- stand-in for the header main.cpp:18, tx.cpp:16 and rx.cpp:17 include.  The reference tree does
// NOT ship this file; class name, members and methods were recovered from the debug information of its prebuilt
// objects and from the recorded golden frame:
//   8-byte little-endian header {u16 tx_id, u16 rx_id, u16 seq_num, u16 cs} + payload of frame_len - 8 bytes;
//   cs = 16-bit sum of every byte of the frame taken with cs = 0  (golden frame: header 01 00 00 00 00 00 7E 57).
// Call sites: MAC mac(1, 0, frame.usefull_size); mac.payload; mac.write(bytes, seq); mac.read(bytes);
// mac.input_tx_id / input_rx_id / input_seq_num  (main.cpp:26-37,82,92; tx.cpp:26-34; rx.cpp:52,221,225).
// Host-side byte shuffling (3 us per frame in the reference's own LOG.txt), outside the GPU hot path.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

class MAC {
public:
    uint16_t tx_id, rx_id, seq_num = 0, cs = 0;
    uint16_t input_tx_id = 0, input_rx_id = 0, input_seq_num = 0, input_cs = 0;
    size_t header_len = 8, frame_len, payload;
    std::vector<uint8_t> mes;

    MAC(unsigned tx, unsigned rx, size_t frame_bytes)
        : tx_id((uint16_t)tx), rx_id((uint16_t)rx), frame_len(frame_bytes), payload(frame_bytes > 8 ? frame_bytes - 8 : 0), mes(frame_bytes, 0) {}

    // 16-bit sum of the frame bytes with the checksum field taken as zero
    uint16_t calc_cs() const {
        unsigned s = 0;
        for (size_t i = 0; i < mes.size(); i++)
            if (i != 6 && i != 7) s += mes[i];
        return (uint16_t)s;
    }

    // header + data (zero-padded / truncated to `payload` bytes); a seq argument of 0 keeps the running counter, which
    // advances by one per frame (rx.cpp logs it as SEQ)
    std::vector<uint8_t> write(std::vector<uint8_t> data, size_t seq) {
        if (seq) seq_num = (uint16_t)seq;
        data.resize(payload, 0);
        auto put16 = [&](size_t at, uint16_t v) { mes[at] = (uint8_t)(v & 255); mes[at + 1] = (uint8_t)(v >> 8); };
        put16(0, tx_id); put16(2, rx_id); put16(4, seq_num); put16(6, 0);
        for (size_t i = 0; i < payload; i++) mes[header_len + i] = data[i];
        cs = calc_cs();
        put16(6, cs);
        seq_num++;
        return mes;
    }

    // parses the header into input_*; returns the payload bytes.  checksum_ok() tells whether the frame is intact.
    std::vector<uint8_t> read(std::vector<uint8_t> frame) {
        frame.resize(frame_len, 0);
        auto get16 = [&](size_t at) { return (uint16_t)(frame[at] | (frame[at + 1] << 8)); };
        input_tx_id = get16(0); input_rx_id = get16(2); input_seq_num = get16(4); input_cs = get16(6);
        unsigned s = 0;
        for (size_t i = 0; i < frame.size(); i++)
            if (i != 6 && i != 7) s += frame[i];
        last_ok_ = (uint16_t)s == input_cs;
        return std::vector<uint8_t>(frame.begin() + (long)header_len, frame.end());
    }
    bool checksum_ok() const { return last_ok_; }

private:
    bool last_ok_ = false;
};
