// host_consts.hpp -- host-side (fp64) construction of everything that is constant for a config:
// config parsing, sub-carrier maps, twiddles, sync tone, preamble, matched filter, constellations.
// Product code (no oracle involved): included by cofdm_host.cu and by the emulator harness.
//
// Reference behaviour mirrored here (file:line in /root/reference):
//   config/parser.cpp:4-33      parse_config ("key = long", '#' comments, missing key -> 0)
//   OFDM/Frame.cpp:31-44        pilot / data-segment sub-carrier map
//   OFDM/Frame.cpp:99-154       T2SIN mask and tone
//   OFDM/Frame.cpp:259-294      preamble bytes (mt19937), OFDM preamble, matched filter
//   OFDM/modulation.cpp:4-36    constellations
//   OFDM/Frame.hpp:311-321      coarse-CFO window borders
#pragma once
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "params.h"

namespace cofdmk {

using ConfigMap = std::map<std::string, long>;

inline ConfigMap parse_config_file(const std::string &path) {
    std::ifstream f(path);
    if (!f.is_open()) throw std::runtime_error("Cannot open config file");   // same text as parser.cpp:6
    ConfigMap cfg;
    std::string line;
    while (std::getline(f, line)) {
        size_t a = 0, b = line.size();
        while (a < b && std::isspace((unsigned char)line[a])) a++;
        while (b > a && std::isspace((unsigned char)line[b - 1])) b--;
        if (a == b || line[a] == '#') continue;
        const size_t eq = line.find('=', a);
        if (eq == std::string::npos || eq >= b) continue;
        std::string key, val;
        for (size_t i = a; i < eq; i++) if (!std::isspace((unsigned char)line[i])) key += line[i];
        for (size_t i = eq + 1; i < b; i++) if (!std::isspace((unsigned char)line[i])) val += line[i];
        cfg[key] = std::stol(val);                                             // throws like parser.cpp:30
    }
    return cfg;
}
inline long cfg_get(const ConfigMap &c, const char *k) {
    auto it = c.find(k);
    return it == c.end() ? 0 : it->second;                                     // operator[] semantics
}

struct cd { double re, im; };
inline cd cd_mul(cd a, cd b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }

// plain O(n^2)-free host DFT: recursive radix-2 for powers of two, direct sum otherwise (n <= 8192)
inline void host_dft(std::vector<cd> &x, int sign) {
    const int n = (int)x.size();
    if (n <= 1) return;
    if (n % 2 == 0) {
        std::vector<cd> e(n / 2), o(n / 2);
        for (int i = 0; i < n / 2; i++) { e[i] = x[2 * i]; o[i] = x[2 * i + 1]; }
        host_dft(e, sign);
        host_dft(o, sign);
        for (int k = 0; k < n / 2; k++) {
            const long double ang = sign * 2.0L * 3.14159265358979323846264338327950288L * k / n;
            const cd w = {(double)cosl(ang), (double)sinl(ang)};
            const cd t = cd_mul(w, o[k]);
            x[k] = {e[k].re + t.re, e[k].im + t.im};
            x[k + n / 2] = {e[k].re - t.re, e[k].im - t.im};
        }
        return;
    }
    std::vector<cd> y(n);
    for (int k = 0; k < n; k++) {
        long double sr = 0, si = 0;
        for (int j = 0; j < n; j++) {
            const long double ang = sign * 2.0L * 3.14159265358979323846264338327950288L * ((long long)j * k % n) / n;
            sr += x[j].re * cosl(ang) - x[j].im * sinl(ang);
            si += x[j].re * sinl(ang) + x[j].im * cosl(ang);
        }
        y[k] = {(double)sr, (double)si};
    }
    x = y;
}

struct HostTables {
    Params p{};                       // sizes filled, device pointers left null
    ConfigMap cfg;
    std::vector<float2> tw_fft, tw_pf, tw_t2, t2_tone, preamble_td, matched, mod_preamble;
    std::vector<float2> constell[9];  // index = mod type 1,2,4,6,8
    std::vector<float> t2_mask;
    std::vector<int16_t> bin_map, data_bin, pilot_bin, bin_role;
    std::vector<uint8_t> preamble_bytes;
    // fp64 originals for the double-precision facade
    std::vector<cd> t2_tone_d, preamble_td_d, matched_d, mod_preamble_d, constell_d[9];
    int rx_buf_size = 0, iterations = 0;
    std::vector<uint4> lane_desc;     // rx512n.cuh: per lane, the roles of its registers after warp_fft512
    std::vector<uint2> lane_aux;      //             combinations and straggler bins
    std::vector<uint2> acq_desc;      //             acquire kernel: which phase each slot produces
    std::vector<uint4> big_roles;     // big.cuh: per thread, the roles of its 16 bins
    std::vector<uint4> big_txd;       //          per thread, where the bits of its 16 grid points sit in the symbol's payload
    std::vector<uint4> big_eq;        //          per thread, the equaliser's table of the specialised row layout
    std::vector<uint4> tx_desc;       // tx512w.cuh: per lane, what sits at the bins lane + 32 n1
    std::vector<float2> grid_conj;    // conj(tx grid of the preamble) / sqrt(N), by bin
    std::vector<float2> grid_lane;    //             the same in lane order
    bool fused512_ok = false;         // the specialised kernels apply to this config
    bool generic_ok = false;          // the any-size multi-kernel path applies to this config
    bool big_ok = false;              // the fft-4096 cluster kernels (big.cuh) apply to this config
};

inline float2 f2(cd v) { return make_float2((float)v.re, (float)v.im); }
inline std::vector<float2> twiddles(int n) {
    std::vector<float2> t((size_t)(n > 0 ? n : 0));
    for (int k = 0; k < n; k++) {
        const long double ang = -2.0L * 3.14159265358979323846264338327950288L * k / n;
        t[k] = make_float2((float)cosl(ang), (float)sinl(ang));
    }
    return t;
}

inline std::vector<cd> constellation_d(int mod) {
    std::vector<cd> t((size_t)1 << mod);
    for (int i = 0; i < (1 << mod); i++) {
        if (mod == 1) {                                     // psk(i, 5*pi/4, 2)  modulation.cpp:4-9,30-31
            const double step = M_PI * 2 / 2.0, ang = step * i + M_PI_4 * 5;
            t[i] = {std::cos(ang), std::sin(ang)};
        } else {                                            // qam(i, mod)  modulation.cpp:12-20
            const int num = 1 << (mod / 2);
            t[i] = {2.0 / (num - 1) * (double)(i % num) - 1.0, 2.0 / (num - 1) * (double)(i >> (mod / 2)) - 1.0};
        }
    }
    return t;
}

// Roles of the registers a lane holds after warp_fft512 (fft512w.cuh): lane 2 k1 + g holds mn[i] = X[k1 + 16 i + (g ? 384 : 0)]
// and ot[i] = X[k1 + 16 i + (g ? 128 : 256)].  Derived from the sub-carrier map (Frame.cpp:31-44); the kernels' hard-wired
// pilot and straggler positions (COFDM_F512_PILOTS / COFDM_F512_STRAG in rx512n.cuh) are checked against it.
inline void build_f512_roles(HostTables &T) {
    Params &p = T.p;
    static const int pil[8] = {33, 66, 99, 132, 380, 413, 446, 479};
    static const int strag[7] = {128, 129, 130, 131, 381, 382, 383};
    for (int q = 0; q < 8; q++)
        if (T.pilot_bin[q] != pil[q]) throw std::runtime_error("fft-512 sub-carrier map differs from the kernels' pilot positions");
    if (p.pf_size != 640 || p.pf_pilot_w != 41 || p.pf_border0 != 134)          // kF512Pf* in rx512n.cuh
        throw std::runtime_error("fft-512 geometry: coarse-CFO windows differ from the acquire kernel's");
    auto g_of = [](int k) { return (k >> 7) & 1; };
    auto main_of = [](int k) { return k < 128 || k >= 384; };
    auto lane_of = [&](int k) { return 2 * (k & 15) + g_of(k); };
    auto i_of = [](int k) { return (k & 127) >> 4; };
    auto seg_of = [&](int di) { return di / p.seg_size; };
    // combinations in order of first appearance over the bins
    std::map<std::vector<int>, int> combo_of;
    std::vector<int> c_seg, c_off;
    auto combo = [&](int k) {
        const int di = T.bin_map[k];
        const std::vector<int> key = {main_of(k) ? 1 : 0, g_of(k), i_of(k), seg_of(di)};
        const int off = (di < 128 ? di : di - 256) - (k & 15);
        auto it = combo_of.find(key);
        if (it == combo_of.end()) {
            it = combo_of.emplace(key, (int)c_seg.size()).first;
            c_seg.push_back(seg_of(di));
            c_off.push_back(off);
        } else if (c_off[it->second] != off) {
            throw std::runtime_error("fft-512 map: data index is not linear inside a combination");
        }
        return it->second;
    };
    const unsigned dummy = 256u << 7;                          // data index 256: a slot nobody reads
    T.lane_desc.assign(32, make_uint4(dummy | (dummy << 16), dummy | (dummy << 16), dummy | (dummy << 16), dummy | (dummy << 16)));
    T.lane_aux.assign(32, make_uint2(0, 0));
    int nstrag = 0;
    for (int k = 0; k < 512; k++) {
        const int di = T.bin_map[k];
        if (di < 0) continue;
        const unsigned d = ((unsigned)di << 7) | ((unsigned)combo(k) << 2);
        if (!main_of(k)) {
            // used bins outside mn[]: ot[0] of odd lanes (128..131), ot[7] of even lanes (381..383)
            if (nstrag >= 7 || strag[nstrag] != k || i_of(k) != (g_of(k) ? 0 : 7))
                throw std::runtime_error("fft-512 map differs from the kernels' straggler bins");
            T.lane_aux[nstrag++].y = d | ((unsigned)(k & 15) << 16);
            continue;
        }
        unsigned *u = &T.lane_desc[lane_of(k)].x;
        const int i = i_of(k);
        u[i >> 1] = (u[i >> 1] & ~(0xffffu << (16 * (i & 1)))) | (d << (16 * (i & 1)));
    }
    if (nstrag != 7) throw std::runtime_error("fft-512 map differs from the kernels' straggler bins");
    for (int q : {3, 4})                                       // the two pilots outside mn[] sit in the same ot[] registers
        if (main_of(pil[q]) || i_of(pil[q]) != (g_of(pil[q]) ? 0 : 7)) throw std::runtime_error("fft-512 map differs from the kernels' pilot registers");
    p.n_combos = (int)c_seg.size();
    if (p.n_combos > 24) throw std::runtime_error("fft-512 map: more than 24 combinations");
    for (int q = 0; q < p.n_combos; q++) T.lane_aux[q].x = (unsigned)c_seg[q] | ((unsigned)(c_off[q] & 0xffff) << 16);
    // routing bits of the lanes that hold a pilot or a straggler bin (a lane holds at most one of them)
    for (int q = 0; q < 8; q++) {
        unsigned &x = T.lane_aux[lane_of(pil[q])].x;
        if (x & 0x300u) throw std::runtime_error("fft-512 map: a lane holds two special bins");
        x |= 0x100u | ((unsigned)q << 4);
    }
    for (int q = 0; q < 7; q++) {
        unsigned &x = T.lane_aux[lane_of(strag[q])].x;
        if (x & 0x300u) throw std::runtime_error("fft-512 map: a lane holds two special bins");
        x |= 0x200u | ((unsigned)q << 4);
    }
    // acquire kernel: the first 128 data sub-carriers (segments 0..3) are mn[] of the even lanes (bins k1 + 16 i < 128) -- except bins
    // 128..131, whose phases are produced by the four phase slots that hold no data (bin 0 and the pilots 33, 66, 99).  Even lane:
    // slots mn[0..3]; odd lane: slots mn[4..7] of its even neighbour.
    T.acq_desc.assign(32, make_uint2(0, 0));
    {
        int idle = 0;
        for (int k = 0; k < 128; k++) {
            const int di = T.bin_map[k], i = i_of(k), lane = 2 * (k & 15) + (i >> 2), u4 = i & 3;
            unsigned d;
            if (di >= 0) {
                if (di >= 128) throw std::runtime_error("fft-512 map: bins 0..127 hold a data index >= 128");
                d = (unsigned)di;
            } else {
                if (idle >= 4 || T.bin_map[128 + idle] != 124 + idle) throw std::runtime_error("fft-512 map differs from the acquire kernel's straggler bins");
                d = 0x8000u | ((unsigned)idle << 8) | (unsigned)(124 + idle);
                idle++;
            }
            unsigned *u = &T.acq_desc[lane].x;
            u[u4 >> 1] |= d << (16 * (u4 & 1));
        }
        if (idle != 4) throw std::runtime_error("fft-512 map differs from the acquire kernel's straggler bins");
    }
    // transmit side: lane l feeds bins l + 32 n1 to the transform; rows n1 = 5..10 (bins 160..351) must be unused
    T.tx_desc.assign(64, make_uint4(0x40004000u, 0x40004000u, 0x40004000u, 0x40004000u));
    for (int k = 0; k < 512; k++) {
        const int m = T.bin_map[k], lane = k & 31, n1 = k >> 5;
        if (n1 >= 5 && n1 <= 10) {
            if (m != -1) throw std::runtime_error("fft-512 map: a used bin in rows 5..10 of the transmit grid");
            continue;
        }
        const int sl = n1 < 5 ? n1 : n1 - 6;
        // data: [7:0] byte offset of the symbol's bits in the staged payload, [11:8] right shift of the byte (of the 16-bit window
        // when modType does not divide 8: a 6-bit symbol may straddle two bytes)
        const int mod = p.mod_type, bit = (m >= 0 ? m : 0) * mod;
        const unsigned shf = (8 % mod) != 0 ? (unsigned)(16 - mod - (bit & 7)) : (unsigned)(8 - mod - (bit & 7));
        const unsigned d = m >= 0 ? ((unsigned)(bit >> 3) | (shf << 8)) : (m == -2 ? 0x8000u : 0x4000u);
        unsigned *u = &T.tx_desc[2 * lane].x;                  // 8 consecutive words per lane
        u[sl >> 1] = (u[sl >> 1] & ~(0xffffu << (16 * (sl & 1)))) | (d << (16 * (sl & 1)));
    }
    // conj(tx grid) / sqrt(N) in lane order: mn[0..7], the used ot[] register, padding
    T.grid_lane.assign(32 * 10, make_float2(0.f, 0.f));
    for (int lane = 0; lane < 32; lane++) {
        const int k1 = lane >> 1, g = lane & 1;
        for (int i = 0; i < 8; i++) T.grid_lane[lane * 10 + i] = T.grid_conj[k1 + 16 * i + (g ? 384 : 0)];
        T.grid_lane[lane * 10 + 8] = T.grid_conj[g ? k1 + 128 : k1 + 112 + 256];
    }
}

inline HostTables build_tables(const ConfigMap &cfg) {
    HostTables T;
    T.cfg = cfg;
    Params &p = T.p;
    p.fft_size = (int)cfg_get(cfg, "fft_size");
    p.num_data_subc = (int)cfg_get(cfg, "num_data_subc");
    p.num_pilot_subc = (int)cfg_get(cfg, "num_pilot_subc");
    p.cp_size = (int)cfg_get(cfg, "cp_size");
    p.num_symb = (int)cfg_get(cfg, "num_symb");
    p.num_pr_symb = (int)cfg_get(cfg, "num_pr_symb");
    p.pr_sin_len = (int)cfg_get(cfg, "pr_sin_len");
    p.t2sin_size = (int)cfg_get(cfg, "T2sin_size");
    p.mod_type = (int)cfg_get(cfg, "modType");
    p.mult = (float)cfg_get(cfg, "mult");
    p.pilot_ampl = (float)((double)cfg_get(cfg, "pilot_ampl") / 1000);
    p.t2_level = (float)((double)cfg_get(cfg, "T2_sin_level") / 1000);
    p.pr_level = (float)((double)cfg_get(cfg, "pr_level") / 1000);
    T.rx_buf_size = (int)cfg_get(cfg, "rx_buf_size");
    T.iterations = (int)cfg_get(cfg, "iterations");
    const int N = p.fft_size, ND = p.num_data_subc, NP = p.num_pilot_subc;
    if (N <= 0 || N > 8192 || NP <= 0 || NP > kMaxPilots || ND <= 0 || p.num_symb <= 0 || p.num_pr_symb <= 0 ||
        p.cp_size < 0 || p.t2sin_size < 0)
        throw std::runtime_error("config: unsupported sizes");
    if (!(p.mod_type == 1 || p.mod_type == 2 || p.mod_type == 4 || p.mod_type == 6 || p.mod_type == 8))
        throw std::runtime_error("config: modType must be 1, 2, 4, 6 or 8");
    p.ofdm_len = N + p.cp_size;
    p.n_sym_rx = p.num_pr_symb + p.num_symb;
    p.rx_len = p.ofdm_len * p.n_sym_rx;
    p.frame_len = p.t2sin_size + p.rx_len;
    p.seg_step = ND / NP + 1;                               // Frame.cpp:9
    p.seg_size = p.seg_step - 1;                            // Frame.cpp:10
    p.bytes_per_frame = ND * p.num_symb * p.mod_type / 8;   // Frame.cpp:223
    p.pts_per_sym = ND;
    p.cor_size = p.t2sin_size * 2 + p.pr_sin_len;           // Frame.cpp:266
    if (1 + p.seg_step * (NP / 2) > N / 2 + 1 || NP % 2)
        throw std::runtime_error("config: sub-carrier map does not fit the FFT");
    // coarse-CFO windows, evaluated in double exactly as Frame.hpp:311-321 does
    p.pf_size = p.ofdm_len * p.num_pr_symb;
    {
        const double rel_bw = double(ND + NP) / (N);
        const double rel_pilot_w = rel_bw / NP;
        p.pf_pilot_w = int(p.pf_size * rel_pilot_w);
        p.pf_border0 = int((1.0 - rel_bw - rel_pilot_w) / 2.0 * p.pf_size);
        p.pf_den = NP * p.pf_size;
        p.pf_bins512 = 512.0f / (float)p.pf_den;
        p.pf_binsN = (float)N / (float)p.pf_den;
        p.inv_pilot_norm = 1.0f / ((float)(p.num_symb * NP) * p.pilot_ampl);
    }

    // sub-carrier map (Frame.cpp:31-44): pilot follows its segment in the positive half, precedes it
    // in the negative half; data points are consumed segment by segment (Frame.cpp:59-62)
    T.bin_map.assign((size_t)N, (int16_t)-1);
    T.data_bin.assign((size_t)NP * p.seg_size, 0);
    T.pilot_bin.assign((size_t)NP, 0);
    {
        int j = 0;
        for (int pos = 1 + p.seg_size; j < NP / 2; j++, pos += p.seg_step) {
            T.pilot_bin[j] = (int16_t)pos;
            for (int e = 0; e < p.seg_size; e++) T.data_bin[(size_t)j * p.seg_size + e] = (int16_t)(pos - p.seg_size + e);
        }
        for (int pos = N - p.seg_step * (NP / 2); j < NP; j++, pos += p.seg_step) {
            T.pilot_bin[j] = (int16_t)pos;
            for (int e = 0; e < p.seg_size; e++) T.data_bin[(size_t)j * p.seg_size + e] = (int16_t)(pos + 1 + e);
        }
        for (int i = 0; i < NP * p.seg_size; i++) T.bin_map[T.data_bin[i]] = (int16_t)i;
        for (int q = 0; q < NP; q++) T.bin_map[T.pilot_bin[q]] = (int16_t)-2;
    }

    T.tw_fft = twiddles(N);
    T.tw_pf = twiddles(p.pf_size);
    T.tw_t2 = twiddles(p.t2sin_size);

    for (int m : {1, 2, 4, 6, 8}) {
        T.constell_d[m] = constellation_d(m);
        for (auto v : T.constell_d[m]) T.constell[m].push_back(f2(v));
    }

    // T2SIN mask (Frame.cpp:120-133) and tone (Frame.cpp:139-154: unnormalised backward DFT of two
    // deltas of 0.5 => s[n] = .5 e^{j2pi f1 n/N} + .5 e^{j2pi f2 n/N})
    {
        const int n = p.t2sin_size, f1 = (int)cfg_get(cfg, "T2_sin_f1"), f2b = (int)cfg_get(cfg, "T2_sin_f2");
        const int sm = (int)cfg_get(cfg, "smooth");
        T.t2_mask.assign((size_t)n, 0.f);
        T.t2_tone_d.assign((size_t)n, cd{0, 0});
        if (n > 0) {
            auto clampi = [&](int v) { return v < 0 ? 0 : (v > n - 1 ? n - 1 : v); };
            for (int i = std::max(0, f1 - sm); i <= clampi(f1 + sm); i++) T.t2_mask[i] += 1.0f;
            for (int i = std::max(0, f2b - sm); i <= clampi(f2b + sm); i++) T.t2_mask[i] += 1.0f;
            std::vector<cd> spec((size_t)n, cd{0, 0});
            if (f1 >= 0 && f1 < n) spec[f1] = {0.5, 0};
            if (f2b >= 0 && f2b < n) spec[f2b] = {0.5, 0};   // same bin twice: the second store wins, as in Frame.cpp:143-144
            host_dft(spec, +1);
            T.t2_tone_d = spec;
        }
        for (auto v : T.t2_tone_d) T.t2_tone.push_back(f2(v));
    }

    // preamble: bytes from std::mt19937(pr_seed) through uniform_int_distribution<int>(0,255)
    // (Frame.cpp:269-272; libstdc++ >= 11 maps a draw to rng() >> 24), BPSK, IFFT/sqrt(N), CP
    {
        const int nb = ND * p.num_pr_symb / 8;
        T.preamble_bytes.resize((size_t)nb);
        std::mt19937 rng((uint32_t)cfg_get(cfg, "pr_seed"));
        for (auto &b : T.preamble_bytes) b = (uint8_t)(rng() >> 24);
        const auto &tab = T.constell_d[1];
        T.mod_preamble_d.resize((size_t)nb * 8);
        for (int i = 0; i < nb * 8; i++) T.mod_preamble_d[i] = tab[(T.preamble_bytes[i >> 3] >> (7 - (i & 7))) & 1];
        T.preamble_td_d.assign((size_t)p.pf_size, cd{0, 0});
        const double amp = (double)cfg_get(cfg, "pilot_ampl") / 1000, nf = std::sqrt((double)N);
        for (int s = 0; s < p.num_pr_symb; s++) {
            std::vector<cd> grid((size_t)N, cd{0, 0});
            for (int q = 0; q < NP; q++) grid[T.pilot_bin[q]] = {amp, 0};
            for (int i = 0; i < NP * p.seg_size; i++) grid[T.data_bin[i]] = T.mod_preamble_d[(size_t)s * ND + i];
            host_dft(grid, +1);
            cd *dst = T.preamble_td_d.data() + (size_t)s * p.ofdm_len;
            for (int n = 0; n < N; n++) dst[p.cp_size + n] = {grid[n].re / nf, grid[n].im / nf};
            for (int n = 0; n < p.cp_size; n++) dst[n] = dst[n + N];
        }
        // matched filter = conj(first pr_sin_len samples) / ||.||_2  (Frame.cpp:285-293)
        T.matched_d.resize((size_t)p.pr_sin_len);
        double norm = 0;
        for (int i = 0; i < p.pr_sin_len; i++) {
            const cd v = i < p.pf_size ? T.preamble_td_d[i] : cd{0, 0};
            T.matched_d[i] = {v.re, -v.im};
            norm += v.re * v.re + v.im * v.im;
        }
        norm = std::sqrt(norm);
        for (auto &v : T.matched_d) { v.re /= norm; v.im /= norm; }
        // conj(tx grid of the first preamble symbol) / sqrt(N): Parseval form of the pr_phase_sinh correlation
        T.grid_conj.assign((size_t)N, make_float2(0.f, 0.f));
        for (int q = 0; q < NP; q++) T.grid_conj[T.pilot_bin[q]] = make_float2((float)(amp / nf), 0.f);
        for (int i = 0; i < NP * p.seg_size; i++)
            T.grid_conj[T.data_bin[i]] = make_float2((float)(T.mod_preamble_d[i].re / nf), (float)(-T.mod_preamble_d[i].im / nf));
        for (auto v : T.preamble_td_d) T.preamble_td.push_back(f2(v));
        for (auto v : T.matched_d) T.matched.push_back(f2(v));
        for (auto v : T.mod_preamble_d) T.mod_preamble.push_back(f2(v));
    }

    // radix schedules for the generic path (Stockham passes of radix 16/8/4/2/5)
    auto schedule = [](int n, int *rad, int &nr) {
        nr = 0;
        for (int r : {8, 4, 2, 5})          // radix 8 keeps twice as many threads busy per pass as radix 16 (measured faster)
            while (n % r == 0 && nr < 8 && n > 1) { rad[nr++] = r; n /= r; }
        return n == 1;
    };
    T.generic_ok = schedule(N, p.fft_radix, p.fft_nr) && schedule(p.pf_size, p.pf_radix, p.pf_nr) && p.num_pr_symb >= 1 &&
                   ND % 8 == 0 && ND % NP == 0 && p.pf_size <= 12288 && N <= 8192 && p.n_sym_rx <= kGenMaxSym;
    T.fused512_ok = (N == 512 && p.cp_size == 128 && ND == 256 && NP == 8 && p.num_pr_symb == 1 &&
                     p.num_symb >= 1 && p.num_symb <= kMaxFusedSymb && p.t2sin_size % 2 == 0 && p.pr_sin_len <= 128 * 5);
    if (T.fused512_ok) build_f512_roles(T);
    T.bin_role.assign((size_t)N, (int16_t)-1);
    for (int k = 0; k < N; k++) T.bin_role[(size_t)k] = T.bin_map[(size_t)k];
    for (int q = 0; q < NP; q++) T.bin_role[(size_t)T.pilot_bin[(size_t)q]] = (int16_t)(-2 - q);
    // big.cuh: fft 4096 / cp 1024 (ofdm_len / fft_size = 5 / 4 like the fft-512 geometry), one preamble symbol, up to 8 message
    // symbols (one CTA each, a portable cluster), at most 128 pilots and 3840 data sub-carriers per symbol
    if (N == 4096) {
        p.big_tmask = 0;
        for (int k = 0; k < N; k++) if (T.bin_role[(size_t)k] != -1) p.big_tmask |= 1 << (k >> 8);
        p.big_phmask = 0;
        for (int k = 0; k < N; k++) if (T.bin_role[(size_t)k] >= 0 && T.bin_role[(size_t)k] < ND / 2) p.big_phmask |= 1 << (k >> 8);
        p.big_dstep = (256 % p.seg_step) == 0 ? 256 / p.seg_step * p.seg_size : 0;
        T.big_txd.assign(1024, make_uint4(0, 0, 0, 0));
        for (int j = 0; j < 256; j++)
            for (int u = 0; u < 16; u++) {
                const int role = T.bin_role[(size_t)(j + 256 * u)];
                unsigned w = 2u << 24;                                          // null
                if (role <= -2) w = 1u << 24;                                   // pilot
                else if (role >= 0) w = (unsigned)((role * p.mod_type) >> 3) | (unsigned)((role * p.mod_type) & 7) << 16;
                (&T.big_txd[4 * (size_t)j + (u >> 2)].x)[u & 3] = w;
            }
        T.big_roles.assign(512, make_uint4(0, 0, 0, 0));
        for (int j = 0; j < 256; j++)
            for (int t = 0; t < 16; t++) {
                unsigned *u = &T.big_roles[2 * (size_t)j].x;          // 8 consecutive words per thread
                u[t >> 1] |= ((unsigned)(uint16_t)T.bin_role[(size_t)(j + 256 * t)]) << (16 * (t & 1));
            }
    }
    // The specialised demod instance (big_demod_kernel<.., LAY = true>): data sub-carriers only in rows 0..3 and 12..15 of the
    // 256 x 16 bin matrix, pilots also in row 4 (1920 + 128 sub-carriers: bins 1..1024 and 3072..4095), and inside each of the
    // two row groups every thread's channel-line abscissa i' advances by big_dstep per row.  The table carries, per thread, where
    // each of its 8 data rows goes and the abscissa at the head of each group (extrapolated when the head itself holds no data).
    p.big_lay = 0;
    if (N == 4096 && p.big_dstep > 0 && ND <= 3840) {
        unsigned dmask = 0;
        for (int k = 0; k < N; k++) if (T.bin_role[(size_t)k] >= 0) dmask |= 1u << (k >> 8);
        bool ok = dmask == 0xF00Fu && (unsigned)p.big_tmask == 0xF01Fu;
        T.big_eq.assign(768, make_uint4(0, 0, 0, 0));
        for (int j = 0; j < 256 && ok; j++) {
            int head[2] = {0, 0};
            for (int grp = 0; grp < 2; grp++) {
                bool have = false;
                for (int r = 0; r < 4; r++) {
                    const int t = (grp ? 12 : 0) + r, role = T.bin_role[(size_t)(j + 256 * t)];
                    unsigned word = (unsigned)(3840 + (j & 15));                          // dump slot, coefficient 0
                    if (role >= 0) {
                        const int ip = role < ND / 2 ? role : role - ND, h0 = ip - r * p.big_dstep;
                        if (have && h0 != head[grp]) ok = false;
                        head[grp] = h0; have = true;
                        word = (unsigned)role | ((unsigned)(role / p.seg_size) * 8u) << 16;
                    }
                    (&T.big_eq[3 * (size_t)j + grp].x)[r] = word;
                }
                if (head[grp] < -32768 || head[grp] > 32767) ok = false;
            }
            T.big_eq[3 * (size_t)j + 2].x = ((unsigned)head[0] & 0xffffu) | ((unsigned)head[1] << 16);
        }
        p.big_lay = ok ? 1 : 0;
    }
    T.big_ok = T.generic_ok && N == 4096 && p.cp_size == 1024 && p.num_pr_symb == 1 && p.num_symb >= 1 && p.num_symb <= 8 &&
               NP <= kMaxPilots && ND <= 3840 && ND % 8 == 0 && ND % NP == 0;
    return T;
}

}  // namespace cofdmk
