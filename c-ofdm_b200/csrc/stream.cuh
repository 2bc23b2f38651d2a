// stream.cuh -- the stream receiver's acquisition loop (rx.cpp:101-235) on the device.
//
// rx.cpp is a sequential state machine over a ring of rx_buf_size+1 frames: find_t2sin from `pos`
// (Frame.hpp:150-197) -> find_preamble (Frame.cpp:338-378) -> demodulate -> pos += message.size, with a
// one-frame carry-over each time the position passes the last frame of the ring (rx.cpp:147-156,180-189) and
// a fresh SDR block whenever the ring is exhausted (buf_update, rx.cpp:73-91).  Where it looks next depends on
// what it found last, so the loop itself cannot be spread over threads -- but
//   * each of its two searches is data parallel (5 candidate sync-tone blocks at a time, one FFT-256 per warp;
//     the lags of the preamble correlation, lower half first, over 80 threads with 4 lags each), and
//   * a long capture can be cut into shards of whole SDR blocks that are scanned independently and merged
//     (two chains that start from different states coincide from the first frame both detect; see
//     c-ofdm_b200/stream.py and merge_stream_shards() in cofdm_host.cu).
// stream_scan_kernel runs ONE CTA per shard; the CTA executes the state machine of rx.cpp literally, with the
// capture resident in HBM and the ring VIRTUAL: ring sample r is capture[cur_block*block + r - out_sz] for
// r >= out_sz, and the carry-over region r < out_sz maps to where the last carry copied from (zeros before the
// first carry, rx.cpp's calloc'd buffer).  It emits the absolute preamble positions; the frames are then
// gathered and demodulated in one batch by the rx kernels.
#pragma once
#include "kernels.cuh"

namespace cofdmk {

struct StreamShard {
    long long first_sample;   // where the shard starts in the capture
    long long n_blocks;       // whole SDR blocks it spans (overlap block included)
    long long own_blocks;     // blocks it owns; frames found beyond them only serve to meet the next shard's chain
};
// frames a shard keeps detecting inside its overlap block.  The next shard starts cold at that block and may miss the first
// frame or two (its sync-tone grid is not yet aligned to the traffic); both chains coincide from the first frame both detect,
// so a short run into the overlap is enough -- the merge counts the boundaries where the chains did not meet (none in practice).
constexpr int kScanOverlapFrames = 16;

#ifndef COFDM_SCAN_WARPS
#define COFDM_SCAN_WARPS 5
#endif
constexpr int kScanWarps = COFDM_SCAN_WARPS;
constexpr int kScanThreads = 32 * kScanWarps;

COFDM_HD size_t stream_scan_smem_bytes(int cor_size, int pr_sin_len) {
    const size_t plane = (size_t)(cor_size + pr_sin_len) / 4 + 4;
    return (size_t)kScanWarps * 2 * kT2Slots * sizeof(float2) + 4 * plane * sizeof(float2) + 2 * (size_t)pr_sin_len * sizeof(float2) +
           (size_t)kScanWarps * sizeof(float) + 16;
}

// preconditions (checked by the host): t2sin_size == 256, pr_sin_len % 4 == 0, cor_size % 4 == 0
#ifndef COFDM_SCAN_MINB
#define COFDM_SCAN_MINB 6
#endif
__global__ void __launch_bounds__(kScanThreads, COFDM_SCAN_MINB)
stream_scan_kernel(const Params P, const unsigned *__restrict__ capture /* int16 I,Q pairs */,
                   const StreamShard *__restrict__ shards, int n_shards, int rx_buf_size, long long iterations,
                   long long *__restrict__ pos_out /* [n_shards][max_per_shard] */, int max_per_shard,
                   int *__restrict__ count_out /* [n_shards] */) {
    COFDM_DYN_SMEM(smem_raw);
    const int s = blockIdx.x;
    if (s >= n_shards) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = P.pr_sin_len, NC = P.cor_size, WN = NC + L, plane = WN / 4 + 4;
    float2 *fft = reinterpret_cast<float2 *>(smem_raw);                 // [kScanWarps][2][kT2Slots]
    float2 *win4 = fft + (size_t)kScanWarps * 2 * kT2Slots;                      // [4][plane]: sample idx at win4[idx & 3][idx >> 2]
    float4 *hf = reinterpret_cast<float4 *>(win4 + 4 * (size_t)plane);
    float *relv = reinterpret_cast<float *>(hf + L);
    int *first = reinterpret_cast<int *>(relv + kScanWarps);

    const StreamShard sh = shards[s];
    const unsigned *base = capture + sh.first_sample;
    const long long out_sz = P.frame_len, block = out_sz * rx_buf_size;   // SDR::rx_buf_size, sdr.hpp:141
    const long long ring = out_sz * (rx_buf_size + 1);                    // from_sdr_buf.size(), Frame.cpp:221
    const long long threshold = ring - out_sz;                            // rx.cpp:116
    const long long msg = (long long)P.ofdm_len * P.num_symb;             // message.size
    long long next_block = 0, cur_block = -1, carry_base = -1;

    auto sample = [&](long long r) -> float2 {
        if (r < 0 || r >= ring) return make_float2(0.f, 0.f);
        long long g;
        if (r < out_sz) {
            if (carry_base < 0) return make_float2(0.f, 0.f);
            g = carry_base + r;
        } else {
            g = cur_block * block + (r - out_sz);
        }
        const unsigned w = __ldg(base + g);
        return make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
    };
    auto buf_update = [&]() -> bool {                                      // rx.cpp:73-91
        if (next_block >= sh.n_blocks) return false;
        cur_block = next_block++;
        return true;
    };
    auto carry = [&]() { carry_base = cur_block * block + (threshold - out_sz); };   // rx.cpp:149-153 / 182-186

    for (int i = tid; i < L; i += kScanThreads) { const float2 hm = __ldg(&P.matched[i]); hf[i] = make_float4(hm.x, hm.y, -hm.y, hm.x); }
    int found = 0, in_overlap = 0;
    if (buf_update()) {                                                    // rx.cpp:103-112
        long long pos = 0;
        for (long long it = 0; it < iterations && found < max_per_shard && in_overlap < kScanOverlapFrames; it++) {   // rx.cpp:126
            // ---- T2SIN_FORM::find_t2sin from pos (Frame.hpp:150-197): kScanWarps blocks per round ----
            long long hit = -1;
            const long long cyc = (ring - pos) / 256;
            for (long long c0 = 0; c0 < cyc && hit < 0; c0 += kScanWarps) {
                const long long c = c0 + warp;
                float rel = 0.f;
                if (c < cyc) {
                    float2 *A = fft + (size_t)warp * 2 * kT2Slots, *B = A + kT2Slots;
                    const long long r0 = pos + c * 256;
                    if (r0 >= out_sz && r0 + 256 <= ring) {
                        // the whole block lies in the current SDR block: 8 coalesced loads, no per-sample index logic
                        const unsigned *src = base + cur_block * block + (r0 - out_sz) + lane;
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const unsigned w = __ldg(src + 32 * i);
                            A[lane + 32 * i] = make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; i++) A[lane + 32 * i] = sample(r0 + lane + 32 * i);
                    }
                    rel = t2sin_block_rel(P, A, B, lane);
                }
                if (lane == 0) relv[warp] = rel;
                __syncthreads();
                for (int w = 0; w < kScanWarps; w++)
                    if (relv[w] > P.t2_level) { hit = pos + (c0 + w) * 256; break; }
                __syncthreads();
            }
            pos = hit;                                                     // rx.cpp:133
            if (pos == -1) {                                               // :137-145
                pos = out_sz;
                if (!buf_update()) break;
                continue;
            }
            if (pos >= threshold) {                                        // :147-156
                pos -= threshold;
                carry();
                if (!buf_update()) break;
            }
            // ---- PREAMBLE_FORM::find_preamble from pos (Frame.cpp:338-378): thread t owns lags 4t..4t+3 ----
            if (tid == 0) *first = 0x7fffffff;
            for (int i = tid; i < 4 * plane; i += kScanThreads) {
                const int q = i / plane, k = i - q * plane, idx = 4 * k + q;
                win4[i] = idx < WN ? sample(pos + idx) : make_float2(0.f, 0.f);
            }
            __syncthreads();
            // only the FIRST lag above the level matters (Frame.cpp:364-376), and it sits one sync tone behind the block
            // find_t2sin returned: the lower half of the lags is searched first, the upper half only if it holds no hit
            int lag = 0x7fffffff;
            const int half_thr = (NC / 4 + 1) / 2;                         // threads (4 lags each) per half
            for (int hp = 0; hp < 2 && lag == 0x7fffffff; hp++) {
                const int t4 = tid + hp * half_thr;
                if (tid < half_thr && 4 * t4 < NC) {
                    float2 a[4];
                    float e[4];
                    corr4_lags(win4, plane, hf, L, t4, a, e);
                    const float lvl2 = P.pr_level * P.pr_level;
                    int best = 0x7fffffff;                                 // Frame.cpp:319,364: first lag with c_i > pr_level
#pragma unroll
                    for (int q = 3; q >= 0; q--)
                        if (e[q] > 1.0f && cnorm2(a[q]) > lvl2 * e[q]) best = 4 * t4 + q;
                    if (best != 0x7fffffff) atomicMin(first, best);
                }
                __syncthreads();
                lag = *first;
                __syncthreads();
            }
            const long long preamble_begin = (lag == 0x7fffffff ? -10 : pos + lag) + 1;   // rx.cpp:158
            if (preamble_begin < -2) { pos += msg; continue; }             // :160-166
            pos = preamble_begin;                                          // :168
            if (pos == -1) {                                               // :170-178
                pos = out_sz;
                if (!buf_update()) break;
                continue;
            }
            if (pos >= threshold + P.t2sin_size) {                         // :180-189
                pos -= threshold;
                carry();
                if (!buf_update()) break;
            }
            // rx.cpp:192-196: the frame's rx_len samples start at ring position pos
            if (tid == 0) pos_out[(size_t)s * max_per_shard + found] = sh.first_sample + cur_block * block + pos - out_sz;
            pos += msg;                                                    // :198
            found++;
            if (cur_block >= sh.own_blocks) in_overlap++;
        }
    }
    if (tid == 0) count_out[s] = found;
}

// frames found by the scanner -> contiguous [n][rx_len] int16 records for the rx kernels
__global__ void stream_gather_kernel(const unsigned *__restrict__ capture, long long n_samples,
                                     const long long *__restrict__ pos, int n_frames, int rx_len, unsigned *__restrict__ out) {
    const int f = blockIdx.x;
    if (f >= n_frames) return;
    const long long p = pos[f];
    unsigned *dst = out + (size_t)f * rx_len;
    for (int i = threadIdx.x; i < rx_len; i += blockDim.x) {
        const long long g = p + i;
        dst[i] = (g >= 0 && g < n_samples) ? __ldg(capture + g) : 0u;
    }
}

}  // namespace cofdmk
