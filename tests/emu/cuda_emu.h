// cuda_emu.h -- a tiny CPU thread emulator for the CUDA execution model.  TEST INFRASTRUCTURE ONLY.
//
// Purpose: there is no GPU in the build container, so the kernels in c-ofdm_b200/csrc/*.cuh are
// ALSO compiled by g++ with -DCOFDM_EMU against this header and run under `pytest -m "not gpu"`
// (tests/test_emu_kernels.py).  Every CUDA thread becomes a pthread; __syncthreads/__syncwarp are
// pthread barriers; warp shuffles go through a per-warp exchange buffer.  This checks the index
// maps, exchange layouts and arithmetic of the real kernel source on the CPU.  It is NOT a product
// path: nothing in c-ofdm_b200/ or the C ABI can reach it, it is slow by design (one OS thread per
// CUDA thread), and TMA / mbarrier / packed-f32x2 instructions are replaced by plain C++ here.
#pragma once
#include <pthread.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
#define __grid_constant__

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct short2 { short x, y; };
struct double2 { double x, y; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
static inline short2 make_short2(short x, short y) { return {x, y}; }
static inline double2 make_double2(double x, double y) { return {x, y}; }

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace emu {
struct Block {
    pthread_barrier_t bar;
    pthread_mutex_t nb_mu = PTHREAD_MUTEX_INITIALIZER;
    pthread_barrier_t nbar[16];
    bool nbar_init[16] = {};
    std::vector<pthread_barrier_t> wbar;
    std::vector<uint64_t> xbuf;   // 32 slots per warp
    unsigned char *smem;
};
inline Block *g_block = nullptr;
inline thread_local unsigned t_tid = 0;
inline unsigned char *dyn_smem() { return g_block->smem; }
}  // namespace emu

inline thread_local dim3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

static inline void __syncthreads() { pthread_barrier_wait(&emu::g_block->bar); }
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu::g_block->wbar[emu::t_tid >> 5]); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

namespace emu {
// bar.sync id, nthreads: a barrier among a fixed sub-team of the block's threads
inline void named_barrier(int id, int nthreads) {
    Block *b = g_block;
    pthread_mutex_lock(&b->nb_mu);
    if (!b->nbar_init[id]) { pthread_barrier_init(&b->nbar[id], nullptr, (unsigned)nthreads); b->nbar_init[id] = true; }
    pthread_mutex_unlock(&b->nb_mu);
    pthread_barrier_wait(&b->nbar[id]);
}

template <class T>
inline T shfl_idx(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    uint64_t *xb = &g_block->xbuf[(t_tid >> 5) * 32];
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    xb[t_tid & 31] = raw;
    __syncwarp();
    uint64_t got = xb[src & 31];
    __syncwarp();
    T r;
    std::memcpy(&r, &got, sizeof(T));
    return r;
}
}  // namespace emu

template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu::shfl_idx(v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu::shfl_idx(v, (int)(emu::t_tid & 31) ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    int lane = emu::t_tid & 31;
    return emu::shfl_idx(v, lane + (int)d < 32 ? lane + (int)d : lane);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    int lane = emu::t_tid & 31;
    return emu::shfl_idx(v, lane - (int)d >= 0 ? lane - (int)d : lane);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (emu::shfl_idx(pred ? 1 : 0, i) ? 1u : 0u) << i;
    return r;
}

static inline int __double2loint(double d) { uint64_t u; std::memcpy(&u, &d, 8); return (int)(uint32_t)u; }
static inline int __double2hiint(double d) { uint64_t u; std::memcpy(&u, &d, 8); return (int)(uint32_t)(u >> 32); }
static inline double __hiloint2double(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; std::memcpy(&d, &u, 8); return d; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    for (int o = 16; o > 0; o >>= 1) v = std::max(v, emu::shfl_idx(v, (int)(emu::t_tid & 31) ^ o));
    return v;
}
static inline int __reduce_add_sync(unsigned, int v) {
    for (int o = 16; o > 0; o >>= 1) v += emu::shfl_idx(v, (int)(emu::t_tid & 31) ^ o);
    return v;
}
static inline int __reduce_min_sync(unsigned, int v) {
    for (int o = 16; o > 0; o >>= 1) v = std::min(v, emu::shfl_idx(v, (int)(emu::t_tid & 31) ^ o));
    return v;
}

template <class T> static inline T __ldg(const T *p) { return *p; }

static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicMin(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline float atomicAdd(float *p, float v) {
    static pthread_mutex_t m = PTHREAD_MUTEX_INITIALIZER;
    pthread_mutex_lock(&m);
    float old = *p;
    *p = old + v;
    pthread_mutex_unlock(&m);
    return old;
}

static inline void sincospif(float x, float *s, float *c) { *s = (float)std::sin(M_PI * (double)x); *c = (float)std::cos(M_PI * (double)x); }
static inline void sincospi(double x, double *s, double *c) { *s = std::sin(M_PI * x); *c = std::cos(M_PI * x); }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline int __float2int_rz(float x) { return (int)x; }
static inline unsigned __float2uint_rz(float x) { return x > 0.0f ? (x >= 4294967296.0f ? 0xffffffffu : (unsigned)x) : 0u; }   // saturating, NaN -> 0
static inline int __float2int_rn(float x) { return (int)std::nearbyint(x); }
static inline float __int2float_rn(int x) { return (float)x; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
using std::max;
using std::min;

namespace emu {
// Run `f()` once per CUDA thread of a grid.  Blocks run one after another; the threads of a block
// run concurrently as pthreads.
template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F f) {
    gridDim = grid;
    blockDim = block;
    const unsigned nthr = block.x * block.y * block.z;
    const unsigned nwarp = (nthr + 31) / 32;
    if (nthr % 32) std::abort();   // keep the emulator simple: whole warps only
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                Block blk;
                pthread_barrier_init(&blk.bar, nullptr, nthr);
                blk.wbar.resize(nwarp);
                for (auto &w : blk.wbar) pthread_barrier_init(&w, nullptr, 32);
                blk.xbuf.assign((size_t)nwarp * 32, 0);
                std::vector<unsigned char> sm(smem_bytes + 256, 0xCD);   // garbage-filled like real smem
                blk.smem = (unsigned char *)(((uintptr_t)sm.data() + 127) & ~(uintptr_t)127);
                g_block = &blk;
                std::vector<std::thread> th;
                th.reserve(nthr);
                for (unsigned t = 0; t < nthr; t++)
                    th.emplace_back([&, t] {
                        t_tid = t;
                        threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        blockIdx = dim3(bx, by, bz);
                        f();
                    });
                for (auto &x : th) x.join();
                pthread_barrier_destroy(&blk.bar);
                for (auto &w : blk.wbar) pthread_barrier_destroy(&w);
                g_block = nullptr;
            }
}
}  // namespace emu
