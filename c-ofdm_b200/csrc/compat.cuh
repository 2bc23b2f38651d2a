// compat.cuh -- thin layer between the kernels and (a) real CUDA for sm_100a, (b) the CPU thread
// emulator used by the `not gpu` tests (tests/emu/cuda_emu.h, -DCOFDM_EMU).  Everything that is
// Blackwell/Hopper-specific PTX (mbarrier, cp.async.bulk = TMA 1-D bulk copy, packed f32x2 math)
// is wrapped here so the kernel bodies are ordinary C++.
#pragma once

#ifdef COFDM_EMU
#include "cuda_emu.h"
#define COFDM_DYN_SMEM(name) unsigned char *name = emu::dyn_smem()
#else
#include <cuda_runtime.h>
#include <stdint.h>
#define COFDM_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

#define COFDM_DEV __device__ __forceinline__
#define COFDM_HD __host__ __device__ inline

namespace cofdmk {

// ---- complex helpers (float2 = re, im) ---------------------------------------------------------
COFDM_DEV float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
COFDM_DEV float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
COFDM_DEV float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
COFDM_DEV float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
COFDM_DEV float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
COFDM_DEV float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
COFDM_DEV float cnorm2(float2 a) { return a.x * a.x + a.y * a.y; }
// acc += a * b
COFDM_DEV void cmac(float2 &acc, float2 a, float2 b) {
    acc.x += a.x * b.x - a.y * b.y;
    acc.y += a.x * b.y + a.y * b.x;
}
// acc += conj(a) * b
COFDM_DEV void cmac_conj(float2 &acc, float2 a, float2 b) {
    acc.x += a.x * b.x + a.y * b.y;
    acc.y += a.x * b.y - a.y * b.x;
}

// exp(-j*2*pi*turns): the angle is carried in TURNS as a double so that long ramps (thousands of
// samples times a CFO) lose nothing before the reduction to (-0.5, 0.5]; the sin/cos itself is fp32.
COFDM_DEV float2 cis_neg_turns(double turns) {
    turns -= rint(turns);
    float s, c;
    sincospif(-2.0f * (float)turns, &s, &c);
    return make_float2(c, s);
}
COFDM_DEV float2 cis_turns(double turns) {
    turns -= rint(turns);
    float s, c;
    sincospif(2.0f * (float)turns, &s, &c);
    return make_float2(c, s);
}

COFDM_DEV float2 warp_sum(float2 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}
COFDM_DEV float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
COFDM_DEV double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- mbarrier + TMA 1-D bulk copy (global -> shared) ---------------------------------------------
// SASS: UBLKCP (cp.async.bulk) + SYNCS (mbarrier).  Under the emulator the copy is a memcpy done by
// the issuing thread; callers always __syncthreads() between issue and first wait, which is also
// what makes the mbarrier initialisation visible on the GPU.
#ifdef COFDM_EMU
COFDM_DEV void mbar_init(uint64_t *bar, int) { *bar = 0; }
COFDM_DEV void mbar_fence_init() {}
COFDM_DEV void mbar_arrive_expect_tx(uint64_t *, uint32_t) {}
COFDM_DEV void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *) { memcpy(dst, src, bytes); }
COFDM_DEV void mbar_wait(uint64_t *, uint32_t) {}
#else
COFDM_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
COFDM_DEV void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
COFDM_DEV void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
COFDM_DEV void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
COFDM_DEV void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
COFDM_DEV void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif

}  // namespace cofdmk
