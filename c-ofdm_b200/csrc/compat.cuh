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

// ---- packed f32x2 math (Blackwell FADD2/FMUL2/FFMA2: two fp32 lanes per issue slot) --------------
// A float2 used as a PACKED PAIR carries the same quantity of two independent problems
// (.x = symbol A, .y = symbol B of the pair a warp works on).  Scalar fallback under the emulator.
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
COFDM_DEV float2 p_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
COFDM_DEV float2 p_sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
COFDM_DEV float2 p_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
COFDM_DEV float2 p_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
COFDM_DEV float2 p_fms(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, make_float2(-c.x, -c.y)); }   // a*b - c
#else
COFDM_DEV float2 p_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
COFDM_DEV float2 p_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
COFDM_DEV float2 p_mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
COFDM_DEV float2 p_fma(float2 a, float2 b, float2 c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
COFDM_DEV float2 p_fms(float2 a, float2 b, float2 c) { return make_float2(a.x * b.x - c.x, a.y * b.y - c.y); }
#endif
COFDM_DEV float2 p_bcast(float s) { return make_float2(s, s); }
COFDM_DEV float2 p_neg(float2 a) { return make_float2(-a.x, -a.y); }

// packed complex pair: two complex numbers (A, B) held as (re_A, re_B), (im_A, im_B)
struct pc { float2 re, im; };
COFDM_DEV pc make_pc(float2 a, float2 b) { pc r; r.re = make_float2(a.x, b.x); r.im = make_float2(a.y, b.y); return r; }
COFDM_DEV float2 pc_a(pc v) { return make_float2(v.re.x, v.im.x); }
COFDM_DEV float2 pc_b(pc v) { return make_float2(v.re.y, v.im.y); }
COFDM_DEV pc cadd(pc a, pc b) { pc r; r.re = p_add(a.re, b.re); r.im = p_add(a.im, b.im); return r; }
COFDM_DEV pc csub(pc a, pc b) { pc r; r.re = p_sub(a.re, b.re); r.im = p_sub(a.im, b.im); return r; }
// both members times the same complex scalar w
COFDM_DEV pc cmul(pc a, float2 w) {
    pc r;
    r.re = p_fms(a.re, p_bcast(w.x), p_mul(a.im, p_bcast(w.y)));
    r.im = p_fma(a.re, p_bcast(w.y), p_mul(a.im, p_bcast(w.x)));
    return r;
}
// member-wise complex product
COFDM_DEV pc cmul(pc a, pc w) {
    pc r;
    r.re = p_fms(a.re, w.re, p_mul(a.im, w.im));
    r.im = p_fma(a.re, w.im, p_mul(a.im, w.re));
    return r;
}
COFDM_DEV pc cconj(pc a) { pc r; r.re = a.re; r.im = p_neg(a.im); return r; }

// ---- NATURAL-layout complex math on packed f32x2 ---------------------------------------------------
// One complex number = one aligned register pair (re, im), exactly as LDS.64 / LDS.128 deliver it.  Blackwell's
// FADD2 / FMUL2 / FFMA2 take per-operand modifiers -- swap halves (.LO_HI), negate one half (.NP), broadcast a
// scalar (.F32) -- so a complex add is ONE instruction, a complex multiply TWO (FMUL2 + FFMA2), and a multiply
// by -+j folds into the add that consumes it.  Same arithmetic rate as the (re_A, re_B)/(im_A, im_B) pairing of
// `pc` below, without pairing two transforms and without the register shuffles that pairing needs after a load.
// ptxas recognises the swap / half-negate patterns written here (checked in SASS: profiles/r02_sass_*.txt).
COFDM_DEV float2 nadd(float2 a, float2 b) { return p_add(a, b); }
COFDM_DEV float2 nsub(float2 a, float2 b) { return p_add(a, make_float2(-b.x, -b.y)); }
COFDM_DEV float2 nadd_mj(float2 a, float2 b) { return p_add(a, make_float2(b.y, -b.x)); }    // a + (-j) b
COFDM_DEV float2 nadd_pj(float2 a, float2 b) { return p_add(a, make_float2(-b.y, b.x)); }    // a + (+j) b
COFDM_DEV float2 nmul(float2 a, float2 w) {                                                  // a * w
    return p_fma(a, make_float2(w.x, w.x), p_mul(make_float2(a.y, a.x), make_float2(-w.y, w.y)));
}
COFDM_DEV float2 nmulc(float2 a, float2 w) {                                                 // a * conj(w)
    return p_fma(a, make_float2(w.x, w.x), p_mul(make_float2(a.y, a.x), make_float2(w.y, -w.y)));
}
COFDM_DEV float2 nscale(float2 a, float s) { return p_mul(a, make_float2(s, s)); }
// acc + a * w  /  acc + conj(a) * w
COFDM_DEV float2 nmac(float2 acc, float2 a, float2 w) {
    return p_fma(make_float2(a.y, a.x), make_float2(-w.y, w.y), p_fma(a, make_float2(w.x, w.x), acc));
}
COFDM_DEV float2 nmac_conj(float2 acc, float2 a, float2 w) {                                 // acc + conj(a) * w
    return p_fma(make_float2(w.y, w.x), make_float2(a.y, -a.y), p_fma(w, make_float2(a.x, a.x), acc));
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

// exp(+j*2*pi*t) for moderate |t| (a few turns) in ~20 instructions: quarter-turn reduction, then minimax
// polynomials in f^2 on |f| <= 1/8 turn (fitted offline; fp32 evaluation error 1e-7 abs, i.e. rounding level).
COFDM_DEV float2 fast_cis_turns(float t) {
    const float k = rintf(4.0f * t);
    const float f = fmaf(k, -0.25f, t);
    const float u = f * f;
    const float s = f * fmaf(u, fmaf(u, fmaf(u, fmaf(u, 4.1414680329e+01f, -7.6695821688e+01f), 8.1605180747e+01f), -4.1341702049e+01f), 6.2831853070e+00f);
    const float c = fmaf(u, fmaf(u, fmaf(u, fmaf(u, 5.9220407194e+01f, -8.5442852118e+01f), 6.4939316208e+01f), -1.9739208650e+01f), 9.9999999995e-01f);
    const int q = (int)k;
    float cr = (q & 1) ? -s : c, sr = (q & 1) ? c : s;
    if (q & 2) { cr = -cr; sr = -sr; }
    return make_float2(cr, sr);
}
COFDM_DEV float2 cis_neg_turns_f(float turns) { return fast_cis_turns(-turns); }
// the same for two angles at once (packed f32x2 polynomial evaluation): result .re = (cos a, cos b), .im = (sin a, sin b)
COFDM_DEV pc fast_cis_turns2(float ta, float tb) {
    const float ka = rintf(4.0f * ta), kb = rintf(4.0f * tb);
    const float2 f = p_fma(make_float2(ka, kb), p_bcast(-0.25f), make_float2(ta, tb));
    const float2 u = p_mul(f, f);
    float2 sp = p_fma(u, p_bcast(4.1414680329e+01f), p_bcast(-7.6695821688e+01f));
    sp = p_fma(u, sp, p_bcast(8.1605180747e+01f));
    sp = p_fma(u, sp, p_bcast(-4.1341702049e+01f));
    sp = p_fma(u, sp, p_bcast(6.2831853070e+00f));
    const float2 s = p_mul(f, sp);
    float2 c = p_fma(u, p_bcast(5.9220407194e+01f), p_bcast(-8.5442852118e+01f));
    c = p_fma(u, c, p_bcast(6.4939316208e+01f));
    c = p_fma(u, c, p_bcast(-1.9739208650e+01f));
    c = p_fma(u, c, p_bcast(9.9999999995e-01f));
    const int qa = (int)ka, qb = (int)kb;
    pc r;
    float cra = (qa & 1) ? -s.x : c.x, sra = (qa & 1) ? c.x : s.x;
    float crb = (qb & 1) ? -s.y : c.y, srb = (qb & 1) ? c.y : s.y;
    if (qa & 2) { cra = -cra; sra = -sra; }
    if (qb & 2) { crb = -crb; srb = -srb; }
    r.re = make_float2(cra, crb);
    r.im = make_float2(sra, srb);
    return r;
}

// atan2(y, x) / (2*pi) in (-0.5, 0.5]: one fast division + a degree-17 odd minimax polynomial on [0,1]
// (fp32 evaluation error 2.3e-8 turns = 1.4e-7 rad, the level of atan2f itself)
COFDM_DEV float fast_atan2_turns(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = mx > 0.0f ? __fdividef(mn, mx) : 0.0f;
    const float u = a * a;
    float r = fmaf(u, 3.9099854500e-04f, -2.2920431106e-03f);
    r = fmaf(u, r, 6.3313740304e-03f);
    r = fmaf(u, r, -1.1514632549e-02f);
    r = fmaf(u, r, 1.6709593699e-02f);
    r = fmaf(u, r, -2.2538297827e-02f);
    r = fmaf(u, r, 3.1808559006e-02f);
    r = fmaf(u, r, -5.3050475890e-02f);
    r = fmaf(u, r, 1.5915492501e-01f);
    r *= a;
    if (ay > ax) r = 0.25f - r;
    if (x < 0.0f) r = 0.5f - r;
    return y < 0.0f ? -r : r;
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

// ---- named barrier for a sub-team of warps (bar.sync id, nthreads) --------------------------------
#ifdef COFDM_EMU
COFDM_DEV void named_bar_sync(int id, int nthreads) { emu::named_barrier(id, nthreads); }
#else
// id must be a compile-time constant at every call site (see team_bar_sync below)
#define named_bar_sync(id, nthreads) asm volatile("bar.sync %0, %1;" ::"n"(id), "r"(nthreads) : "memory")
#endif

// Team barriers with IMMEDIATE ids.  An SM has 64 hardware barriers; a kernel whose barrier id is a run-time
// value is charged all 16 per CTA (ptxas cannot bound it), which caps residency at 4 CTAs per SM.  With
// immediates ptxas counts max id + 1.  Ids: 0 = __syncthreads, 1 = coarse-CFO warps, 2 = channel-fit warps,
// 3 + team = FFT team.  MAXT = number of teams the instantiation can have.
#ifdef COFDM_EMU
template <int MAXT> COFDM_DEV void team_bar_sync(int team) { emu::named_barrier(3 + team, 64); }
#else
template <int ID> COFDM_DEV void bar_sync_imm(int nthreads) { asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(nthreads) : "memory"); }
template <int MAXT> COFDM_DEV void team_bar_sync(int team) {
    // a single-team kernel (the acquire kernel) gets an immediate id and therefore a small barrier count (it wants
    // 8 CTAs per SM); multi-team kernels are held to <= 4 CTAs per SM by registers / shared memory anyway, so a
    // run-time id (all 16 barriers charged) costs them nothing and saves the dispatch on `team`.
    if (MAXT == 1) bar_sync_imm<3>(64);
    else asm volatile("bar.sync %0, %1;" ::"r"(3 + team), "r"(64) : "memory");
}
#endif

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
COFDM_DEV void tma_store_fence() {}
COFDM_DEV void tma_store_1d(void *dst, const void *src, uint32_t bytes) { memcpy(dst, src, bytes); }
COFDM_DEV void tma_store_commit_and_wait_read() {}
COFDM_DEV void tma_store_commit() {}
COFDM_DEV void tma_store_wait_read() {}
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
// TMA 1-D bulk copy shared -> global (SASS: UBLKCP ... S2G).  Order: generic-proxy writes to shared memory, then
// tma_store_fence() by the writers, a barrier, then ONE thread issues the copies, commits and waits until the
// source has been read (the CTA's shared memory must outlive the copy).
COFDM_DEV void tma_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
COFDM_DEV void tma_store_1d(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
COFDM_DEV void tma_store_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
COFDM_DEV void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
COFDM_DEV void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
#endif

}  // namespace cofdmk
