// PTX wrappers shared by the tcgen05 / TMEM kernels of libfnm_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include "fnm_stages.cuh"

namespace fnm {
namespace tc {

constexpr int kSlabK = 32;                // fp32 elements per 128-byte swizzle row

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// Warp-uniform wait (ALL 32 lanes must call it converged): the loop branches on a vote, i.e. on a predicate ptxas knows to
// be uniform, so the code that follows stays in uniform control flow and the MMA issuer's descriptor arithmetic is kept in
// the uniform datapath (UIADD3 on uniform registers).  After the per-lane loop of mbar_wait the compiler treats the warp as
// possibly diverged, builds every descriptor in vector registers and moves it over with R2UR right before the MMA that
// needs it -- ~15 clk of exposed latency each (tools/probes/mma_rate_probe.cu: 30 clk per MMA instead of 16 at N = 32).
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "WAITW_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "vote.sync.all.pred q, p, 0xffffffff;\n\t"
        "@q bra WAITW_DONE;\n\t"
        "bra WAITW_LOOP;\n\t"
        "WAITW_DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// Non-blocking, warp-uniform test of a phase (ALL 32 lanes converged): true when the phase with `parity` has completed.
__device__ __forceinline__ bool mbar_test_warp(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "vote.sync.all.pred q, p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// warp-uniform value of a (lane-invariant) quantity, produced by REDUX so that ptxas can keep it in a uniform register
__device__ __forceinline__ uint32_t uniform_u32(uint32_t v) { return __reduce_min_sync(0xffffffffu, v); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {      // warp-uniform call, one elected lane commits
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
}
// The issue loop runs WARP-UNIFORMLY (all 32 lanes compute the same descriptors, so they stay in
// uniform registers) and one elected lane executes the instruction -- a divergent single-thread loop
// costs ~14 dependent scalar instructions (~100 clk) per MMA instead of ~4.
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::tf32
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]^T, kind::tf32
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// two floats -> one 32-bit word of bf16 (round to nearest even): `first` in the low half (the lower k index of a
// kind::f16 operand: 16-bit A elements sit in TMEM two per column, 16-bit B elements two per word of a K-major row)
__device__ __forceinline__ uint32_t pack_bf16x2(float first, float second) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(second), "f"(first));
    return r;
}
// 16 values -> TMEM as the three A operands of the "TF32 main product + bf16 corrections" scheme:
//   hi  = rna_tf32(v)            16 fp32 columns   (kind::tf32, times the raw B tile)
//   lo16 = bf16(v - hi)           8 columns of pairs (kind::f16, times bf16(B))
//   hi16 = bf16(v)                8 columns of pairs (kind::f16, times bf16(B - trunc_tf32(B)))
__device__ __forceinline__ void split_to_tmem16_corr16(uint32_t taddr_hi, uint32_t taddr_lo16, uint32_t taddr_hi16, const float* v) {
    uint32_t w[16], q[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        uint32_t h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v[j]));
        w[j] = h;
    }
    tmem_st16(taddr_hi, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) q[j] = pack_bf16x2(v[2 * j] - __uint_as_float(w[2 * j]), v[2 * j + 1] - __uint_as_float(w[2 * j + 1]));
    tmem_st8(taddr_lo16, q);
#pragma unroll
    for (int j = 0; j < 8; ++j) q[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
    tmem_st8(taddr_hi16, q);
}
// hi/lo TF32 split of 16 values straight into TMEM: hi -> columns [col, col+16), lo -> [col+lo_off, ...)
__device__ __forceinline__ void split_to_tmem16(uint32_t taddr_hi, uint32_t taddr_lo, const float* v, bool want_lo) {
    uint32_t w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        uint32_t h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v[j]));
        w[j] = h;
    }
    tmem_st16(taddr_hi, w);
    if (want_lo) {
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = __float_as_uint(v[j] - __uint_as_float(w[j]));
        tmem_st16(taddr_lo, w);
    }
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), LBO unused for swizzled K-major layouts.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;                  // leading byte offset (16 B units)
    d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset
    d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
    return d;
}
// The same descriptor as (constant high word, 32-bit low word): an issue loop then spends ONE 32-bit uniform add per MMA
// operand instead of re-assembling 64-bit descriptors (the issuing warp is latency bound on its scalar instruction stream).
constexpr uint32_t kKmajDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t kmaj_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3fffu) | (1u << 16); }
// D[tmem] (+)= A[tmem] * B[K-major SWIZZLE_128B smem, low descriptor word]^T, kind::tf32
__device__ __forceinline__ void umma_ts_lo(uint32_t d, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 bd;\n\tmov.b64 bd, {%2, %5};\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %3, p;\n\t}" ::"r"(d),
        "r"(a_tmem), "r"(bdesc_lo), "r"(idesc), "r"(accum), "r"(kKmajDescHi)
        : "memory");
}
// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, K-major A and B
__host__ __device__ inline uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 instruction descriptor, bf16 operands, fp32 accumulate, K-major A and B (K = 16 per instruction)
__host__ __device__ inline uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// streaming 16-byte load: the activation is read exactly once per pass
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// exact-erf GELU to ~1e-7 absolute: erf via Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7) on the MUFU
// pipe (one rcp, one ex2) -- the libdevice erff costs ~3x the instructions and is branchy.  The
// absolute error of x*Phi(x) stays below 1e-7*|x|, two orders under the 1e-5 parity budget.
__device__ __forceinline__ float rcp_fast(float x) {          // MUFU.RCP (1 ulp); __frcp_rn expands to a fix-up sequence with a slow-path call
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// exp(-z^2) for the GELU formulas: MUFU.EX2 with flush-to-zero (results below 2^-126 only occur for |x| > 13, where
// erfc is 0 to fp32 anyway); __expf adds a denormal-range fix-up (compare + two multiplies) per element
__device__ __forceinline__ float exp_neg_sq(float z) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z * z * -1.4426950408889634074f));
    return r;
}
__device__ __forceinline__ float gelu_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float tt = rcp_fast(fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(1.061405429f, tt, -1.453152027f);
    p = fmaf(p, tt, 1.421413741f);
    p = fmaf(p, tt, -0.284496736f);
    p = fmaf(p, tt, 0.254829592f);
    const float e = p * tt * exp_neg_sq(z);           // erfc(z) for z >= 0
    const float cdf = x >= 0.f ? 1.0f - 0.5f * e : 0.5f * e;
    return x * cdf;
}

// GELU'(x) = Phi(x) + x phi(x) with the same erf approximation; exp(-x^2/2) is shared by both terms
__device__ __forceinline__ float gelu_grad_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float tt = rcp_fast(fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(1.061405429f, tt, -1.453152027f);
    p = fmaf(p, tt, 1.421413741f);
    p = fmaf(p, tt, -0.284496736f);
    p = fmaf(p, tt, 0.254829592f);
    const float ex = exp_neg_sq(z);
    const float e = p * tt * ex;
    const float cdf = x >= 0.f ? 1.0f - 0.5f * e : 0.5f * e;
    return fmaf(x * 0.39894228040143267794f, ex, cdf);
}

// GELU(x) and GELU'(x) from ONE rcp / ex2 pair (the head's backward needs both of the same element)
__device__ __forceinline__ void gelu_both_fast(float x, float& g, float& dg) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float tt = rcp_fast(fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(1.061405429f, tt, -1.453152027f);
    p = fmaf(p, tt, 1.421413741f);
    p = fmaf(p, tt, -0.284496736f);
    p = fmaf(p, tt, 0.254829592f);
    const float ex = exp_neg_sq(z);
    const float e = p * tt * ex;
    const float cdf = x >= 0.f ? 1.0f - 0.5f * e : 0.5f * e;
    g = x * cdf;
    dg = fmaf(x * 0.39894228040143267794f, ex, cdf);
}

__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi = __uint_as_float(h);
    lo = v - hi;
}
// byte offset of the 16-byte chunk (row, chunk) inside a [rows][128 B] swizzled slab
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ void store_split4(uint8_t* hi_slab, uint8_t* lo_slab, int row, int chunk, float4 v) {
    float4 h, l;
    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
    const uint32_t off = swz(row, chunk);
    *reinterpret_cast<float4*>(hi_slab + off) = h;
    *reinterpret_cast<float4*>(lo_slab + off) = l;
}


__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// PTX wrappers below are shared helpers -------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }

// cuTensorMapEncodeTiled through the runtime's driver entry point (fnm_tc_pw2.cu); fp32 elements, strides in bytes
bool tc_make_map(CUtensorMap* m, const float* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                 const cuuint32_t* box, bool swizzle128);
bool tc_make_map_swz(CUtensorMap* m, const float* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                     const cuuint32_t* box, CUtensorMapSwizzle swizzle);
int sm_count();                           // CTAs of a persistent launch (honours fnm_set_sm_limit)
int sm_count_max();                       // SMs of the device: workspace sizing
int sm_limit_exchange(int n);
bool tc_enabled();

}  // namespace tc
}  // namespace fnm
