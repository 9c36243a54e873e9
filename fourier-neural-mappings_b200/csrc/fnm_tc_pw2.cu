// tc_pw2 -- the fused "1x1 conv + inverse DFT along y + bias (+ GELU' multiply) (+ GELU copy)" pass on
// TMA + tcgen05 (3xTF32).  It reads an activation once and writes one once:
//
//   out[row][y][n] = ( sum_k X[row][y][k] Wk[k][n] + sum_l by[y][l] Z[row][l][n] + bias[n] ) * GELU'(mul[row][y][n])
//   out_act        = GELU(out)                                                   (optional second output)
//
// Transposed formulation: output channels are the MMA M dimension (TMEM lanes), the positions of one
// grid-row tile are N:      D[n][y] = [Wk^T | Z[row]^T] (A, TMEM)  x  [X | by] (B, shared memory, K-major)
//   A1 = Wk^T  hi/lo : written to TMEM once per CTA (constant)
//   A2 = Z^T   hi/lo : per grid row; the row's Z tile is TMA-loaded, transposed through registers by the
//                      Z warpgroup (lane = channel) and written to TMEM with tcgen05.st
//   B1 = X tile, B2 = by tile : TMA (SWIZZLE_128B) straight into MMA-ready slabs.  The tensor core reads
//                      fp32 smem operands as TF32 by ignoring 13 mantissa bits, so the RAW tile is the
//                      hi operand; converter warps add the remainder tile lo = v - trunc(v) (smem -> smem)
//   D  : 2 x NT fp32 columns in TMEM; 8 epilogue warps (lane = channel, 128-byte coalesced stores)
// Per tile (fp32-parity mode): D = A_hi*B_raw + A_lo*B_raw + A_hi*B_lo, all kind::tf32 (72 MMAs per 64-position tile at
// C = 128).  FNM_MODE_TF32 issues A_hi*B_raw only.
// FNM_PW2_CORR=bf16 (opt-in, measured and NOT the default): the two correction products as kind::f16 MMAs on bf16 operands,
// D = A_hi*B_raw + bf16(A_lo)*bf16(B) + bf16(A)*bf16(B_lo) (K = 16 per instruction: 48 MMAs per tile; the converter warps
// write bf16(B) | bf16(B - trunc(B)) in place of the fp32 remainder tile, and the table's pair can come precomputed through
// fnm_plan.by16).  It is ~8 % faster at config 5 (0.59 vs 0.64 ms forward) but each product then carries ~1e-6 relative error
// instead of ~1.7e-7, and on outputs with cancellation (golden model fnd2d: 1.7e-5) that breaks the 1e-5 parity bar.
//
// Warps (18 -> 5 per SM sub-partition, 96 registers):
//   0 TMA producer | 1 TMEM allocator + MMA issuer | 2-5 Z warpgroup | 6-9 converters | 10-17 epilogue
#include "fnm_tc_ptx.cuh"

#include <cstdlib>

namespace fnm {
namespace tc {

constexpr int kPwThreads = 18 * 32;
constexpr int kPwMmaWarp = 1;
constexpr int kPwMaxStages = 6;
constexpr uint32_t kPwSmemBudget = 226 * 1024;

struct PwArgs2 {
    const float* wc; const float* bias; const float* mul; float* out; float* out_act;
    long long rows;
    int S2, LY, CI, CO, wc_is_kn, single_pass;       // CI / CO: FOLDED channel extents F*CIo / F*COo (see pw2_fold)
    int CIo, COo, F;                                 // the caller's channel counts; F grid rows share one lane set
    int NT, tiles_per_row, seg_tiles, nseg, units;
    int K1slabs, K2slabs, LYp, stages, nlo, nz, debug;
    uint32_t stage_bytes, x_bytes, z_bytes;
    int corr16, k1s16, k2s16;                        // bf16 correction passes; 64-channel slabs of the bf16 X / by tiles
    int by16;                                        // the bf16 pair of the by tile comes from a precomputed table (TMA)
    uint32_t lo_bytes;                               // one remainder buffer: fp32 lo tile, or {bf16(B) | bf16(B_lo)}
    int mul_raw, out_grad;                           // FNM_PW_MUL_IS_GRAD / FNM_PW_OUT_IS_GRAD (see include/fnm_b200.h)
    // F2V local branch (fnm_local_functional_*): wsum = 1: out[row][seg][half][o] = wx[x] sum_y wy[y] GELU(v) (nothing spatial
    // is written); wsum = 2: out = gS[b][o] wx[x] wy[y] GELU'(v).  row = b * P1w + x.
    int wsum, P1w;
    const float* wx; const float* wy; const float* gS;
    uint32_t tmem0;                                  // always 0: the TMEM base as a kernel parameter (see the MMA issuer)
};

// Optional per-role timeline of CTA 0 (build with -DFNM_PW2_TRACE; tools/pw2_trace.py reads it).  Never in the product build.
#ifdef FNM_PW2_TRACE
constexpr int kTraceTiles = 96, kTraceEv = 24;
__device__ long long g_pw2_trace[kTraceTiles * kTraceEv];
#define PW2_TRACE(tile_idx, ev)                                                                     \
    do {                                                                                            \
        if (blockIdx.x == 0 && (tile_idx) < (uint32_t)kTraceTiles && (threadIdx.x & 31) == 0)       \
            g_pw2_trace[(tile_idx) * kTraceEv + (ev)] = clock64();                                  \
    } while (0)
#else
#define PW2_TRACE(tile_idx, ev) do { } while (0)
#endif

// TMEM column map
constexpr uint32_t kA1Hi = 0, kA1Lo = 128, kA2Hi = 256, kA2Lo = 320, kD0 = 384;
// bf16 corrections: the remainder region of each operand holds two half-width operands, bf16(A_lo) | bf16(A)
constexpr uint32_t kA1Lo16 = 128, kA1Hi16 = 192;


// ---- MMA issue: the issuing warp is latency bound on its scalar (uniform-datapath) instruction stream,
// so the per-tile sequence is fully unrolled for the common (channel slabs, mode k-steps) shapes: one
// 32-bit add per descriptor / TMEM address, no loop control, no descriptor re-assembly.  The high word of
// a K-major SWIZZLE_128B descriptor is constant; k-steps and slabs only move the 14-bit start address.
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3fffu) | (1u << 16); }
__device__ __forceinline__ void umma_ts32(uint32_t d, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 bd;\n\tmov.b64 bd, {%2, %5};\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %3, p;\n\t}" ::"r"(d),
        "r"(a_tmem), "r"(bdesc_lo), "r"(idesc), "r"(accum), "r"(kDescHi)
        : "memory");
}

// raw passes of one tile: D = (A_hi [+ A_lo]) * B_raw over K1S channel slabs and NKS2 mode k-steps.
// K1S / NKS2 < 0: runtime extents (generic shapes).
// The Z^T operand (A2) of a new grid row is awaited only after the channel part has been issued (a2_bar != nullptr
// on the row's first tile): the Z warpgroup converts the next row while those MMAs keep the tensor pipe busy.
// `early` runs before the last four k-steps: the issuer does there the barrier waits of what it issues NEXT, so that the
// ~90 clk of each mbarrier wait pass while the tensor pipe still has queued MMAs.  (The MMA queue is only a few
// instructions deep: with every wait at the top of a tile the pipe idled ~1600 clk per 2304-clk tile -- per-role
// timeline in DESIGN.md section 4.)
template <int K1S, int NKS2, bool SINGLE, typename Early>
__device__ __forceinline__ void pw2_issue_raw(uint32_t d, uint32_t tb, uint32_t a2hi, uint32_t a2lo, uint32_t x_lo,
                                              uint32_t by_lo, uint32_t slab16, uint32_t idesc, int k1s_rt, int nks2_rt,
                                              uint64_t* a2_bar, uint32_t a2_par, Early&& early) {
    const int k1s = K1S >= 0 ? K1S : k1s_rt, nks2 = NKS2 >= 0 ? NKS2 : nks2_rt;
    const int total = 4 * k1s + nks2, e_at = total > 4 ? total - 4 : 0;
    uint32_t accum = 0;
#pragma unroll
    for (int kb = 0; kb < (K1S >= 0 ? K1S : 4); ++kb) {
        if (kb < k1s) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                if (4 * kb + ks == e_at) early();
                const uint32_t b = x_lo + kb * slab16 + 2 * ks;
                umma_ts32(d, tb + kA1Hi + kb * 32 + ks * 8, b, idesc, accum);
                accum = 1;
                if (!SINGLE) umma_ts32(d, tb + kA1Lo + kb * 32 + ks * 8, b, idesc, 1);
            }
        }
    }
    if (a2_bar) {
        mbar_wait_warp(a2_bar, a2_par);
        tc_fence_after();
    }
#pragma unroll
    for (int k8 = 0; k8 < (NKS2 >= 0 ? NKS2 : 8); ++k8) {
        if (k8 < nks2) {
            if (4 * k1s + k8 == e_at) early();
            const uint32_t b = by_lo + (k8 >> 2) * slab16 + 2 * (k8 & 3);
            umma_ts32(d, a2hi + k8 * 8, b, idesc, accum);
            accum = 1;
            if (!SINGLE) umma_ts32(d, a2lo + k8 * 8, b, idesc, 1);
        }
    }
}

// remainder pass of one tile: D += A_hi * B_lo
template <int K1S, int NKS2, typename Early>
__device__ __forceinline__ void pw2_issue_lo(uint32_t d, uint32_t tb, uint32_t a2hi, uint32_t x_lo, uint32_t by_lo,
                                             uint32_t slab16, uint32_t idesc, int k1s_rt, int nks2_rt, Early&& early) {
    const int k1s = K1S >= 0 ? K1S : k1s_rt, nks2 = NKS2 >= 0 ? NKS2 : nks2_rt;
    const int total = 4 * k1s + nks2, e_at = total > 8 ? total - 8 : 0;
#pragma unroll
    for (int kb = 0; kb < (K1S >= 0 ? K1S : 4); ++kb) {
        if (kb < k1s) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                if (4 * kb + ks == e_at) early();
                umma_ts32(d, tb + kA1Hi + kb * 32 + ks * 8, x_lo + kb * slab16 + 2 * ks, idesc, 1);
            }
        }
    }
#pragma unroll
    for (int k8 = 0; k8 < (NKS2 >= 0 ? NKS2 : 8); ++k8) {
        if (k8 < nks2) {
            if (4 * k1s + k8 == e_at) early();
            umma_ts32(d, a2hi + k8 * 8, by_lo + (k8 >> 2) * slab16 + 2 * (k8 & 3), idesc, 1);
        }
    }
}

__device__ __forceinline__ void umma_ts16(uint32_t d, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 bd;\n\tmov.b64 bd, {%2, %5};\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %3, p;\n\t}" ::"r"(d),
        "r"(a_tmem), "r"(bdesc_lo), "r"(idesc), "r"(accum), "r"(kDescHi)
        : "memory");
}

// bf16 correction passes of one tile: D += bf16(A_lo) * bf16(B) + bf16(A) * bf16(B_lo), K = 16 per instruction.
// N16_1 / N16_2: k-steps over the channel / mode axis (< 0: runtime).  xh / byh / xl / byl: descriptor low words of the
// bf16(B) X and by tiles and of the bf16(B_lo) ones; 64 elements (4 k-steps) per slab.
template <int N16_1, int N16_2, typename Early>
__device__ __forceinline__ void pw2_issue_lo16(uint32_t d, uint32_t tb, uint32_t a2lo16, uint32_t a2hi16, uint32_t xh, uint32_t byh,
                                               uint32_t xl, uint32_t byl, uint32_t slab16, uint32_t idesc, int n1_rt, int n2_rt,
                                               Early&& early) {
    const int n1 = N16_1 >= 0 ? N16_1 : n1_rt, n2 = N16_2 >= 0 ? N16_2 : n2_rt;
    const int total = 2 * (n1 + n2), e_at = total > 8 ? total - 8 : 0;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const uint32_t a1 = tb + (pass ? kA1Hi16 : kA1Lo16), a2 = pass ? a2hi16 : a2lo16;
        const uint32_t bx = pass ? xl : xh, bb = pass ? byl : byh;
#pragma unroll
        for (int j = 0; j < (N16_1 >= 0 ? N16_1 : 8); ++j) {
            if (j < n1) {
                if (pass * (n1 + n2) + j == e_at) early();
                umma_ts16(d, a1 + j * 8, bx + (j >> 2) * slab16 + 2 * (j & 3), idesc, 1);
            }
        }
#pragma unroll
        for (int j = 0; j < (N16_2 >= 0 ? N16_2 : 4); ++j) {
            if (j < n2) {
                if (pass * (n1 + n2) + n1 + j == e_at) early();
                umma_ts16(d, a2 + j * 8, bb + (j >> 2) * slab16 + 2 * (j & 3), idesc, 1);
            }
        }
    }
}

struct PwBars {
    uint64_t *full, *empty, *lo_full, *lo_empty, *a2_full, *a2_empty, *acc_full, *acc_empty;
};

// The MMA issuer's loop over this CTA's tiles, specialised on the shape.  Barrier waits are software pipelined: a tile
// starts with its accumulator and stage already awaited (by the previous tile's `early` hook), awaits its remainder tile
// before the last raw k-steps and the NEXT tile's accumulator / stage before the last remainder k-steps.
template <int K1S, int NKS2, int MODE>
__device__ __forceinline__ void pw2_mma_loop(const PwArgs2& a, const PwBars& B, uint32_t sbase, uint32_t lo_base0, int my_units,
                                             uint32_t na2, bool has_z) {
    constexpr bool SINGLE = MODE == 1, CORR16 = MODE == 2;
    const uint32_t idesc = make_idesc(128, a.NT), idesc16 = make_idesc_bf16(128, a.NT);
    const int n16_1 = 2 * a.K1slabs, n16_2 = (a.LYp / 8 + 1) / 2;     // bf16 k-steps (16 elements) over channels / modes
    const uint32_t byh_off = ((uint32_t)a.k1s16 * a.NT * 128u) >> 4, half16 = (a.lo_bytes / 2) >> 4;
    const int nks2 = a.LYp / 8;                                   // k-steps over the (padded) mode axis
    const uint32_t slab16 = (uint32_t)a.NT * 8u;                  // slab stride in 16-byte descriptor units
    const uint32_t xb16 = a.x_bytes >> 4;
    // tiles of this CTA (for "is there a next tile")
    uint32_t total = 0;
    {
        int unit = (int)blockIdx.x;
        for (int u = 0; u < my_units; ++u, unit += (int)gridDim.x) {
            const int row = unit / a.nseg, seg = unit - row * a.nseg;
            const int t0 = seg * a.seg_tiles;
            total += (uint32_t)(min(a.tiles_per_row, t0 + a.seg_tiles) - t0);
        }
    }
    // stage / phase counters advance incrementally (a `t % stages` with a run-time divisor is a vector-datapath sequence
    // and drags every value derived from it out of the uniform registers)
    uint32_t t = 0, s = 0, sph = 0;
    bool tile_ready = false;                                      // this tile's accumulator and stage were seen complete early
    int unit = (int)blockIdx.x;
    for (int u = 0; u < my_units; ++u, unit += (int)gridDim.x) {
        int t0, t1;
        {
            const int row = unit / a.nseg, seg = unit - row * a.nseg;
            t0 = seg * a.seg_tiles;
            t1 = min(a.tiles_per_row, t0 + a.seg_tiles);
        }
        const uint32_t ub = na2 == 2 ? ((uint32_t)u & 1u) : 0u, upar = (na2 == 2 ? ((uint32_t)u >> 1) : (uint32_t)u) & 1u;
        const uint32_t a2off = kA2Hi + (na2 == 2 ? 64u * ub : 0u);
        for (int tile = t0; tile < t1; ++tile, ++t) {
            uint64_t* a2_bar = (has_z && tile == t0) ? &B.a2_full[ub] : nullptr;
            const uint32_t acc = t & 1;
            if (!tile_ready) {
                mbar_wait_warp(&B.acc_empty[acc], ((t >> 1) & 1) ^ 1);
                mbar_wait_warp(&B.full[s], sph);
            }
            tc_fence_after();
            PW2_TRACE(t, 1);
            // per-tile bases produced by REDUX (destination = a uniform register, defined INSIDE the tile loop so the 48
            // operand addresses are not hoisted out of it as 48 long-lived values): every MMA operand below is
            // `uniform base + immediate`, one UIADD3 away from its UTCHMMA
            const uint32_t tmem_base = uniform_u32(a.tmem0 + (t & 0u));
            const uint32_t d = tmem_base + kD0 + acc * 64u;
            const uint32_t a2hi = tmem_base + a2off, a2lo = a2hi + (na2 == 2 ? 32u : 64u);      // lo, or (lo16 | hi16)
            const uint32_t x_lo = uniform_u32(desc_lo(sbase + s * a.stage_bytes)), by_lo = x_lo + xb16;
            const uint32_t lb = a.nlo == 2 ? (t & 1) : 0u;
            const uint32_t xl_lo = uniform_u32(desc_lo(lo_base0 + lb * a.lo_bytes)), byl_lo = xl_lo + xb16;
            // the next tile's stage / accumulator
            uint32_t s1 = s + 1, sph1 = sph;
            if (s1 == (uint32_t)a.stages) { s1 = 0; sph1 ^= 1u; }
            const bool more = t + 1 < total;
            // The hooks only TEST (non-blocking): a barrier that is not complete yet is awaited where it is needed, as
            // before -- blocking here would tie this tile's completion to the NEXT tile's data and cost a pipeline stage.
            // (Issuing the next tile's raw passes while the converters still work on this one was measured at config 5 and
            // dropped: 0.59 -> 0.62 ms forward; the accumulator's completion is then signalled behind 24 more queued MMAs.)
            bool next_ready = false, lo_ready = false;
            auto test_next = [&]() {
                next_ready = more && mbar_test_warp(&B.acc_empty[acc ^ 1u], (((t + 1) >> 1) & 1) ^ 1) && mbar_test_warp(&B.full[s1], sph1);
            };
            auto test_lo = [&]() { lo_ready = mbar_test_warp(&B.lo_full[lb], a.nlo == 2 ? ((t >> 1) & 1) : (t & 1)); };
            // raw passes: (A_hi [+ A_lo]) * B_raw
            if (SINGLE) pw2_issue_raw<K1S, NKS2, true>(d, tmem_base, a2hi, a2lo, x_lo, by_lo, slab16, idesc, a.K1slabs, nks2, a2_bar, upar, test_next);
            else if (CORR16) pw2_issue_raw<K1S, NKS2, true>(d, tmem_base, a2hi, a2lo, x_lo, by_lo, slab16, idesc, a.K1slabs, nks2, a2_bar, upar, test_lo);
            else pw2_issue_raw<K1S, NKS2, false>(d, tmem_base, a2hi, a2lo, x_lo, by_lo, slab16, idesc, a.K1slabs, nks2, a2_bar, upar, test_lo);
            umma_commit(&B.empty[s]);                              // the raw tile is free once the raw passes retire
            PW2_TRACE(t, 2);
            if (!SINGLE) {
                // remainder pass: A_hi * B_lo
                if (!lo_ready) mbar_wait_warp(&B.lo_full[lb], a.nlo == 2 ? ((t >> 1) & 1) : (t & 1));
                tc_fence_after();
                if (CORR16)
                    pw2_issue_lo16<(K1S >= 0 ? 2 * K1S : -1), (NKS2 >= 0 ? (NKS2 + 1) / 2 : -1)>(
                        d, tmem_base, a2lo, a2lo + (na2 == 2 ? 16u : 32u), xl_lo, xl_lo + byh_off, xl_lo + half16,
                        xl_lo + half16 + byh_off, slab16, idesc16, n16_1, n16_2, test_next);
                else
                    pw2_issue_lo<K1S, NKS2>(d, tmem_base, a2hi, xl_lo, byl_lo, slab16, idesc, a.K1slabs, nks2, test_next);
                umma_commit(&B.lo_empty[lb]);
            }
            umma_commit(&B.acc_full[acc]);
            PW2_TRACE(t, 4);
            if (has_z && tile == t1 - 1) umma_commit(&B.a2_empty[ub]);
            s = s1;
            sph = sph1;
            tile_ready = next_ready;
        }
    }
}

#define PW2_SHAPES(X) X(4, 8) X(2, 4) X(1, 3) X(4, 4) X(2, 8) X(1, 4) X(4, 0) X(2, 0) X(1, 0)


// Epilogue of one warp: lane = output channel, 8 positions per step.  The loop is deliberately NOT unrolled over the
// tile: a fully unrolled 32-position body (with GELU / GELU' inline) is ~50 KB of code that every epilogue warp
// streams through once per tile, and ncu showed those warps stalled on instruction fetch for 40 % of their samples
// while pacing the whole kernel.  The GELU' operand of step c+1 is loaded while step c is computed.
// HAS_MUL: 0 none, 1 multiply by GELU'(mul), 2 multiply by mul (a cached derivative); HAS_ACT: 0 none, 1 out_act = GELU(out),
// 2 out = GELU'(v) and out_act = GELU(v) from one rcp / ex2 pair (the next layer's backward then multiplies without any math),
// 3 trapezoid-weighted GELU sum over the row (F2V local branch: one partial per (row, segment, column half, channel), no
// spatial store), 4 its gradient out = gS[b][o] wx[x] wy[y] GELU'(v)
template <int HAS_MUL, int HAS_ACT>
__device__ __forceinline__ void pw2_epilogue(const PwArgs2& a, int warp, int lane, uint32_t tmem_base, int my_units,
                                             uint64_t* acc_full, uint64_t* acc_empty) {
    const int q4 = warp & 3, o = q4 * 32 + lane;
    const int half = (warp - 10) >> 2;
    const int hc = a.NT / 2;                                       // columns per half: 16, 24 or 32
    const int nsteps = hc >> 3;
    const int fr = o / a.COo, oo = o - fr * a.COo;              // lane = (row within the group, out channel)
    const float bias = (o < a.CO && a.bias) ? __ldg(a.bias + oo) : 0.f;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16) + kD0;
    const size_t cstride = (size_t)a.COo;
    uint32_t t = 0;
    for (int u = 0; u < my_units; ++u) {
        const int unit = (int)blockIdx.x + u * (int)gridDim.x;
        const int rg = unit / a.nseg, seg = unit - rg * a.nseg;
        const int t0 = seg * a.seg_tiles, t1 = min(a.tiles_per_row, t0 + a.seg_tiles);
        const long long row = (long long)rg * a.F + fr;
        const bool o_ok = o < a.CO && row < a.rows;
        float wsum_acc = 0.f, wrow = 0.f;                           // HAS_ACT 3 / 4: running sum, wx[x] (times gS[b][o])
        if (HAS_ACT >= 3 && o_ok) {
            const long long bb = row / a.P1w;
            wrow = __ldg(a.wx + (int)(row - bb * a.P1w));
            if (HAS_ACT == 4) wrow *= __ldg(a.gS + bb * a.COo + oo);
        }
        for (int tile = t0; tile < t1; ++tile, ++t) {
            const uint32_t acc = t & 1;
            const int yb = tile * a.NT + half * hc;                 // first position of this warp's columns
            const int nv = o_ok ? max(0, min(hc, a.S2 - yb)) : 0;   // valid positions of this lane
            const size_t base = ((size_t)row * a.S2 + yb) * cstride + oo;
            float mn[8];
            if (HAS_MUL) {
                // pull the NEXT tile's operand lines into L2 (lane j: the 128-byte channel slice of position j) ...
                int ntile = tile + 1, nrg = rg;
                if (ntile == t1) {                                  // first tile of this CTA's next unit
                    const int nunit = unit + (int)gridDim.x;
                    nrg = nunit / a.nseg;
                    ntile = nunit < a.units ? (nunit - nrg * a.nseg) * a.seg_tiles : -1;
                }
                const int frw = (q4 * 32) / a.COo, sb = q4 * 32 - frw * a.COo;      // warp-uniform (row in group, channel slice)
                const long long nrow = (long long)nrg * a.F + frw;
                const int ny = ntile * a.NT + half * hc + lane;
                if (ntile >= 0 && lane < hc && ny < a.S2 && nrow < a.rows && q4 * 32 < a.CO)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.mul + ((size_t)nrow * a.S2 + ny) * cstride + sb));
                // ... and this tile's first step into registers before the accumulator wait
                const float* mp = a.mul + base;
                if (__all_sync(0xffffffffu, 8 <= nv)) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) mn[j] = __ldg(mp + j * a.COo);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) mn[j] = j < nv ? __ldg(mp + j * a.COo) : 0.f;
                }
            }
            if (warp == 10) PW2_TRACE(t, 7);
            mbar_wait(&acc_full[acc], (t >> 1) & 1);
            tc_fence_after();
            if (warp == 10) PW2_TRACE(t, 8);
            const uint32_t taddr = lane_base + acc * 64u + (uint32_t)(half * hc);
            // element offsets of the 8 positions of a step (shared by the operand, the output and the GELU copy)
            uint32_t offs[8];                                       // in bytes
#pragma unroll
            for (int j = 0; j < 8; ++j) offs[j] = (uint32_t)(j * a.COo) * 4u;
            auto at = [](const float* p0, uint32_t byte_off) {
                return reinterpret_cast<const float*>(reinterpret_cast<const char*>(p0) + byte_off);
            };
            auto at_w = [](float* p0, uint32_t byte_off) { return reinterpret_cast<float*>(reinterpret_cast<char*>(p0) + byte_off); };
#pragma unroll 1
            for (int c = 0; c < nsteps; ++c) {
                uint32_t r[8];
#ifdef FNM_PW2_DEBUG_SWITCHES
                if (a.debug & 8) { for (int j = 0; j < 8; ++j) r[j] = 0; } else
#endif
                tmem_ld8(taddr + c * 8, r);
                // warp-uniform fast path (all 8 positions valid in every lane): no per-element predicates or branches,
                // so the eight GELU chains interleave
                const bool full = __all_sync(0xffffffffu, (c + 1) * 8 <= nv);
                float m[8];
                if (HAS_MUL) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) m[j] = mn[j];
                    if (c + 1 < nsteps) {
                        const float* mp = a.mul + base + (size_t)((c + 1) * 8) * cstride;
                        if (__all_sync(0xffffffffu, (c + 2) * 8 <= nv)) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) mn[j] = __ldg(at(mp, offs[j]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) mn[j] = (c + 1) * 8 + j < nv ? __ldg(at(mp, offs[j])) : 0.f;
                        }
                    }
                }
                tmem_ld_wait();
                float* op = a.out + base + (size_t)(c * 8) * cstride;
                float* ap = (HAS_ACT == 1 || HAS_ACT == 2) ? a.out_act + base + (size_t)(c * 8) * cstride : nullptr;
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[j] = __uint_as_float(r[j]) + bias;
                    if (HAS_MUL == 1) v[j] *= gelu_grad_fast(m[j]);
                    if (HAS_MUL == 2) v[j] *= m[j];
                }
#ifdef FNM_PW2_DEBUG_SWITCHES
                if ((a.debug & 1) && v[0] != 123.456f) continue;
#endif
                if (HAS_ACT >= 3) {
                    // F2V local branch: the weights of the 8 positions are warp-uniform loads
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int y = yb + c * 8 + j;
                        const float wyj = (c * 8 + j < nv) ? __ldg(a.wy + y) : 0.f;
                        if (HAS_ACT == 3) wsum_acc = fmaf(wyj, gelu_fast(v[j]), wsum_acc);
                        else v[j] = wrow * wyj * gelu_grad_fast(v[j]);
                    }
                    if (HAS_ACT == 3) continue;
                }
                float ga[8];
                if (HAS_ACT == 2) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { float g, dg; gelu_both_fast(v[j], g, dg); ga[j] = g; v[j] = dg; }
                } else if (HAS_ACT == 1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) ga[j] = gelu_fast(v[j]);
                }
                if (full) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) *at_w(op, offs[j]) = v[j];
                    if (HAS_ACT == 1 || HAS_ACT == 2) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) *at_w(ap, offs[j]) = ga[j];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (c * 8 + j < nv) {
                            *at_w(op, offs[j]) = v[j];
                            if (HAS_ACT == 1 || HAS_ACT == 2) *at_w(ap, offs[j]) = ga[j];
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
            if (warp == 10) PW2_TRACE(t, 9);
        }
        if (HAS_ACT == 3 && o_ok)                                   // one partial per (row, segment, column half, channel)
            a.out[(((size_t)row * a.nseg + seg) * 2 + half) * a.COo + oo] = wrow * wsum_acc;
    }
}

// CORR16: the opt-in bf16-correction variant as its own instantiation, so that the default kernel's code (register allocation,
// scheduling of the Z and converter warps) does not depend on it: with run-time `corr16` branches in one kernel the default
// path lost 25 % at the 32-channel shapes.
template <bool CORR16>
__global__ void __launch_bounds__(kPwThreads, 1)
pw2_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapBy,
           const __grid_constant__ CUtensorMap mapZ, const __grid_constant__ CUtensorMap mapBy16, const PwArgs2 a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    // [stages x {X raw slabs, by raw slabs}] [lo copy of one stage] [Z staging] [barriers]
    uint8_t* lo_buf = smem + (size_t)a.stages * a.stage_bytes;
    float* zs = reinterpret_cast<float*>(lo_buf + (size_t)a.nlo * a.lo_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(zs) + (size_t)a.nz * a.z_bytes);
    uint64_t* full = bars;                          // [kPwMaxStages]
    uint64_t* empty = bars + kPwMaxStages;          // [kPwMaxStages]
    uint64_t* lo_full = bars + 2 * kPwMaxStages;    // [2]
    uint64_t* lo_empty = lo_full + 2;               // [2]
    uint64_t* z_full = lo_full + 4;                 // [4]
    uint64_t* z_empty = lo_full + 8;                // [4]
    uint64_t* a2_full = lo_full + 12;               // [2]
    uint64_t* a2_empty = lo_full + 14;              // [2]
    uint64_t* acc_full = lo_full + 16;              // [2]
    uint64_t* acc_empty = lo_full + 18;             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lo_full + 20);

    const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
    const bool has_x = a.CI > 0, has_z = a.LY > 0;
    // Z^T hi/lo (A2): one 128-column buffer for up to 64 mode columns, or -- when LYp <= 32 -- two 64-column
    // buffers so that the next grid row's Z is converted while the current row is still being multiplied
    const uint32_t na2 = a.LYp <= 32 ? 2u : 1u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], a.single_pass ? 1 : 5); }
        // remainder tile complete: the 4 converter warps (+ the producer's expect_tx when the table's bf16 pair is TMA-loaded)
        for (int s = 0; s < 2; ++s) { mbar_init(&lo_full[s], (CORR16 && a.by16) ? 5 : 4); mbar_init(&lo_empty[s], 1); }
        for (int s = 0; s < 4; ++s) { mbar_init(&z_full[s], 1); mbar_init(&z_empty[s], 4); }
        for (int s = 0; s < 2; ++s) { mbar_init(&a2_full[s], 4); mbar_init(&a2_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
        fence_barrier_init();
        tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapBy); tma_prefetch_desc(&mapZ); tma_prefetch_desc(&mapBy16);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    // ---- A1: constant pointwise weights, Wk^T[o][k], hi/lo -> TMEM (epilogue warps 10-13, lane = o)
    if (has_x && warp >= 10 && warp < 14) {
        const int q4 = warp & 3, o = q4 * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
        for (int k0 = 0; k0 < a.K1slabs * 32; k0 += 16) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int k = k0 + j;
                v[j] = 0.f;
                // folded problem: block-diagonal weights, lane = (sub-position, out channel), k = (sub-position, in channel)
                const int ro = o / a.COo, oo = o - ro * a.COo, rk = k / a.CIo, ii = k - rk * a.CIo;
                if (o < a.CO && k < a.CI && ro == rk)
                    v[j] = a.wc_is_kn ? __ldg(a.wc + (size_t)ii * a.COo + oo) : __ldg(a.wc + (size_t)oo * a.CIo + ii);
            }
            if (CORR16) split_to_tmem16_corr16(lane_base + kA1Hi + k0, lane_base + kA1Lo16 + k0 / 2, lane_base + kA1Hi16 + k0 / 2, v);
            else split_to_tmem16(lane_base + kA1Hi + k0, lane_base + kA1Lo + k0, v, !a.single_pass);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const int my_units = (a.units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto unit_tiles = [&](int u, int& row, int& t0, int& t1) {
        const int unit = (int)blockIdx.x + u * (int)gridDim.x;
        row = unit / a.nseg;
        const int seg = unit - row * a.nseg;
        t0 = seg * a.seg_tiles;
        t1 = min(a.tiles_per_row, t0 + a.seg_tiles);
    };

    if (warp == 0) {
        // =========================================================================== TMA producer
        if (lane == 0) {
            uint32_t t = 0;
            for (int u = 0; u < my_units; ++u) {
                int row, t0, t1;
                unit_tiles(u, row, t0, t1);
                if (has_z) {                                       // ring of nz staging buffers for the rows' Z tiles
                    const uint32_t zb = (uint32_t)u % (uint32_t)a.nz, zpar = ((uint32_t)u / (uint32_t)a.nz) & 1;
                    mbar_wait(&z_empty[zb], zpar ^ 1);
                    mbar_expect_tx(&z_full[zb], a.z_bytes);
                    tma_load_2d(&mapZ, &z_full[zb], reinterpret_cast<uint8_t*>(zs) + (size_t)zb * a.z_bytes, 0, row * a.F * a.LY);
                }
                for (int tile = t0; tile < t1; ++tile, ++t) {
                    const uint32_t s = t % (uint32_t)a.stages;
                    mbar_wait(&empty[s], ((t / (uint32_t)a.stages) & 1) ^ 1);
                    PW2_TRACE(t, 10);
#ifdef FNM_PW2_DEBUG_SWITCHES
                    if ((a.debug & 4) && t >= (uint32_t)a.stages) { mbar_arrive(&full[s]); continue; }
#endif
                    mbar_expect_tx(&full[s], a.stage_bytes);
                    uint8_t* st = smem + (size_t)s * a.stage_bytes;
                    const int y0 = tile * a.NT;
                    // K-slab kb of the X tile: 32 channels of grid row (row group * F + kb / slabs per row)
                    const int spr = a.F > 1 ? a.CIo / 32 : a.K1slabs;
                    for (int kb = 0; kb < a.K1slabs; ++kb)
                        tma_load_3d(&mapX, &full[s], st + (size_t)kb * a.NT * 128, (kb % spr) * 32, y0, row * a.F + kb / spr);
                    for (int kb = 0; kb < a.K2slabs; ++kb)
                        tma_load_2d(&mapBy, &full[s], st + a.x_bytes + (size_t)kb * a.NT * 128, kb * 32, y0);
                }
            }
        } else if (CORR16 && lane == 1 && a.by16) {
            // second producer lane: the bf16(by) | bf16(by_lo) tiles of every tile column, from the precomputed table pair
            // straight into the remainder buffers (paced by lo_empty, independently of the raw stages lane 0 runs ahead with)
            const uint32_t half = a.lo_bytes / 2, slab_b = (uint32_t)a.NT * 128u;
            uint32_t t = 0;
            for (int u = 0; u < my_units; ++u) {
                int row, t0, t1;
                unit_tiles(u, row, t0, t1);
                for (int tile = t0; tile < t1; ++tile, ++t) {
                    const uint32_t lb = a.nlo == 2 ? (t & 1) : 0u;
                    mbar_wait(&lo_empty[lb], (a.nlo == 2 ? ((t >> 1) & 1) : (t & 1)) ^ 1);
                    uint8_t* lob = lo_buf + (size_t)lb * a.lo_bytes + (size_t)a.k1s16 * slab_b;
                    mbar_expect_tx(&lo_full[lb], 2u * (uint32_t)a.k2s16 * slab_b);
                    for (int kb = 0; kb < a.k2s16; ++kb) {
                        tma_load_3d(&mapBy16, &lo_full[lb], lob + (size_t)kb * slab_b, kb * 64, tile * a.NT, 0);
                        tma_load_3d(&mapBy16, &lo_full[lb], lob + half + (size_t)kb * slab_b, kb * 64, tile * a.NT, 1);
                    }
                }
            }
        }
    } else if (warp == kPwMmaWarp) {
        // =========================================================================== MMA issuer
        // One CTA per SM owns all 512 TMEM columns, so the allocation starts at lane 0 / column 0; the issue loop takes the
        // base from a kernel PARAMETER (always 0: a uniform-register source) so that every TMEM operand address stays in
        // the uniform datapath.
        if (tmem_base != 0u) __trap();
        const uint32_t sbase = smem_u32(smem);
        const uint32_t lo_base0 = smem_u32(lo_buf);
        // One MMA shape for every tile.  Issuing the half-empty last tile of a row with a narrower N (idesc N = 32) was
        // measured 10-14 % SLOWER overall: alternating instruction shapes costs more than the columns it saves.
        PwBars B{full, empty, lo_full, lo_empty, a2_full, a2_empty, acc_full, acc_empty};
        const int nks2 = a.LYp / 8;
        bool done = false;
#define PW2_LOOP(K1, K2)                                                                                                 \
        if (!done && a.K1slabs == K1 && nks2 == K2) {                                                                    \
            if (CORR16) pw2_mma_loop<K1, K2, 2>(a, B, sbase, lo_base0, my_units, na2, has_z);                            \
            else if (a.single_pass) pw2_mma_loop<K1, K2, 1>(a, B, sbase, lo_base0, my_units, na2, has_z);                \
            else pw2_mma_loop<K1, K2, 0>(a, B, sbase, lo_base0, my_units, na2, has_z);                                   \
            done = true;                                                                                                 \
        }
        PW2_SHAPES(PW2_LOOP)
#undef PW2_LOOP
        if (!done) {
            if (CORR16) pw2_mma_loop<-1, -1, 2>(a, B, sbase, lo_base0, my_units, na2, has_z);
            else if (a.single_pass) pw2_mma_loop<-1, -1, 1>(a, B, sbase, lo_base0, my_units, na2, has_z);
            else pw2_mma_loop<-1, -1, 0>(a, B, sbase, lo_base0, my_units, na2, has_z);
        }
    } else if (warp >= 2 && warp < 6) {
        // =========================================================================== Z warpgroup: Z[row] (smem) -> A2 hi/lo (TMEM)
        const int q4 = warp & 3, c = q4 * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
        const bool c_ok = c < a.CO;
        const int zr = c / a.COo, zo = c - zr * a.COo;                  // lane = (row within the group, channel)
        for (int u = 0; u < (has_z ? my_units : 0); ++u) {
            const uint32_t ub = (uint32_t)u % na2, upar = ((uint32_t)u / na2) & 1;
            const uint32_t a2w = na2 == 2 ? 64u : 128u;
            const uint32_t a2hi = kA2Hi + (na2 == 2 ? 64u * ub : 0u), a2lo = a2hi + a2w / 2;
            const uint32_t zb = (uint32_t)u % (uint32_t)a.nz, zpar = ((uint32_t)u / (uint32_t)a.nz) & 1;
            const float* zsb = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(zs) + (size_t)zb * a.z_bytes);
            mbar_wait(&z_full[zb], zpar);
            mbar_wait(&a2_empty[ub], upar ^ 1);
            tc_fence_after();
            for (int l0 = 0; l0 < a.LYp; l0 += 16) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int kk = l0 + j;                                        // each row of the group brings its own Z tile
                    v[j] = (c_ok && kk < a.LY) ? zsb[(zr * a.LY + kk) * a.COo + zo] : 0.f;
                }
                if (CORR16) split_to_tmem16_corr16(lane_base + a2hi + l0, lane_base + a2lo + l0 / 2, lane_base + a2lo + a2w / 4 + l0 / 2, v);
                else split_to_tmem16(lane_base + a2hi + l0, lane_base + a2lo + l0, v, !a.single_pass);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&z_empty[zb]); mbar_arrive(&a2_full[ub]); }
        }
    } else if (warp >= 6 && warp < 10) {
        // =========================================================================== converters: lo = raw - trunc(raw)
        if (!a.single_pass) {
            const int ct = threadIdx.x - 6 * 32;                       // 0..127
            uint32_t t = 0;
            for (int u = 0; u < my_units; ++u) {
                int row, t0, t1;
                unit_tiles(u, row, t0, t1);
                for (int tile = t0; tile < t1; ++tile, ++t) {
                    const uint32_t s = t % (uint32_t)a.stages;
                    mbar_wait(&full[s], (t / (uint32_t)a.stages) & 1);
                    const uint32_t lb = a.nlo == 2 ? (t & 1) : 0u;
                    mbar_wait(&lo_empty[lb], (a.nlo == 2 ? ((t >> 1) & 1) : (t & 1)) ^ 1);
                    if (warp == 6) PW2_TRACE(t, 5);
                    if (CORR16) {
                        // two bf16 tiles: bf16(B) and bf16(B - trunc(B)).  Warp w converts the `up = w & 1` half of every
                        // 64-channel output slab, i.e. fp32 slab 2 so + up whole: lane = (row, q) reads the row's chunks 2q, 2q + 1
                        // (8 channels) and writes output chunk 4 up + q.  Rows of a quarter-warp differ by XOR 5 in their low bits,
                        // which keeps the swizzled 16-byte accesses of its 8 lanes on 8 different bank groups for the loads (bit 0
                        // flips even <-> odd chunks) and for the stores (bit 2 flips the half of the row).
                        const uint32_t srcb = smem_u32(smem) + s * a.stage_bytes, dstb = smem_u32(lo_buf) + lb * a.lo_bytes;
                        const uint32_t half = a.lo_bytes / 2, slab_b = (uint32_t)a.NT * 128u;
                        const int up = (warp - 6) & 1, q = lane & 3, rsel = lane >> 2;
                        const int rr = (rsel >> 1) ^ ((rsel & 1) ? 5 : 0);
                        const int rbase = 8 * ((warp - 6) >> 1) + rr;                 // rows rbase + 16 i
                        const uint32_t in0 = (uint32_t)(((2 * q) ^ rr) << 4), in1 = (uint32_t)(((2 * q + 1) ^ rr) << 4);
                        const uint32_t outc = (uint32_t)(((4 * up + q) ^ rr) << 4);
                        const int nparts = a.by16 ? 1 : 2;
#pragma unroll 1
                        for (int part = 0; part < nparts; ++part) {
                            const int nin = part ? a.K2slabs : a.K1slabs, nout = part ? a.k2s16 : a.k1s16;
                            const uint32_t src0 = srcb + (part ? a.x_bytes : 0u), dst0 = dstb + (part ? (uint32_t)a.k1s16 * slab_b : 0u);
#pragma unroll 1
                            for (int so = 0; so < nout; ++so) {
                                if (2 * so + up >= nin) break;                        // odd slab count: that half is never multiplied
                                const uint32_t sin = src0 + (uint32_t)(2 * so + up) * slab_b, sout = dst0 + (uint32_t)so * slab_b;
                                // all (up to four) rows of this thread first: 8 loads in flight, then convert and store
                                float4 e[4], o[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int r = rbase + 16 * i;
                                    e[i] = make_float4(0.f, 0.f, 0.f, 0.f); o[i] = e[i];
                                    if (r < a.NT) {
                                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(e[i].x), "=f"(e[i].y), "=f"(e[i].z), "=f"(e[i].w) : "r"(sin + (uint32_t)r * 128u + in0));
                                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o[i].x), "=f"(o[i].y), "=f"(o[i].z), "=f"(o[i].w) : "r"(sin + (uint32_t)r * 128u + in1));
                                    }
                                }
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int r = rbase + 16 * i;
                                    if (r >= a.NT) break;
                                    auto lo_of = [](float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); };
                                    const uint32_t h0 = pack_bf16x2(e[i].x, e[i].y), h1 = pack_bf16x2(e[i].z, e[i].w);
                                    const uint32_t h2 = pack_bf16x2(o[i].x, o[i].y), h3 = pack_bf16x2(o[i].z, o[i].w);
                                    const uint32_t l0 = pack_bf16x2(lo_of(e[i].x), lo_of(e[i].y)), l1 = pack_bf16x2(lo_of(e[i].z), lo_of(e[i].w));
                                    const uint32_t l2 = pack_bf16x2(lo_of(o[i].x), lo_of(o[i].y)), l3 = pack_bf16x2(lo_of(o[i].z), lo_of(o[i].w));
                                    const uint32_t dsta = sout + (uint32_t)r * 128u + outc;
                                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dsta), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
                                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dsta + half), "r"(l0), "r"(l1), "r"(l2), "r"(l3) : "memory");
                                }
                            }
                        }
                    } else {
                        const uint32_t src = smem_u32(smem) + s * a.stage_bytes + (uint32_t)ct * 16u;
                        const uint32_t dst = smem_u32(lo_buf) + lb * a.lo_bytes + (uint32_t)ct * 16u;
                        // linear copy (the remainder tile has the raw tile's swizzled layout, so offsets are identical):
                        // stage_bytes is a multiple of 4 KB; 2 KB per pass of the 128 threads, two passes per iteration
                        uint32_t off = 0;
#ifdef FNM_PW2_DEBUG_SWITCHES
                        if (a.debug & 2) off = a.stage_bytes;
#endif
                        for (; off + 4096u <= a.stage_bytes; off += 4096u) {
                            float4 v0, v1;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w) : "r"(src + off));
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w) : "r"(src + off + 2048u));
                            v0.x -= __uint_as_float(__float_as_uint(v0.x) & 0xffffe000u); v0.y -= __uint_as_float(__float_as_uint(v0.y) & 0xffffe000u);
                            v0.z -= __uint_as_float(__float_as_uint(v0.z) & 0xffffe000u); v0.w -= __uint_as_float(__float_as_uint(v0.w) & 0xffffe000u);
                            v1.x -= __uint_as_float(__float_as_uint(v1.x) & 0xffffe000u); v1.y -= __uint_as_float(__float_as_uint(v1.y) & 0xffffe000u);
                            v1.z -= __uint_as_float(__float_as_uint(v1.z) & 0xffffe000u); v1.w -= __uint_as_float(__float_as_uint(v1.w) & 0xffffe000u);
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + off), "f"(v0.x), "f"(v0.y), "f"(v0.z), "f"(v0.w) : "memory");
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + off + 2048u), "f"(v1.x), "f"(v1.y), "f"(v1.z), "f"(v1.w) : "memory");
                        }
                        if (off < a.stage_bytes) {                        // odd number of 2 KB passes
                            float4 v0;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w) : "r"(src + off));
                            v0.x -= __uint_as_float(__float_as_uint(v0.x) & 0xffffe000u); v0.y -= __uint_as_float(__float_as_uint(v0.y) & 0xffffe000u);
                            v0.z -= __uint_as_float(__float_as_uint(v0.z) & 0xffffe000u); v0.w -= __uint_as_float(__float_as_uint(v0.w) & 0xffffe000u);
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + off), "f"(v0.x), "f"(v0.y), "f"(v0.z), "f"(v0.w) : "memory");
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(&lo_full[lb]); mbar_arrive(&empty[s]); }
                    if (warp == 6) PW2_TRACE(t, 6);
                }
            }
        }
    } else if (warp >= 10 && warp < 18) {
        // =========================================================================== epilogue (8 warps)
        // explicitly specialised on (GELU' operand, GELU copy): with the flags tested per element the loop is
        // 2x slower, and whether ptxas unswitches it on its own changes with unrelated edits (measured)
        const int mm = a.mul ? (a.mul_raw ? 2 : 1) : 0, am = a.wsum ? 2 + a.wsum : (a.out_act ? (a.out_grad ? 2 : 1) : 0);
        if (am == 3) pw2_epilogue<0, 3>(a, warp, lane, tmem_base, my_units, acc_full, acc_empty);
        else if (am == 4) pw2_epilogue<0, 4>(a, warp, lane, tmem_base, my_units, acc_full, acc_empty);
        else if (mm == 0 && am == 0) pw2_epilogue<0, 0>(a, warp, lane, tmem_base, my_units, acc_full, acc_empty);
        else if (mm == 0 && am == 1) pw2_epilogue<0, 1>(a, warp, lane, tmem_base, my_units, acc_full, acc_empty);
        else if (mm == 0 && am == 2) pw2_epilogue<0, 2>(a, warp, lane, tmem_base, my_units, acc_full, acc_empty);
        else if (mm == 1 && am == 0) pw2_epilogue<1, 0>(a, warp, lane, tmem_base, my_units, acc_full, acc_empty);
        else if (mm == 2 && am == 0) pw2_epilogue<2, 0>(a, warp, lane, tmem_base, my_units, acc_full, acc_empty);
        else if (mm == 1) pw2_epilogue<1, 1>(a, warp, lane, tmem_base, my_units, acc_full, acc_empty);       // (mul, GELU copy): not used by the layers
        else pw2_epilogue<2, 1>(a, warp, lane, tmem_base, my_units, acc_full, acc_empty);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

bool tc_make_map_swz(CUtensorMap* m, const float* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                     const cuuint32_t* box, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(ptr), dims, strides_bytes,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool tc_make_map(CUtensorMap* m, const float* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                 const cuuint32_t* box, bool swizzle128) {
    return tc_make_map_swz(m, ptr, rank, dims, strides_bytes, box, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
}

// Row folding for narrow layers: F consecutive grid rows share one CTA pass.  TMEM lane = (row in group, out
// channel), the K axis of the pointwise term is (row in group, in channel) with block-diagonal weights -- K-slab
// kb of the X tile is simply TMA-loaded from grid row kb / (C/32) -- and every row of the group contributes its own
// Z^T rows while the twiddle tile `by` is shared.  All 128 lanes, full MMA K and F times fewer tiles for C = 32 / 64,
// for any row length and mode count.
static int pw2_fold(int64_t rows, int CI, int CO) {
    if ((CI != 0 && CI != CO) || CO % 32 != 0) return 1;
    for (int F = 4; F >= 2; F >>= 1)
        if (CO * F <= 128 && rows >= 2 * F) return F;
    return 1;
}

static bool pw2_corr16_default() {
    const char* e = getenv("FNM_PW2_CORR");
    return e && e[0] == 'b';                                     // FNM_PW2_CORR=bf16: bf16 correction passes (see the header note)
}

static bool pw2_configure(PwArgs2& a, int64_t rows, int S2, int LY, int CI, int CO, int by_ld = 0, int mode = FNM_MODE_FP32) {
    a.corr16 = (mode != FNM_MODE_TF32 && pw2_corr16_default()) ? 1 : 0;
    // what TMA can address (16-byte strides: the twiddle table's leading dimension by_ld, LY unless the caller padded it)
    // and what the TMEM map holds (128 lanes, 64 mode columns)
    if (by_ld == 0) by_ld = LY;
    if (rows < 1 || rows > (1LL << 28) || S2 < 1 || LY < 0 || LY > 64 || by_ld % 4 != 0 || by_ld < LY) return false;
    if (LY == 0 && CI == 0) return false;
    if (CO < 4 || CO > 128 || CO % 4 != 0) return false;
    if (CI != 0 && (CI < 4 || CI > 128 || CI % 4 != 0)) return false;
    a.CIo = CI > 0 ? CI : 1; a.COo = CO;
    {
        const char* e = getenv("FNM_PW2_NOFOLD");
        a.F = (e && e[0] == '1') ? 1 : pw2_fold(rows, CI, CO);
        CI *= a.F; CO *= a.F;
    }
    if (a.F > 1 && (uint64_t)a.F * LY * a.COo * 4u > 32768u) return false;
    a.rows = rows; a.S2 = S2; a.LY = LY; a.CI = CI; a.CO = CO;
    a.LYp = (LY + 7) / 8 * 8;
    a.K1slabs = (CI + 31) / 32;
    a.K2slabs = (a.LYp + 31) / 32;
    a.z_bytes = (uint32_t)a.F * LY * a.COo * 4u;
    int best = 0, best_cost = 1 << 30;
    for (int nt = 64; nt >= 32; nt -= 16) {                    // padded positions + a per-tile overhead worth ~32 positions
        const int tiles = (S2 + nt - 1) / nt;                  // (measured at config 5: 5 x 64 beats 6 x 48); ties -> wider tile
        const int cost = tiles * (nt + 32);
        if (cost < best_cost) { best_cost = cost; best = nt; }
    }
    a.NT = best;
    { const char* e = getenv("FNM_PW2_NT"); if (e && (atoi(e) == 32 || atoi(e) == 48 || atoi(e) == 64)) a.NT = atoi(e); }
    a.tiles_per_row = (S2 + a.NT - 1) / a.NT;
    a.x_bytes = (uint32_t)a.K1slabs * a.NT * 128u;
    a.stage_bytes = a.x_bytes + (uint32_t)a.K2slabs * a.NT * 128u;
    // Z staging: a ring when the tiles are small (few tiles per row: the row-to-row hand-over must pipeline)
    a.nz = 1;
    if (a.z_bytes > 0 && a.z_bytes % 128 == 0) { a.nz = (int)(32768u / a.z_bytes); a.nz = a.nz < 1 ? 1 : (a.nz > 4 ? 4 : a.nz); }
    const uint32_t fixed = (uint32_t)a.nz * a.z_bytes + 1024 /*align*/ + 512 /*barriers*/;
    // remainder buffer: the fp32 lo tile (stage layout), or the two bf16 tiles (64-channel slabs; an odd slab count rounds up)
    a.k1s16 = (a.K1slabs + 1) / 2; a.k2s16 = (a.K2slabs + 1) / 2;
    a.lo_bytes = a.corr16 ? 2u * (uint32_t)(a.k1s16 + a.k2s16) * a.NT * 128u : a.stage_bytes;
    if (fixed + 2 * a.stage_bytes + a.lo_bytes > kPwSmemBudget) return false;        // 2 raw stages + 1 remainder buffer at least
    // bytes in flight are what hides HBM latency (~90 KB per SM): everything beyond the remainder buffers is raw stages
    a.nlo = fixed + 2 * a.stage_bytes + 2 * a.lo_bytes <= kPwSmemBudget ? 2 : 1;
    { const char* e = getenv("FNM_PW2_NLO"); if (e && a.nlo == 2) a.nlo = atoi(e) == 2 ? 2 : 1; }
    const int slots = (int)((kPwSmemBudget - fixed - a.nlo * a.lo_bytes) / a.stage_bytes);
    a.stages = slots > kPwMaxStages ? kPwMaxStages : slots;
    { const char* e = getenv("FNM_PW2_STAGES"); if (e && atoi(e) >= 2 && atoi(e) <= a.stages) a.stages = atoi(e); }
    // units: whole rows when there are enough of them, otherwise row segments (1-D problems have few rows)
    int nseg = 1;
    const int want = 2 * sm_count();
    const int64_t grows = (rows + a.F - 1) / a.F;               // row groups
    if (grows < want) nseg = (int)((want + grows - 1) / grows);
    if (nseg > a.tiles_per_row) nseg = a.tiles_per_row;
    a.seg_tiles = (a.tiles_per_row + nseg - 1) / nseg;
    a.nseg = (a.tiles_per_row + a.seg_tiles - 1) / a.seg_tiles;
    const long long units = grows * a.nseg;
    if (units > (1LL << 30)) return false;
    a.units = (int)units;
    return true;
}

}  // namespace tc

using namespace tc;

bool tc_pw2_supported(int64_t rows, int S2, int LY, int K, int N, int act, const void* x, const void* Z, const void* by, int by_ld) {
    if (!tc_enabled() || act != 0 || !encode_fn()) return false;
    { const char* e = getenv("FNM_DISABLE_PW2"); if (e && e[0] == '1') return false; }
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(by)) & 15) return false;
    PwArgs2 a{};
    return pw2_configure(a, rows, S2, LY, K, N, by_ld);
}

int tc_pw2(const float* x, const float* wc, int wc_is_kn, const float* bias, const float* Z, const float* by,
           const float* mul, float* out, float* out_act, int64_t rows, int S2, int LY, int K, int N, int mode,
           cudaStream_t st, int by_ld, int flags, const void* by16) {
    PwArgs2 a{};
    const int CI = (x && wc) ? K : 0;
    if (by_ld == 0) by_ld = LY;
    if (!pw2_configure(a, rows, S2, LY, CI, N, by_ld, mode)) return fail(FNM_ENOTSUP, "pointwise_ysyn: shape not served by the TMA kernel");
    a.wc = wc; a.bias = bias; a.mul = mul; a.out = out; a.out_act = out_act; a.wc_is_kn = wc_is_kn;
    a.single_pass = mode == FNM_MODE_TF32;
    a.mul_raw = (flags & FNM_PW_MUL_IS_GRAD) ? 1 : 0;
    a.out_grad = (flags & FNM_PW_OUT_IS_GRAD) ? 1 : 0;
    { const char* e = getenv("FNM_PW2_DEBUG"); a.debug = e ? atoi(e) : 0; }
    CUtensorMap mx, mb, mz, mb16;
    a.by16 = 0;
    if (LY > 0 && a.corr16 && by16 && by_ld == LY && (reinterpret_cast<uintptr_t>(by16) & 15) == 0) {
        // [2][S2][LY8] bfloat16: boxes of 64 columns x NT rows of one of the two tables, columns past LY8 read as zero
        const cuuint64_t ly8 = (cuuint64_t)((LY + 7) / 8 * 8);
        const cuuint64_t dims[3] = {ly8, (cuuint64_t)a.S2, 2};
        const cuuint64_t str[2] = {ly8 * 2, ly8 * 2 * (cuuint64_t)a.S2};
        const cuuint32_t box[3] = {64, (cuuint32_t)a.NT, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        EncodeTiledFn fn = encode_fn();
        if (fn && fn(&mb16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(by16), dims, str, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
            a.by16 = 1;
    }
    if (LY > 0) {
        const cuuint64_t dims[2] = {(cuuint64_t)by_ld, (cuuint64_t)a.S2};        // columns LY .. by_ld - 1 of a padded table are zero
        const cuuint64_t str[1] = {(cuuint64_t)by_ld * 4};
        const cuuint32_t box[2] = {32, (cuuint32_t)a.NT};
        if (!tc_make_map(&mb, by, 2, dims, str, box, true)) return fail(FNM_ECUDA, "pointwise_ysyn: tensor map (by) failed");
    }
    if (LY > 0) {
        const cuuint64_t dims[2] = {(cuuint64_t)a.COo, (cuuint64_t)rows * a.LY};   // Z[row][l][o]; one box = the F rows of a group
        const cuuint64_t str[1] = {(cuuint64_t)a.COo * 4};
        const cuuint32_t box[2] = {(cuuint32_t)a.COo, (cuuint32_t)(a.F * a.LY)};
        if (!tc_make_map(&mz, Z, 2, dims, str, box, false)) return fail(FNM_ECUDA, "pointwise_ysyn: tensor map (Z) failed");
    }
    if (CI > 0) {
        const cuuint64_t dims[3] = {(cuuint64_t)a.CIo, (cuuint64_t)a.S2, (cuuint64_t)rows};
        const cuuint64_t str[2] = {(cuuint64_t)a.CIo * 4, (cuuint64_t)a.S2 * a.CIo * 4};
        const cuuint32_t box[3] = {32, (cuuint32_t)a.NT, 1};
        if (!tc_make_map(&mx, x, 3, dims, str, box, true)) return fail(FNM_ECUDA, "pointwise_ysyn: tensor map (x) failed");
    } else {
        mx = mb;
    }
    if (LY == 0) { mb = mx; mz = mx; }                          // never dereferenced without a spectral term
    if (!a.by16) mb16 = mx;
    const size_t smem = (size_t)a.stages * a.stage_bytes + (size_t)a.nlo * a.lo_bytes + (size_t)a.nz * a.z_bytes + 1024 + 512;
    static std::atomic<uint64_t> attr_set{0};
    int attr_set_dev = 0;
    if (once_needed(attr_set, attr_set_dev)) {
        cudaError_t e = cudaFuncSetAttribute(pw2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(pw2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return fail(FNM_ECUDA, "pointwise_ysyn: smem attribute: %s", cudaGetErrorString(e));
        once_done(attr_set, attr_set_dev);
    }
    const int grid = a.units < sm_count() ? a.units : sm_count();
    {
        LaunchScope ls(LY > 0 ? "pointwise_ysyn" : "pointwise_conv", st);
        if (a.corr16) pw2_kernel<true><<<grid, kPwThreads, smem < 120 * 1024 ? 120 * 1024 : smem, st>>>(mx, mb, mz, mb16, a);
        else pw2_kernel<false><<<grid, kPwThreads, smem < 120 * 1024 ? 120 * 1024 : smem, st>>>(mx, mb, mz, mb16, a);
    }
    return check_launch("pointwise_ysyn");
}

// F2V local branch on the fused pointwise kernel (no spectral term): the hidden tensor h1 = fc1(x) never reaches HBM.
//   forward : partial[row][seg][half][h] = wx[x] sum_{y in segment half} wy[y] GELU(h1[row][y][h])
//   backward: g_h1[row][y][h] = gS[b][h] wx[x] wy[y] GELU'(h1[row][y][h])      (h1 recomputed from x)
bool tc_local_functional_supported(int64_t rows, int S2, int K, int N, const void* x) {
    return tc_pw2_supported(rows, S2, 0, K, N, 0, x, nullptr, nullptr);
}

int tc_local_functional_parts(int64_t rows, int S2, int K, int N) {      // partials per (row, channel): 2 * segments
    PwArgs2 a{};
    if (!pw2_configure(a, rows, S2, 0, K, N)) return 0;
    return 2 * a.nseg;
}

int tc_local_functional(const float* x, const float* w1, const float* b1, const float* wx, const float* wy, const float* gS,
                        float* out, int64_t rows, int P1, int S2, int K, int N, int mode, cudaStream_t st) {
    PwArgs2 a{};
    if (!pw2_configure(a, rows, S2, 0, K, N, 0, mode)) return fail(FNM_ENOTSUP, "local_functional: shape not served by the TMA kernel");
    a.wc = w1; a.bias = b1; a.mul = nullptr; a.out = out; a.out_act = nullptr; a.wc_is_kn = 0;
    a.single_pass = mode == FNM_MODE_TF32;
    a.wsum = gS ? 2 : 1; a.P1w = P1; a.wx = wx; a.wy = wy; a.gS = gS;
    CUtensorMap mx;
    const cuuint64_t dims[3] = {(cuuint64_t)a.CIo, (cuuint64_t)a.S2, (cuuint64_t)rows};
    const cuuint64_t str[2] = {(cuuint64_t)a.CIo * 4, (cuuint64_t)a.S2 * a.CIo * 4};
    const cuuint32_t box[3] = {32, (cuuint32_t)a.NT, 1};
    if (!tc_make_map(&mx, x, 3, dims, str, box, true)) return fail(FNM_ECUDA, "local_functional: tensor map (x) failed");
    const size_t smem = (size_t)a.stages * a.stage_bytes + (size_t)a.nlo * a.lo_bytes + (size_t)a.nz * a.z_bytes + 1024 + 512;
    static std::atomic<uint64_t> attr_set{0};
    int attr_set_dev = 0;
    if (once_needed(attr_set, attr_set_dev)) {
        cudaError_t e = cudaFuncSetAttribute(pw2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(pw2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return fail(FNM_ECUDA, "local_functional: smem attribute: %s", cudaGetErrorString(e));
        once_done(attr_set, attr_set_dev);
    }
    const int grid = a.units < sm_count() ? a.units : sm_count();
    const char* what = gS ? "local_functional_bwd" : "local_functional_fwd";
    {
        LaunchScope ls(what, st);
        if (a.corr16) pw2_kernel<true><<<grid, kPwThreads, smem < 120 * 1024 ? 120 * 1024 : smem, st>>>(mx, mx, mx, mx, a);
        else pw2_kernel<false><<<grid, kPwThreads, smem < 120 * 1024 ? 120 * 1024 : smem, st>>>(mx, mx, mx, mx, a);
    }
    return check_launch(what);
}

}  // namespace fnm

#ifdef FNM_PW2_TRACE
extern "C" int fnm_debug_pw2_trace(long long* host, int n) {
    const int total = fnm::tc::kTraceTiles * fnm::tc::kTraceEv;
    if (n > total) n = total;
    return (int)cudaMemcpyFromSymbol(host, fnm::tc::g_pw2_trace, sizeof(long long) * n);
}
#endif
