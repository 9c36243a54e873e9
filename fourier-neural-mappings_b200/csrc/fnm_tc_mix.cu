// tc_mix -- the per-mode complex channel mixing (compl_mul, models/shared.py:48-53) and its two
// backward products on tcgen05 (3xTF32), plus the layout permutation that makes the weights mode-major.
//
// The reference keeps spectral weights as (Cin, Cout, kx, ky) complex64: the modes are the FASTEST
// axis, so the (Cin x Cout) matrix of one mode is scattered with an 8 KB stride.  Each layer call first
// permutes the weights into mode-major scratch copies (one streaming pass, ~0.1 ms for 268 MB):
//     W1[m][i][o]   (forward:  lanes = o)          W2[m][o][i]   (backward-x: lanes = i)
// and the weight gradient is produced mode-major and permuted back into the reference layout.
//
// One kernel serves the three products.  Per work item (mode m, tile of 32 "n" rows):
//     D[l][n] = sum_k A[m][k][l] (*) B[m][n][k]          complex, (*) = a*b  or  conj(a)*b pattern
//   apply, forward   : A = W1 (k=i, l=o), B = X^ (n=b)          O[b][o]  = sum_i X[b][i] W[i][o]
//   apply, backward-x: A = W2 (k=o, l=i), B = G^ (n=b), conj    Gx[b][i] = sum_o G[b][o] conj(W[i][o])
//   wgrad            : A = X^ (k=b, l=i), B = G^ (n=o), conj    dW[i][o] = sum_b conj(X[b][i]) G[b][o]
// A is read with coalesced warp loads (lane = l), split hi/lo in registers and written to TMEM;
// B is turned into K-major SWIZZLE_128B rows (raw + truncation remainder) by loader threads, the real rows
// of the tile stacked on its imaginary rows: B' = [B_re ; B_im] is ONE 64-row operand, so the complex
// product takes two accumulators F = A_re B' = [A_re B_re | A_re B_im] and E = A_im B' = [A_im B_re | A_im B_im]
// (6 MMAs with N = 64 per k-step in 3xTF32 instead of 12 with N = 32: the cost of a tf32 MMA does not depend
// on N at this size) and the epilogue forms D_re = F_re -/+ E_im, D_im = F_im +/- E_re (lower sign: conj(A)).
// The contraction is short (K = Cin <= 128 or K = batch), so the 2^-11-sized correction products share the
// accumulator of the main product (see fnm_tc_tabgemm.cu for when they must not).
//
// Warps (16): 0-3 epilogue | 4 MMA issuer | 5-7 B loaders | 8-15 two A-loader warpgroups.
#include "fnm_tc_ptx.cuh"

#include <cstdlib>

namespace fnm {
namespace tc {

constexpr int kMxThreads = 16 * 32;
constexpr int kMxBWarps = 3;
constexpr int kMxN = 32;                                   // n rows per item (MMA N)

struct MxArgs {
    const float* a; const float* b; float* out;
    const float* a1; float* out1; int a_split, o_split;    // modes >= a_split read a1, modes >= o_split write out1 (weights1 | weights2 held mode-major)
    long long a_sm, a_sk, a_sl, a_im;                      // A element (m, k, l): a + m*a_sm + k*a_sk + l*a_sl (+ a_im)
    long long b_sm, b_sn, b_sk, b_im;                      // B element (m, n, k)
    long long o_sm, o_sn, o_sl, o_im;                      // output element (m, n, l)
    int Kt, Lt, Nt;                                        // contraction length, lanes (<= 128), n extent
    int nmodes, ntiles, items, nchunks;
    int conj, single_pass, a_pair, b_kcontig, o_pair;      // a_pair / o_pair: (re, im) adjacent -> 8-byte accesses
    int a_kc;                                              // A element (m, k, l) has k contiguous as (re, im) pairs: each lane reads its own 256-byte row
    int a_reuse;                                           // one A chunk per mode, kept in TMEM across the mode's n tiles
    int b_tma;                                             // TMA brings the raw B tile: 1 = rows K-contiguous (K-major operand), 2 = rows N-contiguous
                                                           // (MN-major operand, SWIZZLE_128B_BASE32B: the weight gradient's spectra)
};

__device__ __forceinline__ float ldf(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ldf2(const float* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

__device__ __forceinline__ uint64_t desc_mn32_mix(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)(4096 >> 4) << 16;                      // LBO: the imaginary half
    d |= (uint64_t)(512 >> 4) << 32;                       // SBO: next group of 4 k rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                                // SWIZZLE_128B_BASE32B
    return d;
}

// smem per B buffer: [raw | lo] x nchunks slabs of [64 rows = 32 re + 32 im][128 B]
__global__ void __launch_bounds__(kMxThreads, 1) mixgemm_tc(const __grid_constant__ CUtensorMap mapB, const MxArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t half_bytes = (uint32_t)a.nchunks * 8192u;       // raw tile | remainder tile, 64 rows per k-chunk slab
    const uint32_t bbuf_bytes = 2u * half_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * bbuf_bytes);
    uint64_t* a_full = bars;             // [2] 4 warps
    uint64_t* a_empty = bars + 2;        // [2]
    uint64_t* b_full = bars + 4;         // [2] kMxBWarps
    uint64_t* b_empty = bars + 6;        // [2]
    uint64_t* d_full = bars + 8;         // [2]
    uint64_t* d_empty = bars + 10;       // [2] 4 warps
    uint64_t* braw_full = bars + 12;     // [2] TMA transaction barrier of the raw B tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1);
            mbar_init(&b_full[s], kMxBWarps); mbar_init(&b_empty[s], 1);
            mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], 4);
            mbar_init(&braw_full[s], 1);
        }
        fence_barrier_init();
        if (a.b_tma) tma_prefetch_desc(&mapB);
    }
    if (warp == 4) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    // TMEM: A buffer s at 128*s: [re hi | re lo | im hi | im lo] x 32 columns;
    //       D buffer s at 256 + 128*s: F = [A_re B_re | A_re B_im], E = [A_im B_re | A_im B_im], 32 columns each
    // a CTA owns whole modes and runs their n tiles back to back; when the contraction fits one A chunk (weight
    // gradient: K = batch) the A operand of a mode stays in TMEM for all its n tiles (a_reuse)
    const int my_modes = (a.nmodes - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int my_items = my_modes * a.ntiles;
    auto item_of = [&](int it, int& m, int& n0) {
        const int mi = it / a.ntiles;
        m = (int)blockIdx.x + mi * (int)gridDim.x;
        n0 = (it - mi * a.ntiles) * kMxN;
    };

    if (warp < 4) {
        // =========================================================================== epilogue
        const int L = warp * 32 + lane;
        const bool l_ok = L < a.Lt;
        for (int it = 0; it < my_items; ++it) {
            int m, n0;
            item_of(it, m, n0);
            const int db = it & 1;
            mbar_wait(&d_full[db], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + 256u + 128u * db;
            float* obase = (m < a.o_split ? a.out + (size_t)m * a.o_sm : a.out1 + (size_t)(m - a.o_split) * a.o_sm) + (size_t)L * a.o_sl;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                uint32_t fr[16], fi[16], er[16], ei[16];
                tmem_ld16(taddr + g * 16, fr);
                tmem_ld16(taddr + 32 + g * 16, fi);
                tmem_ld16(taddr + 64 + g * 16, er);
                tmem_ld16(taddr + 96 + g * 16, ei);
                tmem_ld_wait();
                if (l_ok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = n0 + g * 16 + j;
                        if (n < a.Nt) {
                            // conj: 0 = A B, 1 = conj(A) B, 2 = A conj(B)
                            const float s = a.conj ? 1.f : -1.f;
                            const float vr = fmaf(s, __uint_as_float(ei[j]), __uint_as_float(fr[j]));
                            float vi = fmaf(-s, __uint_as_float(er[j]), __uint_as_float(fi[j]));
                            if (a.conj == 2) vi = -vi;
                            float* p = obase + (size_t)n * a.o_sn;
                            if (a.o_pair) *reinterpret_cast<float2*>(p) = make_float2(vr, vi);
                            else { p[0] = vr; p[a.o_im] = vi; }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[db]);
        }
    } else if (warp == 4) {
        // =========================================================================== MMA issuer
        const uint32_t sbase = smem_u32(smem);
        const bool mn = a.b_tma == 2;
        const uint32_t idesc = make_idesc(128, 2 * kMxN) | (mn ? (1u << 16) : 0u);     // MN-major B for the TMA-fed weight gradient
        const uint32_t kstep = mn ? 64u : 2u;                          // descriptor start-address units (16 B) per k-step of 8
        uint32_t q = 0;
        for (int it = 0; it < my_items; ++it) {
            const int db = it & 1, bb = it & 1;
            mbar_wait(&d_empty[db], (uint32_t)(((it >> 1) & 1) ^ 1));
            mbar_wait(&b_full[bb], (uint32_t)((it >> 1) & 1));
            const uint32_t d_f = tmem_base + 256u + 128u * db, d_e = d_f + 64u;
            const uint32_t bs = sbase + bb * bbuf_bytes;
            const int tile = it % a.ntiles;
            for (int ch = 0; ch < a.nchunks; ++ch) {
                const uint32_t ab = q & 1;
                if (!a.a_reuse || tile == 0) {
                    mbar_wait(&a_full[ab], (q >> 1) & 1);
                    tc_fence_after();
                }
                const uint32_t ar = tmem_base + 128u * ab, arl = ar + 32u, ai = ar + 64u, ail = ar + 96u;
                // K-major: 64 rows (re | im) of 128 bytes.  MN-major (SWIZZLE_128B_BASE32B): 32 k rows of 32 n per half, the im
                // half (second N block) 4096 bytes behind the re half, 4-row k groups 512 bytes apart
                const uint64_t braw = mn ? desc_mn32_mix(bs + (uint32_t)ch * 8192u) : make_desc(bs + (uint32_t)ch * 8192u);
                const uint64_t blo = mn ? desc_mn32_mix(bs + half_bytes + (uint32_t)ch * 8192u) : make_desc(bs + half_bytes + (uint32_t)ch * 8192u);
                const int nks = min(4, (a.Kt - ch * 32 + 7) / 8);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (ks < nks) {
                        const uint32_t first = (ch == 0 && ks == 0) ? 0u : 1u;
                        const uint32_t k8 = ks * 8;
                        const uint64_t br = braw + kstep * ks, bl = blo + kstep * ks;
                        umma_ts(d_f, ar + k8, br, idesc, first);
                        umma_ts(d_e, ai + k8, br, idesc, first);
                        if (!a.single_pass) {
                            umma_ts(d_f, arl + k8, br, idesc, 1);
                            umma_ts(d_e, ail + k8, br, idesc, 1);
                            umma_ts(d_f, ar + k8, bl, idesc, 1);
                            umma_ts(d_e, ai + k8, bl, idesc, 1);
                        }
                    }
                }
                if (!a.a_reuse || tile == a.ntiles - 1) {
                    umma_commit(&a_empty[ab]);
                    ++q;
                }
            }
            umma_commit(&b_empty[bb]);
            umma_commit(&d_full[db]);
            __syncwarp();
        }
    } else if (warp < 8) {
        // =========================================================================== B loaders (96 threads)
        const int bt = threadIdx.x - 5 * 32;
        // TMA variant: one thread streams the raw [re ; im] rows of an item (SWIZZLE_128B boxes of 32 k x 32 rows) one
        // item ahead; the three warps only derive the remainder tile, smem -> smem
        auto tma_item = [&](int it) {
            int m, n0;
            item_of(it, m, n0);
            const int bb = it & 1;
            uint8_t* base = smem + (size_t)bb * bbuf_bytes;
            mbar_wait(&b_empty[bb], (uint32_t)(((it >> 1) & 1) ^ 1));
            mbar_expect_tx(&braw_full[bb], half_bytes);
            for (int ch = 0; ch < a.nchunks; ++ch)
                for (int ri = 0; ri < 2; ++ri) {
                    uint8_t* dst = base + (size_t)ch * 8192 + (size_t)ri * 4096;
                    if (a.b_tma == 1) tma_load_2d(&mapB, &braw_full[bb], dst, ch * 32, (m * 2 + ri) * a.Nt + n0);
                    else tma_load_2d(&mapB, &braw_full[bb], dst, n0, (m * 2 + ri) * a.Kt + ch * 32);   // 32 k rows of 32 n
                }
        };
        if (a.b_tma && warp == 5 && lane == 0 && my_items > 0) tma_item(0);
        for (int it = 0; it < my_items; ++it) {
            int m, n0;
            item_of(it, m, n0);
            const int bb = it & 1;
            uint8_t* base = smem + (size_t)bb * bbuf_bytes;
            const float* src = a.b + (size_t)m * a.b_sm;
            if (a.b_tma) {
                mbar_wait(&braw_full[bb], (uint32_t)((it >> 1) & 1));
                const uint32_t raw0 = smem_u32(base) + (uint32_t)bt * 16u;
                for (uint32_t off = 0; off < half_bytes; off += kMxBWarps * 32 * 16 * 2) {
                    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                    const uint32_t o1 = off + kMxBWarps * 32 * 16;
                    const bool ok0 = off + (uint32_t)bt * 16u < half_bytes, ok1 = o1 + (uint32_t)bt * 16u < half_bytes;
                    if (ok0) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w) : "r"(raw0 + off));
                    if (ok1) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w) : "r"(raw0 + o1));
                    v0.x -= __uint_as_float(__float_as_uint(v0.x) & 0xffffe000u); v0.y -= __uint_as_float(__float_as_uint(v0.y) & 0xffffe000u);
                    v0.z -= __uint_as_float(__float_as_uint(v0.z) & 0xffffe000u); v0.w -= __uint_as_float(__float_as_uint(v0.w) & 0xffffe000u);
                    v1.x -= __uint_as_float(__float_as_uint(v1.x) & 0xffffe000u); v1.y -= __uint_as_float(__float_as_uint(v1.y) & 0xffffe000u);
                    v1.z -= __uint_as_float(__float_as_uint(v1.z) & 0xffffe000u); v1.w -= __uint_as_float(__float_as_uint(v1.w) & 0xffffe000u);
                    if (ok0) asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raw0 + half_bytes + off), "f"(v0.x), "f"(v0.y), "f"(v0.z), "f"(v0.w) : "memory");
                    if (ok1) asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raw0 + half_bytes + o1), "f"(v1.x), "f"(v1.y), "f"(v1.z), "f"(v1.w) : "memory");
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&b_full[bb]);
                if (warp == 5 && lane == 0 && it + 1 < my_items) tma_item(it + 1);
                continue;
            }
            mbar_wait(&b_empty[bb], (uint32_t)(((it >> 1) & 1) ^ 1));
            if (a.b_kcontig) {
                // rows n, k contiguous: 16-byte units (part re/im, row, 16-byte chunk over the padded K)
                const int kc = a.nchunks * 8;
                const int units = 2 * kMxN * kc;
                for (int u0 = 0; u0 < units; u0 += kMxBWarps * 32 * 4) {
                    float4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int u = u0 + bt + j * kMxBWarps * 32;
                        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (u < units) {
                            const int ri = u / (kMxN * kc), r2 = u - ri * kMxN * kc, n = r2 / kc, c = r2 - n * kc, k = c * 4;
                            if (n0 + n < a.Nt && k < a.Kt) {
                                const float* p = src + (size_t)ri * a.b_im + (size_t)(n0 + n) * a.b_sn + k;
                                if (k + 3 < a.Kt && (a.Kt & 3) == 0) v[j] = __ldg(reinterpret_cast<const float4*>(p));
                                else {
                                    v[j].x = __ldg(p);
                                    if (k + 1 < a.Kt) v[j].y = __ldg(p + 1);
                                    if (k + 2 < a.Kt) v[j].z = __ldg(p + 2);
                                    if (k + 3 < a.Kt) v[j].w = __ldg(p + 3);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int u = u0 + bt + j * kMxBWarps * 32;
                        if (u < units) {
                            const int ri = u / (kMxN * kc), r2 = u - ri * kMxN * kc, n = r2 / kc, c = r2 - n * kc;
                            uint8_t* dst = base + (size_t)(c >> 3) * 8192 + (size_t)ri * 4096 + swz(n, c & 7);
                            float4 l;
                            l.x = v[j].x - __uint_as_float(__float_as_uint(v[j].x) & 0xffffe000u);
                            l.y = v[j].y - __uint_as_float(__float_as_uint(v[j].y) & 0xffffe000u);
                            l.z = v[j].z - __uint_as_float(__float_as_uint(v[j].z) & 0xffffe000u);
                            l.w = v[j].w - __uint_as_float(__float_as_uint(v[j].w) & 0xffffe000u);
                            *reinterpret_cast<float4*>(dst) = v[j];
                            *reinterpret_cast<float4*>(dst + half_bytes) = l;
                        }
                    }
                }
            } else {
                // k strided (n contiguous): one thread per (part, row n) gathers its K-row 32 values at a time
                for (int task = bt; task < 2 * kMxN; task += kMxBWarps * 32) {
                    const int ri = task / kMxN, n = task - ri * kMxN;
                    const bool n_ok = n0 + n < a.Nt;
                    const float* p = src + (size_t)ri * a.b_im + (size_t)(n0 + n) * a.b_sn;
                    for (int ch = 0; ch < a.nchunks; ++ch) {
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int k = ch * 32 + j;
                            v[j] = (n_ok && k < a.Kt) ? __ldg(p + (size_t)k * a.b_sk) : 0.f;
                        }
                        uint8_t* dst = base + (size_t)ch * 8192 + (size_t)ri * 4096;
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4) {
                            const float4 w = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
                            float4 l;
                            l.x = w.x - __uint_as_float(__float_as_uint(w.x) & 0xffffe000u);
                            l.y = w.y - __uint_as_float(__float_as_uint(w.y) & 0xffffe000u);
                            l.z = w.z - __uint_as_float(__float_as_uint(w.z) & 0xffffe000u);
                            l.w = w.w - __uint_as_float(__float_as_uint(w.w) & 0xffffe000u);
                            *reinterpret_cast<float4*>(dst + swz(n, c4)) = w;
                            *reinterpret_cast<float4*>(dst + half_bytes + swz(n, c4)) = l;
                        }
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&b_full[bb]);
        }
    } else {
        // =========================================================================== A loaders (two warpgroups, buffer = warpgroup)
        const int wg = (warp - 8) >> 2;
        const int q4 = warp & 3;
        const int L = q4 * 32 + lane;
        const bool l_ok = L < a.Lt;
        const uint32_t tb = tmem_base + ((uint32_t)(q4 * 32) << 16) + 128u * wg;
        const uint32_t total = a.a_reuse ? (uint32_t)my_modes : (uint32_t)my_items * (uint32_t)a.nchunks;
        for (uint32_t q = (uint32_t)wg; q < total; q += 2) {
            const int it = a.a_reuse ? (int)q * a.ntiles : (int)(q / (uint32_t)a.nchunks);
            const int ch = a.a_reuse ? 0 : (int)(q - (uint32_t)it * a.nchunks);
            int m, n0_unused;
            item_of(it, m, n0_unused);
            const float* amode = m < a.a_split ? a.a + (size_t)m * a.a_sm : a.a1 + (size_t)(m - a.a_split) * a.a_sm;
            const float* p = amode + (size_t)(ch * 32) * a.a_sk + (size_t)L * a.a_sl;
            const int klim = l_ok ? a.Kt - ch * 32 : 0;
            float vr[32], vi[32];
            if (a.a_kc) {
                // backward-x on the FORWARD's mode-major copy W1[m][i][o] (lanes = i, k = o): a lane's 32 (re, im) pairs
                // are one 256-byte run; L1 keeps the sectors between the two 16-byte loads that share them
                const float4* p4 = reinterpret_cast<const float4*>(amode + (size_t)L * a.a_sl + (size_t)(ch * 32) * 2);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (2 * j < klim) v = __ldg(p4 + j);
                    vr[2 * j] = v.x; vi[2 * j] = v.y;
                    vr[2 * j + 1] = 2 * j + 1 < klim ? v.z : 0.f; vi[2 * j + 1] = 2 * j + 1 < klim ? v.w : 0.f;
                }
            } else if (a.a_pair) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float2 v = make_float2(0.f, 0.f);
                    if (j < klim) v = ldf2(p);
                    vr[j] = v.x; vi[j] = v.y;
                    p += a.a_sk;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    vr[j] = j < klim ? ldf(p) : 0.f;
                    vi[j] = j < klim ? ldf(p + a.a_im) : 0.f;
                    p += a.a_sk;
                }
            }
            mbar_wait(&a_empty[wg], ((q >> 1) & 1) ^ 1);
            tc_fence_after();
            split_to_tmem16(tb, tb + 32u, vr, !a.single_pass);
            split_to_tmem16(tb + 16u, tb + 48u, vr + 16, !a.single_pass);
            split_to_tmem16(tb + 64u, tb + 96u, vi, !a.single_pass);
            split_to_tmem16(tb + 80u, tb + 112u, vi + 16, !a.single_pass);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[wg]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// out[c][b][a] = in[a][b][c] for 8-byte (complex64) elements; grid (ceil(C/32), ceil(A/32), B * ntens): the
// two weight tensors of a SpectralConv2d (weights1 | weights2) are permuted by ONE launch (blockIdx.z / B picks it)
__global__ void permute_c64_cba(const float2* __restrict__ in0, const float2* __restrict__ in1, float2* __restrict__ out0,
                                float2* __restrict__ out1, int A, int B, int C, long long out_c_stride) {
    __shared__ float2 tile[32][33];
    const int t = blockIdx.z / B, b = blockIdx.z - t * B;
    const float2* in = t ? in1 : in0;
    float2* out = t ? out1 : out0;
    const int c0 = blockIdx.x * 32, a0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;              // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int aa = a0 + r, cc = c0 + tx;
        if (aa < A && cc < C) tile[r][tx] = in[((size_t)aa * B + b) * C + cc];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int cc = c0 + r, aa = a0 + tx;
        if (aa < A && cc < C) out[(size_t)cc * out_c_stride + (size_t)b * A + aa] = tile[tx][r];
    }
}

static int mx_launch(MxArgs& a, const char* what, cudaStream_t st) {
    a.nchunks = (a.Kt + 31) / 32;
    a.ntiles = (a.Nt + kMxN - 1) / kMxN;
    const long long items = (long long)a.nmodes * a.ntiles;
    if (items > (1LL << 30)) return fail(FNM_EINVAL, "%s: too many work items", what);
    a.items = (int)items;
    a.a_reuse = (a.nchunks == 1 && a.ntiles > 1) ? 1 : 0;
    size_t smem = 2 * 4 * (size_t)a.nchunks * 4096 + 1024 + 256;
    if (smem < 120 * 1024) smem = 120 * 1024;
    static std::atomic<uint64_t> attr_set{0};
    int attr_set_dev = 0;
    if (once_needed(attr_set, attr_set_dev)) {
        const cudaError_t e = cudaFuncSetAttribute(mixgemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return fail(FNM_ECUDA, "%s: smem attribute: %s", what, cudaGetErrorString(e));
        once_done(attr_set, attr_set_dev);
    }
    // raw B tile by TMA when its rows are K-contiguous: tensor [nmodes * 2 * Nt rows][Kt], box 32 x 32
    CUtensorMap mapB{};
    a.b_tma = 0;
    if (a.b_kcontig && a.Kt % 4 == 0 && (reinterpret_cast<uintptr_t>(a.b) & 15) == 0 && a.b_sn == a.Kt &&
        a.b_im == (long long)a.Nt * a.Kt && a.b_sm == 2 * a.b_im) {
        const cuuint64_t dims[2] = {(cuuint64_t)a.Kt, (cuuint64_t)a.nmodes * 2 * (cuuint64_t)a.Nt};
        const cuuint64_t str[1] = {(cuuint64_t)a.Kt * 4};
        const cuuint32_t box[2] = {32, (cuuint32_t)kMxN};
        if (tc_make_map(&mapB, a.b, 2, dims, str, box, true)) a.b_tma = 1;
    } else if (!a.b_kcontig && a.Nt % 4 == 0 && (reinterpret_cast<uintptr_t>(a.b) & 15) == 0 && a.b_sn == 1 &&
               a.b_sk == a.Nt && a.b_im == (long long)a.Kt * a.Nt && a.b_sm == 2 * a.b_im) {
        // rows are N-contiguous (weight gradient: B = X^[m][.][b][i]): tensor [nmodes * 2 * Kt rows][Nt], boxes of 32 k x 32 n
        const cuuint64_t dims[2] = {(cuuint64_t)a.Nt, (cuuint64_t)a.nmodes * 2 * (cuuint64_t)a.Kt};
        const cuuint64_t str[1] = {(cuuint64_t)a.Nt * 4};
        const cuuint32_t box[2] = {32, 32};
        if (tc_make_map_swz(&mapB, a.b, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) a.b_tma = 2;
    }
    const int grid = a.nmodes < sm_count() ? a.nmodes : sm_count();
    {
        LaunchScope ls(what, st);
        mixgemm_tc<<<grid, kMxThreads, smem, st>>>(mapB, a);
    }
    return check_launch(what);
}

}  // namespace tc

using namespace tc;

bool tc_mix_supported(int B, int Ci, int Co) {
    if (!tc_enabled()) return false;
    { const char* e = getenv("FNM_DISABLE_TCMIX"); if (e && e[0] == '1') return false; }
    return B >= 1 && B <= 192 && Ci >= 8 && Co >= 8 && Ci <= 128 && Co <= 128;   // wgrad stages K = B in shared memory
}

// scratch (floats) for the mode-major copies: W1 + W2 (2 x nm*Ci*Co complex) or dW (nm*Ci*Co complex)
size_t tc_mix_shadow_floats(int nmodes, int Ci, int Co) { return (size_t)nmodes * Ci * Co * 2; }

// W (Ci,Co,rpw*NY) [x2 tensors] -> W1[m][i][o] (which = 0) or W2[m][o][i] (which = 1)
int tc_mix_pack(const float* w0, const float* w1, int rpw, int R, int NY, int Ci, int Co, int which, float* dst,
                cudaStream_t st) {
    const int nm = rpw * NY;                                    // modes per weight tensor
    const int ntens = (R == rpw) ? 1 : 2;
    const float2* s0 = reinterpret_cast<const float2*>(w0);
    const float2* s1 = reinterpret_cast<const float2*>(w1 ? w1 : w0);
    float2* o0 = reinterpret_cast<float2*>(dst);
    float2* o1 = o0 + (size_t)nm * Ci * Co;
    {
        LaunchScope ls("mix_pack", st);
        if (which == 0) {
            // in[a = (i,o)][b = 1][c = m] -> out[m][(i,o)]
            dim3 grid((nm + 31) / 32, (Ci * Co + 31) / 32, ntens);
            permute_c64_cba<<<grid, 256, 0, st>>>(s0, s1, o0, o1, Ci * Co, 1, nm, (long long)Ci * Co);
        } else {
            // in[a = i][b = o][c = m] -> out[m][o][i]
            dim3 grid((nm + 31) / 32, (Ci + 31) / 32, Co * ntens);
            permute_c64_cba<<<grid, 256, 0, st>>>(s0, s1, o0, o1, Ci, Co, nm, (long long)Ci * Co);
        }
    }
    return check_launch("mix_pack");
}

// dW1[m][i][o] -> dw (Ci,Co,rpw*NY) [x2 tensors]
int tc_mix_unpack(const float* dW1, float* dw0, float* dw1, int rpw, int R, int NY, int Ci, int Co, cudaStream_t st) {
    const int nm = rpw * NY;
    const int ntens = (R == rpw) ? 1 : 2;
    const float2* s0 = reinterpret_cast<const float2*>(dW1);
    const float2* s1 = s0 + (size_t)nm * Ci * Co;
    {
        LaunchScope ls("mix_unpack", st);
        // in[a = m][b = 1][c = (i,o)] -> out[(i,o)][m]
        dim3 grid((Ci * Co + 31) / 32, (nm + 31) / 32, ntens);
        permute_c64_cba<<<grid, 256, 0, st>>>(s0, s1, reinterpret_cast<float2*>(dw0),
                                                 reinterpret_cast<float2*>(dw1 ? dw1 : dw0), nm, 1, Ci * Co, (long long)nm);
    }
    return check_launch("mix_unpack");
}

// Oh[m][ro][b][o] = sum_i Xh[m][ri][b][i] W1[m][i][o]        (conj = 0, W = W1, Kt = Ci, Lt = Co)
// Gxh[m][ri][b][i] = sum_o Gh[m][ro][b][o] conj(W[i][o][m])   (conj = 1, Kt = Co, Lt = Ci) with W given either as
//   W2[m][o][i] (w_is_w1 = 0) or as the forward's copy W1[m][i][o] (w_is_w1 = 1: lanes read k-contiguous rows)
//   Wm1 / m_split: modes >= m_split are read from Wm1 (mode-major parameters: weights1 | weights2 are two arrays)
int tc_mix_apply(const float* Wm, const float* Wm1, int m_split, const float* In, float* Out, int nmodes, int B, int Kt,
                 int Lt, int conj, int w_is_w1, int mode, const char* what, cudaStream_t st) {
    MxArgs a{};
    a.a = Wm; a.b = In; a.out = Out;
    a.a1 = Wm1 ? Wm1 : Wm; a.a_split = Wm1 ? m_split : nmodes; a.out1 = Out; a.o_split = nmodes;
    a.Kt = Kt; a.Lt = Lt; a.Nt = B; a.nmodes = nmodes;
    a.a_sm = (long long)Kt * Lt * 2; a.a_sk = (long long)Lt * 2; a.a_sl = 2; a.a_im = 1; a.a_pair = 1;
    if (conj && w_is_w1) { a.a_sk = 2; a.a_sl = (long long)Kt * 2; a.a_kc = 1; }
    a.b_sm = (long long)2 * B * Kt; a.b_sn = Kt; a.b_sk = 1; a.b_im = (long long)B * Kt; a.b_kcontig = 1;
    a.o_sm = (long long)2 * B * Lt; a.o_sn = Lt; a.o_sl = 1; a.o_im = (long long)B * Lt; a.o_pair = 0;
    a.conj = conj; a.single_pass = mode == FNM_MODE_TF32;
    return mx_launch(a, what, st);
}

// dW1[m][i][o] = sum_b conj(Xh[m][.][b][i]) Gh[m][.][b][o].  The output-gradient spectrum is the TMEM operand
// (lanes = o) so that the epilogue's warp stores run along o, the contiguous axis of dW1 (with lanes = i every
// 8-byte store of a warp hit a different 1 KB row); the conjugate then sits on the shared-memory operand.
int tc_mix_wgrad(const float* Xh, const float* Gh, float* dW1, float* dW1b, int m_split, int nmodes, int B, int Ci, int Co,
                 int mode, cudaStream_t st, const char* what) {
    MxArgs a{};
    a.a = Gh; a.b = Xh; a.out = dW1;
    a.a1 = Gh; a.a_split = nmodes; a.out1 = dW1b ? dW1b : dW1; a.o_split = dW1b ? m_split : nmodes;
    a.Kt = B; a.Lt = Co; a.Nt = Ci; a.nmodes = nmodes;
    a.a_sm = (long long)2 * B * Co; a.a_sk = Co; a.a_sl = 1; a.a_im = (long long)B * Co; a.a_pair = 0;
    a.b_sm = (long long)2 * B * Ci; a.b_sn = 1; a.b_sk = Ci; a.b_im = (long long)B * Ci; a.b_kcontig = 0;
    a.o_sm = (long long)Ci * Co * 2; a.o_sn = (long long)Co * 2; a.o_sl = 2; a.o_im = 1; a.o_pair = 1;
    a.conj = 2; a.single_pass = mode == FNM_MODE_TF32;
    return mx_launch(a, what, st);
}

}  // namespace fnm
