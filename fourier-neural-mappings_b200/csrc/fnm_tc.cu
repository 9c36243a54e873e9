// tcgen05 / TMEM kernels (sm_100a) for the heavy spatial passes of the Fourier layer.
//
// tc_pointwise_ysyn -- the fused "1x1 conv + inverse DFT along y + bias (+ GELU')" pass, i.e. the
// kernel that reads an activation once and writes one once.  Formulated TRANSPOSED so that the
// channel axis is the MMA M dimension and the positions of one (b,x) grid row are the N dimension:
//
//     D[o, y] = sum_i Wk[i,o] f(X[y,i])  +  sum_l' Z[l',o] by[y,l']        (one accumulator in TMEM)
//
//   A = [ Wk^T | Z[row]^T ]  (M = channels): the constant weight block lives in TMEM for the whole
//                            kernel (M = 128) or in shared memory (M = 64); Z[row] is staged per row
//   B = [ f(X) | by ]        (N = positions): staged per tile by producer warps; rows are channel /
//                            mode contiguous, i.e. K-major, written in the canonical SWIZZLE_128B layout
//
// fp32 parity mode is 3xTF32: every operand is split in registers into hi = tf32(v) and lo = v - hi
// and each k-step issues hi*hi + hi*lo + lo*hi into the same TMEM accumulator (error ~2^-21).
// FNM_MODE_TF32 issues hi*hi only.  Because the activation has to pass through registers anyway
// (GELU on load + the split), it is loaded with coalesced 16-byte LDGs rather than TMA.
//
// Warp roles (512 threads, 1 CTA / SM, persistent over grid rows):
//   warps 0-3  epilogue: tcgen05.ld -> +bias, *GELU'(z_prev) -> coalesced 128-byte stores (lane = channel)
//   warp  4    TMEM allocator + single-thread tcgen05.mma issuer
//   warps 5-15 producers (global -> registers -> split -> swizzled shared memory)
// Pipelines: smem stage full/empty (producers <-> MMA), Z-row full/empty, TMEM accumulator full/empty.
#include "fnm_tc_ptx.cuh"

#include <atomic>

#include <cstdlib>

namespace fnm {
namespace tc {

constexpr int kEpiWarps = 4;
constexpr int kProdWarps = 11;           // 16 warps total: 4 per SM sub-partition (128 registers each)
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kStages = 2;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccCol0 = 256;        // accumulators at columns 256 + 128*acc
constexpr int kMaxU1 = 6;                 // activation float4 units per producer thread per tile
constexpr int kMaxU2 = 3;                 // twiddle float4 units per producer thread per tile

// ---------------------------------------------------------------------------------- kernel
struct PwArgs {
    const float* x; const float* wc; const float* bias; const float* Z; const float* by; const float* mul;
    float* out; float* out_act;
    long long rows;
    int S2, LY, CI, CO, act, wc_is_kn, single_pass;
    int NT, tiles_per_row, K2p, K2slabs, M;
    int debug;      // FNM_TC_DEBUG bit mask (perf experiments only): 1 skip MMAs, 2 skip producer stores, 4 skip epilogue stores, 8 hi*hi K1 only
};

struct PwSmem {
    // byte offsets from the 1024-aligned base
    uint32_t stage_bytes, b1_hi, b1_lo, b2_hi, b2_lo;    // within a stage
    uint32_t a2_hi, a2_lo, a1_hi, a1_lo, bars, total;
};

__host__ __device__ inline PwSmem pw_layout(int NT, int CI, int K2slabs, int M, bool a1_smem) {
    PwSmem s;
    const uint32_t b1 = (uint32_t)(CI / kSlabK) * NT * 128, b2 = (uint32_t)K2slabs * NT * 128;
    s.b1_hi = 0; s.b1_lo = b1; s.b2_hi = 2 * b1; s.b2_lo = 2 * b1 + b2;
    s.stage_bytes = 2 * b1 + 2 * b2;
    uint32_t off = kStages * s.stage_bytes;
    const uint32_t a2 = (uint32_t)K2slabs * M * 128;
    s.a2_hi = off; s.a2_lo = off + a2; off += 2 * a2;
    const uint32_t a1 = a1_smem ? (uint32_t)(CI / kSlabK) * M * 128 : 0;
    s.a1_hi = off; s.a1_lo = off + a1; off += 2 * a1;
    s.bars = off; off += 128;
    s.total = off;
    return s;
}

template <bool A1_TMEM>
__global__ void __launch_bounds__(kThreads, 1) pointwise_ysyn_tc(const PwArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const PwSmem L = pw_layout(a.NT, a.CI, a.K2slabs, a.M, !A1_TMEM);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* full = bars;            // [2]
    uint64_t* empty = bars + 2;       // [2]
    uint64_t* a2_full = bars + 4;
    uint64_t* a2_empty = bars + 5;
    uint64_t* acc_full = bars + 6;    // [2]
    uint64_t* acc_empty = bars + 8;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    // warp index through a shuffle so the compiler knows it is warp-uniform (role branches and the MMA
    // issue loop then use uniform registers / uniform branches)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int K1slabs = a.CI / kSlabK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], kProdWarps); mbar_init(&empty[s], 1); }
        mbar_init(a2_full, kProdWarps); mbar_init(a2_empty, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kEpiWarps); }
        fence_barrier_init();
    }
    if (warp == kEpiWarps) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // warp-uniform

    // ---- constant pointwise weight block A1[m = out channel][k = in channel] = Wk[k][m]
    if (a.CI > 0) {
        if (A1_TMEM) {
            if (warp < kEpiWarps) {
                const int m = warp * 32 + lane;          // TMEM lane = output channel (M = 128)
                for (int kc = 0; kc < K1slabs; ++kc) {
                    uint32_t hi[32], lo[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int k = kc * 32 + j;
                        float v = 0.f;
                        if (m < a.CO) v = a.wc_is_kn ? __ldg(a.wc + (size_t)k * a.CO + m) : __ldg(a.wc + (size_t)m * a.CI + k);
                        float h, l;
                        split_tf32(v, h, l);
                        hi[j] = __float_as_uint(h); lo[j] = __float_as_uint(l);
                    }
                    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
                    tmem_st32(lane_base + kc * 32, hi);
                    tmem_st32(lane_base + a.CI + kc * 32, lo);
                }
                tmem_st_wait();
            }
        } else {
            for (int u = threadIdx.x; u < a.M * (a.CI / 4); u += kThreads) {
                const int m = u / (a.CI / 4), c4 = u - m * (a.CI / 4), k = c4 * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < a.CO) {
                    if (a.wc_is_kn) {
                        v.x = __ldg(a.wc + (size_t)(k + 0) * a.CO + m); v.y = __ldg(a.wc + (size_t)(k + 1) * a.CO + m);
                        v.z = __ldg(a.wc + (size_t)(k + 2) * a.CO + m); v.w = __ldg(a.wc + (size_t)(k + 3) * a.CO + m);
                    } else {
                        v = __ldg(reinterpret_cast<const float4*>(a.wc + (size_t)m * a.CI + k));
                    }
                }
                const int kb = k / kSlabK;
                store_split4(smem + L.a1_hi + (size_t)kb * a.M * 128, smem + L.a1_lo + (size_t)kb * a.M * 128, m, c4 & 7, v);
            }
            fence_proxy_async();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const long long my_rows = (a.rows - (long long)blockIdx.x + gridDim.x - 1) / gridDim.x;   // rows handled by this CTA
    const uint32_t idesc = make_idesc(a.M, a.NT);

    if (warp < kEpiWarps) {
        // =============================================================== epilogue
        const int o = (a.M == 128) ? warp * 32 + lane : warp * 16 + lane;       // output channel of this TMEM lane
        const bool lane_ok = (a.M == 128 || lane < 16) && o < a.CO;
        const float bias = (lane_ok && a.bias) ? __ldg(a.bias + o) : 0.f;
        long long t = 0;
        for (long long ri = 0; ri < my_rows; ++ri) {
            const long long row = (long long)blockIdx.x + ri * gridDim.x;
            for (int tile = 0; tile < a.tiles_per_row; ++tile, ++t) {
                const int acc = (int)(t & 1);
                mbar_wait(&acc_full[acc], (uint32_t)((t >> 1) & 1));
                tc_fence_after();
                const int y0 = tile * a.NT;
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + kAccCol0 + acc * 128;
                const int nvalid = min(a.NT, a.S2 - y0);                          // positions of this tile inside the row
                for (int c0 = 0; c0 < nvalid; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(taddr + c0, r);
                    tmem_ld_wait();
                    if (lane_ok && !(a.debug & 4)) {
                        const size_t base = ((size_t)row * a.S2 + y0 + c0) * a.CO + o;
                        const int cnt = min(16, nvalid - c0);
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + bias;
                        if (a.mul) {
                            const float* mp = a.mul + base;
                            float z[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) z[j] = j < cnt ? __ldg(mp + (size_t)j * a.CO) : 0.f;
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] *= gelu_grad_fast(z[j]);
                        }
                        float* op = a.out + base;
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (j < cnt) op[(size_t)j * a.CO] = v[j];
                        if (a.out_act) {
                            float* ap = a.out_act + base;
#pragma unroll
                            for (int j = 0; j < 16; ++j) if (j < cnt) ap[(size_t)j * a.CO] = gelu_fast(v[j]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);
            }
        }
    } else if (warp == kEpiWarps) {
        // =============================================================== MMA issuer (warp-uniform, elected lane)
        {
            const uint32_t sbase = smem_u32(smem);
            long long t = 0;
            for (long long ri = 0; ri < my_rows; ++ri) {
                for (int tile = 0; tile < a.tiles_per_row; ++tile, ++t) {
                    const int s = (int)(t & 1), acc = s;
                    const uint32_t ph = (uint32_t)((t >> 1) & 1);
                    mbar_wait(&acc_empty[acc], ph ^ 1);
                    mbar_wait(&full[s], ph);
                    if (tile == 0) mbar_wait(a2_full, (uint32_t)(ri & 1));
                    tc_fence_after();
                    const uint32_t d = tmem_base + kAccCol0 + acc * 128;
                    const uint32_t st = sbase + s * L.stage_bytes;
                    uint32_t accum = 0;
                    // descriptors differ only in the 14-bit start address: +2 per 32-byte k-step
                    for (int kb = 0; kb < ((a.debug & 1) ? 0 : K1slabs); ++kb) {
                        const uint64_t bhi0 = make_desc(st + L.b1_hi + (uint32_t)kb * a.NT * 128);
                        const uint64_t blo0 = make_desc(st + L.b1_lo + (uint32_t)kb * a.NT * 128);
                        const uint64_t ahi0 = A1_TMEM ? 0 : make_desc(sbase + L.a1_hi + (uint32_t)kb * a.M * 128);
                        const uint64_t alo0 = A1_TMEM ? 0 : make_desc(sbase + L.a1_lo + (uint32_t)kb * a.M * 128);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t bhi = bhi0 + 2 * ks, blo = blo0 + 2 * ks;
                            if (A1_TMEM) {
                                const uint32_t ahi = tmem_base + kb * 32 + ks * 8, alo = ahi + a.CI;
                                umma_ts(d, ahi, bhi, idesc, accum); accum = 1;
                                if (!a.single_pass) { umma_ts(d, ahi, blo, idesc, 1); umma_ts(d, alo, bhi, idesc, 1); }
                            } else {
                                umma_ss(d, ahi0 + 2 * ks, bhi, idesc, accum); accum = 1;
                                if (!a.single_pass) { umma_ss(d, ahi0 + 2 * ks, blo, idesc, 1); umma_ss(d, alo0 + 2 * ks, bhi, idesc, 1); }
                            }
                        }
                    }
                    for (int kb = 0; kb < ((a.debug & 1) ? 0 : a.K2slabs); ++kb) {
                        const int nks = min(4, (a.K2p - kb * kSlabK) / 8);
                        const uint64_t bhi0 = make_desc(st + L.b2_hi + (uint32_t)kb * a.NT * 128);
                        const uint64_t blo0 = make_desc(st + L.b2_lo + (uint32_t)kb * a.NT * 128);
                        const uint64_t ahi0 = make_desc(sbase + L.a2_hi + (uint32_t)kb * a.M * 128);
                        const uint64_t alo0 = make_desc(sbase + L.a2_lo + (uint32_t)kb * a.M * 128);
                        for (int ks = 0; ks < nks; ++ks) {
                            umma_ss(d, ahi0 + 2 * ks, bhi0 + 2 * ks, idesc, accum); accum = 1;
                            if (!a.single_pass) {
                                umma_ss(d, ahi0 + 2 * ks, blo0 + 2 * ks, idesc, 1);
                                umma_ss(d, alo0 + 2 * ks, bhi0 + 2 * ks, idesc, 1);
                            }
                        }
                    }
                    umma_commit(&empty[s]);                 // smem stage reusable once these MMAs retire
                    umma_commit(&acc_full[acc]);            // accumulator ready for the epilogue
                    if (tile == a.tiles_per_row - 1) umma_commit(a2_empty);
                    __syncwarp();
                }
            }
        }
    } else {
        // =============================================================== producers
        // Everything that does not depend on the tile (which unit a thread owns, where it lands in the
        // swizzled stage) is computed ONCE; per tile a thread issues its 16-byte loads, waits for the
        // stage, splits hi/lo and stores.  Loads of tile t+1 are issued before the stage wait of t+1,
        // i.e. right after the stores of tile t, so HBM latency overlaps the MMA of tile t.
        const int pt = threadIdx.x - (kEpiWarps + 1) * 32;
        const int ci4 = a.CI / 4;
        const int k2c = a.K2slabs * 8;                     // 16-byte chunks per row over the padded mode axis
        const int n_a2 = k2c * a.M;
        const uint32_t lo1 = L.b1_lo - L.b1_hi, lo2 = L.b2_lo - L.b2_hi, loa = L.a2_lo - L.a2_hi;
        const bool by_vec = (a.LY % 4) == 0;
        // packed unit descriptors: ux = position index << 16 | global offset inside the tile (-1: unused);
        // sx = byte offset of the hi chunk inside a stage.  ub = position << 24 | first mode << 16 | offset.
        int ux[kMaxU1], ub[kMaxU2];
        uint32_t sx[kMaxU1], sb[kMaxU2];
#pragma unroll
        for (int j = 0; j < kMaxU1; ++j) {
            const int u = pt + j * kProdThreads;
            ux[j] = -1; sx[j] = 0;
            if (ci4 > 0 && u < a.NT * ci4) {
                const int n = u / ci4, c4 = u - n * ci4;
                ux[j] = (n << 16) | (n * a.CI + c4 * 4);
                sx[j] = L.b1_hi + (uint32_t)((c4 * 4) / kSlabK) * a.NT * 128 + swz(n, c4 & 7);
            }
        }
#pragma unroll
        for (int j = 0; j < kMaxU2; ++j) {
            const int u = pt + j * kProdThreads;
            ub[j] = -1; sb[j] = 0;
            if (u < a.NT * k2c) {
                const int n = u / k2c, c = u - n * k2c;
                ub[j] = (n << 24) | ((c * 4) << 16) | (n * a.LY + c * 4);
                sb[j] = L.b2_hi + (uint32_t)((c * 4) / kSlabK) * a.NT * 128 + swz(n, c & 7);
            }
        }
        long long t = 0;
        float4 xr[kMaxU1];                                 // prefetched activation units of the next tile

        auto load_x = [&](long long row, int tile) {
            const int y0 = tile * a.NT;
            const float* src = a.x + ((size_t)row * a.S2 + y0) * a.CI;
            const int lim = a.S2 - y0;
#pragma unroll
            for (int j = 0; j < kMaxU1; ++j) {
                xr[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ux[j] >= 0 && (ux[j] >> 16) < lim) xr[j] = ld_stream4(src + (ux[j] & 0xffff));
            }
        };
        auto fill_b = [&](int tile, long long tt, long long nrow, int ntile, bool has_next) {
            const int s = (int)(tt & 1);
            const int y0 = tile * a.NT;
            const int lim = a.S2 - y0;
            float4 br[kMaxU2];                             // twiddles: L2-resident table, loaded before the wait
            const float* bsrc = a.by + (size_t)y0 * a.LY;
#pragma unroll
            for (int j = 0; j < kMaxU2; ++j) {
                br[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ub[j] >= 0 && (ub[j] >> 24) < lim) {
                    const int k0 = (ub[j] >> 16) & 0xff, off = ub[j] & 0xffff;
                    if (by_vec) {
                        if (k0 < a.LY) br[j] = __ldg(reinterpret_cast<const float4*>(bsrc + off));
                    } else {
                        float vv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int q = 0; q < 4; ++q) if (k0 + q < a.LY) vv[q] = __ldg(bsrc + off + q);
                        br[j] = make_float4(vv[0], vv[1], vv[2], vv[3]);
                    }
                }
            }
            mbar_wait(&empty[s], (uint32_t)(((tt >> 1) & 1) ^ 1));
            uint8_t* st = smem + (size_t)s * L.stage_bytes;
#pragma unroll
            for (int j = 0; j < kMaxU1; ++j) {
                if (ux[j] >= 0 && !(a.debug & 2)) {
                    float4 v = xr[j];
                    if (a.act) { v.x = gelu_fast(v.x); v.y = gelu_fast(v.y); v.z = gelu_fast(v.z); v.w = gelu_fast(v.w); }
                    float4 h, l;
                    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
                    *reinterpret_cast<float4*>(st + sx[j]) = h;
                    *reinterpret_cast<float4*>(st + sx[j] + lo1) = l;
                }
            }
            if (has_next) load_x(nrow, ntile);             // next tile's activation loads are now in flight
#pragma unroll
            for (int j = 0; j < kMaxU2; ++j) {
                if (ub[j] >= 0) {
                    float4 h, l;
                    split_tf32(br[j].x, h.x, l.x); split_tf32(br[j].y, h.y, l.y);
                    split_tf32(br[j].z, h.z, l.z); split_tf32(br[j].w, h.w, l.w);
                    *reinterpret_cast<float4*>(st + sb[j]) = h;
                    *reinterpret_cast<float4*>(st + sb[j] + lo2) = l;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        };

        if (my_rows > 0) load_x((long long)blockIdx.x, 0);
        for (long long ri = 0; ri < my_rows; ++ri) {
            const long long row = (long long)blockIdx.x + ri * gridDim.x;
            const bool last_row = ri + 1 == my_rows;
            {
                const bool one = a.tiles_per_row == 1;
                fill_b(0, t, one ? row + gridDim.x : row, one ? 0 : 1, !(one && last_row));
            }
            // A2: Z[row]^T, rows = output channel, K = l'  (consecutive threads -> consecutive channels)
            const float* zsrc = a.Z + (size_t)row * a.LY * a.CO;
            for (int u0 = 0; u0 < n_a2; u0 += kProdThreads * 4) {
                float4 zr[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int u = u0 + pt + j * kProdThreads;
                    float vv[4] = {0.f, 0.f, 0.f, 0.f};
                    if (u < n_a2) {
                        const int c = u / a.M, m = u - c * a.M, k = c * 4;
                        if (m < a.CO) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) if (k + q < a.LY) vv[q] = __ldg(zsrc + (size_t)(k + q) * a.CO + m);
                        }
                    }
                    zr[j] = make_float4(vv[0], vv[1], vv[2], vv[3]);
                }
                if (u0 == 0) mbar_wait(a2_empty, (uint32_t)((ri & 1) ^ 1));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int u = u0 + pt + j * kProdThreads;
                    if (u < n_a2) {
                        const int c = u / a.M, m = u - c * a.M;
                        uint8_t* dst = smem + L.a2_hi + (size_t)((c * 4) / kSlabK) * a.M * 128 + swz(m, c & 7);
                        float4 h, l;
                        split_tf32(zr[j].x, h.x, l.x); split_tf32(zr[j].y, h.y, l.y);
                        split_tf32(zr[j].z, h.z, l.z); split_tf32(zr[j].w, h.w, l.w);
                        *reinterpret_cast<float4*>(dst) = h;
                        *reinterpret_cast<float4*>(dst + loa) = l;
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a2_full);
            ++t;
            for (int tile = 1; tile < a.tiles_per_row; ++tile, ++t) {
                const bool last_tile = tile + 1 == a.tiles_per_row;
                fill_b(tile, t, last_tile ? row + gridDim.x : row, last_tile ? 0 : tile + 1, !(last_tile && last_row));
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps) tmem_dealloc(tmem_base, kTmemCols);
}

static int g_sm_count = 0;
static std::atomic<int> g_sm_limit{0};
int sm_count_max() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_sm_count <= 0)
            g_sm_count = 148;
    }
    return g_sm_count;
}
// CTAs of a persistent launch: every SM, or the caller's budget (fnm_set_sm_limit: SMs left to a concurrent collective)
int sm_count() {
    const int n = sm_count_max(), lim = g_sm_limit.load(std::memory_order_relaxed);
    return (lim > 0 && lim < n) ? lim : n;
}
int sm_limit_exchange(int n) { return g_sm_limit.exchange(n < 0 ? 0 : n); }

bool tc_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FNM_DISABLE_TC");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

// choose positions per tile: multiple of 16, <= 128, fits shared memory, least padding waste
static int choose_nt(int S2, int CI, int K2slabs, int M, bool a1_smem) {
    int best = 0;
    double best_cost = 1e30;
    for (int nt = 16; nt <= 128; nt += 16) {
        const PwSmem L = pw_layout(nt, CI, K2slabs, M, a1_smem);
        if (L.total + 1024 > 227 * 1024) continue;
        if (nt * (CI / 4) > kMaxU1 * kProdThreads || nt * K2slabs * 8 > kMaxU2 * kProdThreads) continue;
        const int tiles = (S2 + nt - 1) / nt;
        const double cost = (double)tiles * nt / S2 + 0.02 * tiles;     // padding waste + per-tile overhead
        if (cost < best_cost) { best_cost = cost; best = nt; }
    }
    return best;
}

}  // namespace tc

int tc_available() { return 1; }

int tc_set_sm_limit(int n) { return tc::sm_limit_exchange(n); }

bool tc_pointwise_ysyn_supported(int64_t rows, int S2, int LY, int K, int N) {
    if (!tc::tc_enabled()) return false;
    if (rows < 1 || S2 < 8 || LY < 1 || LY > 128 || N < 1 || N > 128) return false;
    if (K != 0 && (K % 32 != 0 || K > 128)) return false;
    const int M = N > 64 ? 128 : 64;
    const int K2slabs = ((LY + 7) / 8 * 8 + 31) / 32;
    return tc::choose_nt(S2, K, K2slabs, M, M == 64) > 0;
}

int tc_pointwise_ysyn(const float* x, int act, const float* wc, int wc_is_kn, const float* bias, const float* Z,
                      const float* by, const float* mul, float* out, float* out_act, int64_t rows, int S2, int LY,
                      int K, int N, int mode, cudaStream_t st) {
    using namespace tc;
    PwArgs a;
    const int CI = (x && wc) ? K : 0;
    a.x = x; a.wc = wc; a.bias = bias; a.Z = Z; a.by = by; a.mul = mul; a.out = out; a.out_act = out_act;
    a.rows = rows; a.S2 = S2; a.LY = LY; a.CI = CI; a.CO = N; a.act = act; a.wc_is_kn = wc_is_kn;
    a.single_pass = (mode == FNM_MODE_TF32) ? 1 : 0;
    { const char* e = getenv("FNM_TC_DEBUG"); a.debug = e ? atoi(e) : 0; }
    a.M = N > 64 ? 128 : 64;
    a.K2p = (LY + 7) / 8 * 8;
    a.K2slabs = (a.K2p + 31) / 32;
    const bool a1_tmem = a.M == 128;
    a.NT = choose_nt(S2, CI, a.K2slabs, a.M, !a1_tmem);
    if (a.NT <= 0) return fail(FNM_ENOTSUP, "tc_pointwise_ysyn: no tile shape fits shared memory");
    a.tiles_per_row = (S2 + a.NT - 1) / a.NT;
    const PwSmem L = pw_layout(a.NT, CI, a.K2slabs, a.M, !a1_tmem);
    size_t smem = L.total + 1024;
    if (smem < 120 * 1024) smem = 120 * 1024;        // keep one CTA per SM: each CTA owns all 512 TMEM columns
    const int grid = (int)(rows < sm_count() ? rows : sm_count());
    cudaError_t e;
    if (a1_tmem) {
        e = cudaFuncSetAttribute(pointwise_ysyn_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(FNM_ECUDA, "tc_pointwise_ysyn: smem attribute: %s", cudaGetErrorString(e));
        LaunchScope ls("pointwise_ysyn.v1", st);
        pointwise_ysyn_tc<true><<<grid, kThreads, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(pointwise_ysyn_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(FNM_ECUDA, "tc_pointwise_ysyn: smem attribute: %s", cudaGetErrorString(e));
        LaunchScope ls("pointwise_ysyn.v1", st);
        pointwise_ysyn_tc<false><<<grid, kThreads, smem, st>>>(a);
    }
    return check_launch("tc_pointwise_ysyn");
}

}  // namespace fnm
