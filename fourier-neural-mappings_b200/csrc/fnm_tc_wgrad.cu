// tc_conv_wgrad -- gradient of the 1x1 "w" convolution on tcgen05 (3xTF32):
//
//     dWc[o][i] = sum_pos g[pos][o] * f(x[pos][i])        dbc[o] = sum_pos g[pos][o]
//
// Both operands are activation-sized and arrive as [pos][channel] (channel contiguous), so either
// becomes a K-major operand through ONE thread per channel that reads 32 consecutive positions with
// coalesced warp loads: the 32 registers are that channel's K-row.
//     A = g^T  : hi/lo TF32 split in registers, written to TMEM with tcgen05.st (lanes = o)
//     B = x^T  : row i of a SWIZZLE_128B K-major slab, written twice: raw and x_lo = x - trunc(x)
//     D[o][i] += g_hi*x_raw   (main)        Dc[o][i] += g_lo*x_raw + g_hi*x_lo   (corrections)
// The tensor core reads an fp32 shared-memory operand as TF32 by ignoring its low 13 mantissa bits,
// so the raw slab IS trunc(x) and no conversion pass over x is needed.  dbc falls out of the same
// registers (running sum per channel).
//
// Accuracy: the tensor core adds into its fp32 accumulator with truncation; measured on B200 the
// relative error grows by ~1.9e-8 per MMA accumulated, i.e. 1.3e-4 over the 6 700 MMAs a CTA would
// chain at config 5.  So (a) the correction passes, 2^-11 smaller, have their own accumulator, and
// (b) the main accumulator is double-buffered and FLUSHED every kCwSegChunks chunks: the epilogue
// warpgroup adds it (round-to-nearest FADD) into an fp32 accumulator in shared memory while the next
// segment accumulates into the other TMEM buffer.  Chains are 4*kCwSegChunks = 64 MMAs (1.2e-6).
//
// Every CTA owns a contiguous range of 32-position chunks and writes its partial D / db to the
// workspace; conv_wgrad_tc_reduce adds the partials up (deterministic).
//
// Warps: 0-7 two loader warpgroups (each thread holds the g AND x K-rows of its channel: 64 KB in
// flight per SM), 8-11 flush/epilogue warpgroup, 12 MMA issuer.  13 warps -> 128 registers.
#include "fnm_tc_ptx.cuh"

#include <cstdlib>

namespace fnm {
namespace tc {

constexpr int kCwLoaderWgs = 3;
constexpr int kCwStages = 3;                               // >= loader warpgroups (mbarrier parity safety)
constexpr int kCwSegChunks = 8;                            // 8 chunks x 12 MMAs per chain (main and corrections share it)
constexpr int kCwEpiWarp0 = 4 * kCwLoaderWgs;              // warps 12-15: flush / epilogue warpgroup
constexpr int kCwMmaWarp = kCwEpiWarp0 + 4;                // warp 16
constexpr int kCwThreads = (kCwMmaWarp + 4) * 32;          // 20 warps (5 per SM sub-partition): 96 registers each at launch.
// setmaxnreg re-balances them: the register pool a warpgroup can grow from holds only what OTHER warps of the
// same CTA released, so the MMA warp's three idle siblings (warps 17-19) exist to donate theirs:
//   released: epilogue 4 x 32 x (96-64) + MMA warp and its 3 donor siblings 4 x 32 x (96-32) = 12 288
//   claimed : loaders 12 x 32 x (120-96)                                              =  9 216
constexpr uint32_t kCwStageBytes = 32768;                  // raw slab 16 KB + lo slab 16 KB
constexpr uint32_t kCwAccBytes = 65536;                    // fp32 [128 cols][128 lanes]

struct CwArgs {
    const float* g; const float* x; float* part;
    long long npos;
    int Ci, Co, act, single_pass;
    long long nchunks;            // total 32-position chunks
    long long per;                // chunks per CTA
};

__device__ __forceinline__ float ld_stream_f(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// per-CTA partial layout (floats): D[Co][Ci], db[kCwLoaderWgs][Co]
__host__ __device__ inline size_t cw_part_floats(int Ci, int Co) { return (size_t)Ci * Co + (size_t)kCwLoaderWgs * Co; }

__global__ void __launch_bounds__(kCwThreads, 1) conv_wgrad_tc(const CwArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* acc = reinterpret_cast<float*>(smem + kCwStages * kCwStageBytes);            // [col i][lane o]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCwStages * kCwStageBytes + kCwAccBytes);
    uint64_t* full = bars;                         // [kCwStages]  the 4 warps of the loading warpgroup
    uint64_t* empty = bars + kCwStages;            // [kCwStages]
    uint64_t* seg_full = bars + 2 * kCwStages;     // [2]  main accumulator buffer ready to flush
    uint64_t* seg_empty = seg_full + 2;            // [2]  flushed (4 epilogue warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(seg_empty + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kCwStages; ++s) { mbar_init(&full[s], 4); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&seg_full[s], 1); mbar_init(&seg_empty[s], 4); }
        fence_barrier_init();
    }
    if (warp == kCwMmaWarp) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    // TMEM columns: [0,128) accumulator buffer 0, [128,256) buffer 1, [256,448) three g stages (hi 32 | lo 32)
    const uint32_t kStage0 = 256u;
    // register re-balancing: the loaders hold 64 loads in flight per thread, the other roles need few


    const long long c_begin = (long long)blockIdx.x * a.per;
    long long c_end = c_begin + a.per;
    if (c_end > a.nchunks) c_end = a.nchunks;
    const int n = c_end > c_begin ? (int)(c_end - c_begin) : 0;
    const int nseg = (n + kCwSegChunks - 1) / kCwSegChunks;
    float* part = a.part + (size_t)blockIdx.x * cw_part_floats(a.Ci, a.Co);

    if (warp > kCwMmaWarp) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;\n");      // register donors (same value as the MMA warp: one warpgroup), nothing else to do
    } else if (warp == kCwMmaWarp) {
        // ======================================================================= MMA issuer
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;\n");
        const uint32_t sbase = smem_u32(smem);
        const uint32_t idesc = make_idesc(128, a.Ci);
        for (int q = 0; q < n; ++q) {
            const int s = q % kCwStages;
            const int seg = q / kCwSegChunks, qs = q - seg * kCwSegChunks, db = seg & 1;
            if (qs == 0) {
                mbar_wait(&seg_empty[db], (uint32_t)(((seg >> 1) & 1) ^ 1));
                tc_fence_after();
            }
            mbar_wait(&full[s], (uint32_t)((q / kCwStages) & 1));
            tc_fence_after();
            const uint32_t dmain = tmem_base + 128u * db;
            const uint32_t g_hi = tmem_base + kStage0 + 64u * s, g_lo = g_hi + 32u;
            const uint64_t bx = make_desc(sbase + s * kCwStageBytes), bxl = make_desc(sbase + s * kCwStageBytes + 16384u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                umma_ts(dmain, g_hi + ks * 8, bx + 2 * ks, idesc, (qs > 0 || ks > 0) ? 1u : 0u);
                if (!a.single_pass) {
                    umma_ts(dmain, g_lo + ks * 8, bx + 2 * ks, idesc, 1);
                    umma_ts(dmain, g_hi + ks * 8, bxl + 2 * ks, idesc, 1);
                }
            }
            umma_commit(&empty[s]);
            if (qs == kCwSegChunks - 1 || q == n - 1) umma_commit(&seg_full[db]);
            __syncwarp();
        }
    } else if (warp < kCwEpiWarp0) {
        // ======================================================================= loaders: warpgroup wg takes chunks q = wg, wg+3, ...
        asm volatile("setmaxnreg.inc.sync.aligned.u32 120;\n");
        const int wg = warp >> 2;
        const int q4 = warp & 3;
        const int ch = q4 * 32 + lane;                             // channel = TMEM lane (g) = smem row (x)
        const bool g_ok = ch < a.Co, x_ok = ch < a.Ci;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16) + kStage0;
        float dbsum = 0.f;

        for (int q = wg; q < n; q += kCwLoaderWgs) {
            const int s = q % kCwStages;
            const long long pos0 = (c_begin + q) * 32;
            float vg[32], vx[32];
            {
                const float* gs = a.g + (size_t)pos0 * a.Co + ch;
                const float* xs = a.x + (size_t)pos0 * a.Ci + ch;
                const int lim = (int)((a.npos - pos0) < 32 ? (a.npos - pos0) : 32);
#pragma unroll
                for (int j = 0; j < 32; ++j) vg[j] = (g_ok && j < lim) ? ld_stream_f(gs + (size_t)j * a.Co) : 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) vx[j] = (x_ok && j < lim) ? ld_stream_f(xs + (size_t)j * a.Ci) : 0.f;
            }
            mbar_wait(&empty[s], (uint32_t)(((q / kCwStages) & 1) ^ 1));
            tc_fence_after();
            uint8_t* bx = smem + s * kCwStageBytes;
            const uint32_t tb = lane_base + 64u * s;
            // g: running bias-gradient sum, hi/lo split to TMEM
#pragma unroll
            for (int j = 0; j < 32; ++j) dbsum += vg[j];
            split_to_tmem16(tb, tb + 32u, vg, !a.single_pass);
            split_to_tmem16(tb + 16u, tb + 48u, vg + 16, !a.single_pass);
            // x: optional GELU, then the raw K-row and its truncation remainder
            if (a.act) {
#pragma unroll
                for (int j = 0; j < 32; ++j) vx[j] = gelu_fast(vx[j]);
            }
            if (x_ok) {
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const uint32_t off = swz(ch, c4);
                    const float4 v = make_float4(vx[4 * c4], vx[4 * c4 + 1], vx[4 * c4 + 2], vx[4 * c4 + 3]);
                    *reinterpret_cast<float4*>(bx + off) = v;
                    if (!a.single_pass) {
                        float4 l;
                        l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                        l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                        l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                        l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                        *reinterpret_cast<float4*>(bx + 16384 + off) = l;
                    }
                }
            }
            tmem_st_wait();
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        }
        if (g_ok) part[(size_t)a.Ci * a.Co + (size_t)wg * a.Co + ch] = dbsum;
    } else {
        // ======================================================================= flush / epilogue warpgroup
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;\n");
        const int q4 = warp & 3;
        const int o = q4 * 32 + lane;                              // TMEM lane = output channel
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
        for (int seg = 0; seg < nseg; ++seg) {
            const int db = seg & 1;
            mbar_wait(&seg_full[db], (uint32_t)((seg >> 1) & 1));
            tc_fence_after();
            const bool last = seg == nseg - 1;
            for (int c0 = 0; c0 < a.Ci; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(lane_addr + 128u * db + c0, r);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
                if (seg > 0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += acc[(c0 + j) * 128 + o];
                }
                if (!last) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[(c0 + j) * 128 + o] = v[j];
                } else {
                    if (o < a.Co) {
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4)
                            *reinterpret_cast<float4*>(part + (size_t)o * a.Ci + c0 + 4 * j4) =
                                make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&seg_empty[db]);
        }
        if (nseg == 0 && o < a.Co) {
            for (int c0 = 0; c0 < a.Ci; c0 += 4)
                *reinterpret_cast<float4*>(part + (size_t)o * a.Ci + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kCwMmaWarp) tmem_dealloc(tmem_base, 512);
}

// dwc / dbc = sum of the per-CTA partials.  fold > 1: the kernel ran on a folded problem -- `fold` consecutive
// positions of a C-channel tensor viewed as ONE position of a (fold*C)-channel tensor, which is the same
// memory -- so all 128 TMEM lanes and loader threads are busy for C = 32 / 64; the wanted C x C products are
// the `fold` diagonal blocks of the (fold*C) x (fold*C) result.
__global__ void conv_wgrad_tc_reduce(const float* __restrict__ part, int nparts, int Ci, int Co, int fold,
                                     float* __restrict__ dwc, float* __restrict__ dbc) {
    // one warp per output element; lanes stride over (partial, diagonal block) and shuffle-reduce.  (A coalesced
    // 32-outputs x 32-groups block was measured slower at config 1: 13 vs 9 us -- the partials are L2 resident and the
    // 1056 warps of this form spread over every SM.)
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int Cif = Ci * fold, Cof = Co * fold;
    const size_t stride = cw_part_floats(Cif, Cof);
    if (j >= Co * Ci + Co) return;
    float s = 0.f;
    if (j < Co * Ci) {
        const int o = j / Ci, i = j - o * Ci;
        for (int t = lane; t < nparts * fold; t += 32) {
            const int k = t / fold, r = t - k * fold;
            s += part[(size_t)k * stride + (size_t)(r * Co + o) * Cif + (r * Ci + i)];
        }
    } else {
        const int o = j - Co * Ci;
        for (int t = lane; t < nparts * fold * kCwLoaderWgs; t += 32) {
            const int k = t / (fold * kCwLoaderWgs), rem = t - k * fold * kCwLoaderWgs, w = rem / fold, r = rem - w * fold;
            s += part[(size_t)k * stride + (size_t)Cif * Cof + (size_t)w * Cof + r * Co + o];
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) {
        if (j < Co * Ci) { if (dwc) dwc[j] = s; }
        else if (dbc) dbc[j - Co * Ci] = s;
    }
}


// ------------------------------------------------------------------------------------------------------------------
// TMA variant for 128 (folded) channels without an activation on load.  Both operands are [pos][channel] with the
// channel contiguous, i.e. MN-major for a contraction over positions: g is the A operand (M = o), x the B operand
// (N = i), both straight from TMA boxes of 32 positions x 32 channels -- no register staging, no transposition.
// kind::tf32 accepts MN-major shared-memory operands only in the SWIZZLE_128B_BASE32B layout (128-byte rows whose 32-byte
// chunks are XORed with row & 3; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B); with the ordinary 128-byte swizzle the
// instruction returns zeros (tools/probes/mn_major_probe.cu).  Descriptor: LBO = distance of two 32-channel blocks,
// SBO = distance of two 4-row k groups (512 B), one k-step of 8 positions = 1024 B.
//   main  D += g_raw * x_raw        corrections  D += g_lo * x_raw + g_raw * x_lo   (lo tiles: converter warps, smem -> smem)
// x_lo sits right behind x_raw, so g_raw * [x_raw ; x_lo] is ONE N = 256 MMA whose 256 accumulator columns are
// (main | g_raw x_lo); g_lo * x_raw (N = 128) is added to the first half: two MMAs and two reads of the g slice per
// k-step instead of three.  The epilogue adds the two halves.
// The accumulator is flushed every kCwSegChunks chunks into REGISTERS of the 8 flush warps (64 columns each).
// Warps: 0 TMA producer | 1 MMA issuer | 2-5 converters (+ column sums of g for db) | 6-13 flush / epilogue.
constexpr int kW2Raw = 4, kW2Lo = 2;
constexpr uint32_t kW2Tile = 49152;                        // g: 4 channel blocks x [32 pos][128 B] | x: the same | x_lo
constexpr uint32_t kW2LoTile = 16384;                      // g_lo
constexpr int kW2Threads = 14 * 32;

__device__ __forceinline__ uint64_t desc_mn32(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)(4096 >> 4) << 16;                      // LBO: next 32-channel block
    d |= (uint64_t)(512 >> 4) << 32;                       // SBO: next group of 4 positions
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                                // SWIZZLE_128B_BASE32B
    return d;
}

__global__ void __launch_bounds__(kW2Threads, 1)
conv_wgrad_tma(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapX, const CwArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* lo_ring = smem + kW2Raw * kW2Tile;
    float* dbs = reinterpret_cast<float*>(lo_ring + kW2Lo * kW2LoTile);                    // [128] column sums of g
    uint64_t* bars = reinterpret_cast<uint64_t*>(dbs + 128);
    uint64_t* raw_full = bars;                     // [kW2Raw] TMA transaction barriers
    uint64_t* raw_empty = bars + kW2Raw;           // [kW2Raw] MMA commit + 4 converter warps (they read the tile for db in every mode)
    uint64_t* lo_full = bars + 2 * kW2Raw;         // [kW2Lo]  4 converter warps
    uint64_t* lo_empty = lo_full + kW2Lo;          // [kW2Lo]  MMA commit
    uint64_t* seg_full = lo_empty + kW2Lo;         // [2]
    uint64_t* seg_empty = seg_full + 2;            // [2] 8 flush warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(seg_empty + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kW2Raw; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 5); }
        for (int s = 0; s < kW2Lo; ++s) { mbar_init(&lo_full[s], 4); mbar_init(&lo_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&seg_full[s], 1); mbar_init(&seg_empty[s], 8); }
        fence_barrier_init();
        tma_prefetch_desc(&mapG); tma_prefetch_desc(&mapX);
    }
    if (threadIdx.x < 128) dbs[threadIdx.x] = 0.f;
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    const long long c0 = (long long)blockIdx.x * a.per;
    const long long c1 = c0 + a.per < a.nchunks ? c0 + a.per : a.nchunks;
    const uint32_t n = c1 > c0 ? (uint32_t)(c1 - c0) : 0u;
    const uint32_t nseg = (n + kCwSegChunks - 1) / kCwSegChunks;

    if (warp == 0) {
        // =========================================================================== TMA producer
        if (lane == 0) {
            for (uint32_t q = 0; q < n; ++q) {
                const uint32_t s = q % kW2Raw;
                mbar_wait(&raw_empty[s], ((q / kW2Raw) & 1) ^ 1);
                mbar_expect_tx(&raw_full[s], 32768u);
                uint8_t* st = smem + (size_t)s * kW2Tile;
                const int row = (int)((c0 + q) * 32);
                for (int j = 0; j < 4; ++j) {
                    tma_load_2d(&mapG, &raw_full[s], st + j * 4096, j * 32, row);
                    tma_load_2d(&mapX, &raw_full[s], st + 16384 + j * 4096, j * 32, row);
                }
            }
        }
    } else if (warp == 1) {
        // =========================================================================== MMA issuer
        const uint32_t sbase = smem_u32(smem), lbase = smem_u32(lo_ring);
        const uint32_t idesc128 = make_idesc(128, 128) | (1u << 15) | (1u << 16);       // A and B MN-major
        const uint32_t idesc256 = make_idesc(128, 256) | (1u << 15) | (1u << 16);
        for (uint32_t q = 0; q < n; ++q) {
            const uint32_t seg = q / kCwSegChunks, buf = seg & 1, first = (q % kCwSegChunks) == 0;
            if (first) mbar_wait(&seg_empty[buf], ((seg >> 1) & 1) ^ 1);
            const uint32_t s = q % kW2Raw, l = q % kW2Lo;
            const uint32_t d = tmem_base + buf * 256u;
            mbar_wait(&raw_full[s], (q / kW2Raw) & 1);
            if (!a.single_pass) mbar_wait(&lo_full[l], (q / kW2Lo) & 1);
            tc_fence_after();
            const uint32_t g_raw = sbase + s * kW2Tile, x_raw = g_raw + 16384u;
            const uint32_t g_lo = lbase + l * kW2LoTile;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)          // (main | g_raw x_lo): the x_lo blocks continue x_raw's at the same 4096-byte stride
                umma_ss(d, desc_mn32(g_raw + ks * 1024u), desc_mn32(x_raw + ks * 1024u), a.single_pass ? idesc128 : idesc256,
                        (first && ks == 0) ? 0u : 1u);
            if (!a.single_pass) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_ss(d, desc_mn32(g_lo + ks * 1024u), desc_mn32(x_raw + ks * 1024u), idesc128, 1u);
                umma_commit(&lo_empty[l]);
            }
            umma_commit(&raw_empty[s]);
            if ((q % kCwSegChunks) == kCwSegChunks - 1 || q == n - 1) umma_commit(&seg_full[buf]);
            __syncwarp();
        }
    } else if (warp < 6) {
        // =========================================================================== converters: lo = raw - trunc(raw), db
        const int ct = (int)threadIdx.x - 64;                        // 0..127
        const int pair = ct & 31, rgrp = ct >> 5;                    // (channel block j, 16-byte chunk c) and row group
        const int j = pair >> 3, c = pair & 7;
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t q = 0; q < n; ++q) {
            const uint32_t s = q % kW2Raw, l = q % kW2Lo;
            mbar_wait(&raw_full[s], (q / kW2Raw) & 1);
            const uint32_t src0 = smem_u32(smem) + s * kW2Tile, dst0 = smem_u32(lo_ring) + l * kW2LoTile;
            if (!a.single_pass) mbar_wait(&lo_empty[l], ((q / kW2Lo) & 1) ^ 1);
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                const uint32_t r = (uint32_t)(rgrp + 4 * rr);
                const uint32_t off = (uint32_t)j * 4096u + r * 128u + ((((uint32_t)c >> 1) ^ (r & 3u)) << 5) + ((uint32_t)c & 1u) * 16u;
                float4 vg, vx;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vg.x), "=f"(vg.y), "=f"(vg.z), "=f"(vg.w) : "r"(src0 + off));
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vx.x), "=f"(vx.y), "=f"(vx.z), "=f"(vx.w) : "r"(src0 + 16384u + off));
                sum.x += vg.x; sum.y += vg.y; sum.z += vg.z; sum.w += vg.w;
                if (!a.single_pass) {
                    vg.x -= __uint_as_float(__float_as_uint(vg.x) & 0xffffe000u); vg.y -= __uint_as_float(__float_as_uint(vg.y) & 0xffffe000u);
                    vg.z -= __uint_as_float(__float_as_uint(vg.z) & 0xffffe000u); vg.w -= __uint_as_float(__float_as_uint(vg.w) & 0xffffe000u);
                    vx.x -= __uint_as_float(__float_as_uint(vx.x) & 0xffffe000u); vx.y -= __uint_as_float(__float_as_uint(vx.y) & 0xffffe000u);
                    vx.z -= __uint_as_float(__float_as_uint(vx.z) & 0xffffe000u); vx.w -= __uint_as_float(__float_as_uint(vx.w) & 0xffffe000u);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst0 + off), "f"(vg.x), "f"(vg.y), "f"(vg.z), "f"(vg.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(src0 + 32768u + off), "f"(vx.x), "f"(vx.y), "f"(vx.z), "f"(vx.w) : "memory");
                }
            }
            if (!a.single_pass) fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                if (!a.single_pass) mbar_arrive(&lo_full[l]);
                mbar_arrive(&raw_empty[s]);
            }
        }
        // bias gradient: channel 32 j + 4 c + {0..3}; the four row groups meet in shared memory
        atomicAdd(&dbs[32 * j + 4 * c + 0], sum.x); atomicAdd(&dbs[32 * j + 4 * c + 1], sum.y);
        atomicAdd(&dbs[32 * j + 4 * c + 2], sum.z); atomicAdd(&dbs[32 * j + 4 * c + 3], sum.w);
    } else {
        // =========================================================================== flush / epilogue (8 warps)
        const int ew = warp - 6, q4 = warp & 3, half = ew >> 2;      // TMEM lane quarter = warp % 4
        const int o = q4 * 32 + lane;
        float acc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = 0.f;
        for (uint32_t seg = 0; seg < nseg; ++seg) {
            const uint32_t buf = seg & 1;
            mbar_wait(&seg_full[buf], (seg >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + buf * 256u + (uint32_t)half * 64u;
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
                uint32_t r[16], r2[16];
                tmem_ld16(taddr + g4 * 16, r);
                if (!a.single_pass) tmem_ld16(taddr + 128u + g4 * 16, r2);          // the g_raw * x_lo half
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    acc[g4 * 16 + i] += __uint_as_float(r[i]) + (a.single_pass ? 0.f : __uint_as_float(r2[i]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&seg_empty[buf]);
        }
        float* pd = a.part + (size_t)blockIdx.x * cw_part_floats(a.Ci, a.Co) + (size_t)o * a.Ci + half * 64;
#pragma unroll
        for (int i = 0; i < 64; i += 4) *reinterpret_cast<float4*>(pd + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 128) {
        float* pb = a.part + (size_t)blockIdx.x * cw_part_floats(a.Ci, a.Co) + (size_t)a.Ci * a.Co;
        pb[threadIdx.x] = dbs[threadIdx.x];
        for (int w = 1; w < kCwLoaderWgs; ++w) pb[w * a.Co + threadIdx.x] = 0.f;
    }
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// gz_pass_tma: the two passes of a layer's backward that each read the incoming gradient g in full, as ONE read of g
// ("read once"; models/func_to_func2d.py:89-95 backward):
//     T[row][l][c] = sum_y tab[l][y] g[row][y][c]                                   (ydft_tma's product, fnm_tc_ydft.cu)
//     dWc[o][i]    = sum_pos g[pos][o] x[pos][i],   dbc[o] = sum_pos g[pos][o]      (conv_wgrad_tma's products, above)
// Both contract over positions with g^T as the A operand (M = channel of g = TMEM lane), so the MN-major TMA tile of g and
// its remainder tile serve both; the B operands are the stacked [x_raw ; x_lo] tile (MN-major, N = 256) and the stacked
// [tab ; tab_lo] chunk (K-major, N = 128).  Per 32-position chunk and k-step:
//     Dy(main | corr) += g_raw * [tab ; tab_lo]      Dy(corr) += g_lo * tab
//     Dc(main | corr) += g_raw * [x_raw ; x_lo]      Dc(corr) += g_lo * x_raw
// A work item is one grid row: its ydft accumulator (128 columns, double buffered) goes to T, the conv accumulator (256
// columns, single buffered) is flushed into the registers of the 8 epilogue warps after every row -- chains of 32 (main) and
// 64 (corrections) MMAs at 256 positions per row.  The next row's first ydft MMAs are issued before the issuer waits for
// that flush, so the tensor pipe has work while the epilogue drains the conv accumulator.
// C = 32 / 64: F = 128 / C consecutive grid rows are folded into the 128 lanes as in ydft_tma (the 32-channel blocks of the g
// and x tiles come from F different rows); the conv product then holds the wanted C x C blocks on its diagonal, which is
// the folded partial format conv_wgrad_tc_reduce already sums.
// Stage: g 16 KB | x 16 KB | x_lo 16 KB | tab 8 KB | tab_lo 8 KB; g_lo in its own ring.
// Warps: 0 TMA producer | 1 MMA issuer | 2-5 converters (+ column sums of g for dbc) | 6-13 epilogue.
constexpr int kGzRaw = 3, kGzLo = 2;
constexpr uint32_t kGzStage = 65536, kGzXOff = 16384, kGzTabOff = 49152;
constexpr uint32_t kGzLoTile = 16384;
constexpr int kGzThreads = 14 * 32;

struct GzArgs {
    float* T; float* part;
    long long rows, groups;               // grid rows; row groups of F rows (the work items)
    int S2, LY, nch, single_pass;
    int C, F, cshift;                     // channels (= 1 << cshift); C = 32 / 64: F = 128 / C consecutive grid rows share the 128 lanes
};

__global__ void __launch_bounds__(kGzThreads, 1)
gz_pass_tma(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapX,
            const __grid_constant__ CUtensorMap mapT, const GzArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* lo_ring = smem + kGzRaw * kGzStage;
    float* dbs = reinterpret_cast<float*>(lo_ring + kGzLo * kGzLoTile);                    // [128] column sums of g
    uint64_t* bars = reinterpret_cast<uint64_t*>(dbs + 128);
    uint64_t* raw_full = bars;                     // [kGzRaw] TMA transaction barriers
    uint64_t* raw_empty = bars + kGzRaw;           // [kGzRaw] MMA commit + 4 converter warps
    uint64_t* lo_full = bars + 2 * kGzRaw;         // [kGzLo]  4 converter warps
    uint64_t* lo_empty = lo_full + kGzLo;          // [kGzLo]  MMA commit
    uint64_t* cv_full = lo_empty + kGzLo;          // [1] conv accumulator of a row complete
    uint64_t* cv_empty = cv_full + 1;              // [1] 8 epilogue warps
    uint64_t* yd_full = cv_empty + 1;              // [2]
    uint64_t* yd_empty = yd_full + 2;              // [2] 8 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(yd_empty + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kGzRaw; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 5); }
        for (int s = 0; s < kGzLo; ++s) { mbar_init(&lo_full[s], 4); mbar_init(&lo_empty[s], 1); }
        mbar_init(cv_full, 1); mbar_init(cv_empty, 8);
        for (int s = 0; s < 2; ++s) { mbar_init(&yd_full[s], 1); mbar_init(&yd_empty[s], 8); }
        fence_barrier_init();
        tma_prefetch_desc(&mapG); tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapT);
    }
    if (threadIdx.x < 128) dbs[threadIdx.x] = 0.f;
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    const uint32_t my_rows = (uint32_t)((a.groups - (long long)blockIdx.x + (long long)gridDim.x - 1) / (long long)gridDim.x);
    const uint32_t nch = (uint32_t)a.nch;
    const uint32_t total = my_rows * nch;

    if (warp == 0) {
        // =========================================================================== TMA producer
        if (lane == 0) {
            uint32_t it = 0, ch = 0;
            for (uint32_t q = 0; q < total; ++q) {
                const uint32_t s = q % kGzRaw;
                const int row0 = (int)(blockIdx.x + it * gridDim.x) * a.F;
                const int bpr = a.C / 32;                          // 32-channel blocks per grid row
                mbar_wait(&raw_empty[s], ((q / kGzRaw) & 1) ^ 1);
                mbar_expect_tx(&raw_full[s], 2u * 16384u + 8192u);
                uint8_t* st = smem + (size_t)s * kGzStage;
                for (int j = 0; j < 4; ++j) {                      // positions / rows past the end read as zeros (TMA bounds)
                    tma_load_3d(&mapG, &raw_full[s], st + j * 4096, (j % bpr) * 32, (int)ch * 32, row0 + j / bpr);
                    tma_load_3d(&mapX, &raw_full[s], st + kGzXOff + j * 4096, (j % bpr) * 32, (int)ch * 32, row0 + j / bpr);
                }
                tma_load_2d(&mapT, &raw_full[s], st + kGzTabOff, (int)ch * 32, 0);
                if (++ch == nch) { ch = 0; ++it; }
            }
        }
    } else if (warp == 1) {
        // =========================================================================== MMA issuer
        const uint32_t sbase = smem_u32(smem), lbase = smem_u32(lo_ring);
        const uint32_t idc128 = make_idesc(128, 128) | (1u << 15) | (1u << 16);       // conv: A and B MN-major
        const uint32_t idc256 = make_idesc(128, 256) | (1u << 15) | (1u << 16);
        const uint32_t idy64 = make_idesc(128, 64) | (1u << 15);                      // ydft: A MN-major, B K-major
        const uint32_t idy128 = make_idesc(128, 128) | (1u << 15);
        uint32_t q = 0;
        for (uint32_t it = 0; it < my_rows; ++it) {
            const uint32_t ybuf = it & 1u;
            const uint32_t dy = tmem_base + 256u + ybuf * 128u;
            for (uint32_t ch = 0; ch < nch; ++ch, ++q) {
                const uint32_t s = q % kGzRaw, l = q % kGzLo;
                mbar_wait(&raw_full[s], (q / kGzRaw) & 1);
                if (!a.single_pass) mbar_wait(&lo_full[l], (q / kGzLo) & 1);
                if (ch == 0) mbar_wait(&yd_empty[ybuf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t g_raw = sbase + s * kGzStage, x_raw = g_raw + kGzXOff, tabb = g_raw + kGzTabOff;
                const uint32_t g_lo = lbase + l * kGzLoTile;
                const uint32_t first = ch == 0 ? 0u : 1u;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)          // (main | corr) of the row's DFT: g_raw * [tab ; tab_lo]
                    umma_ss(dy, desc_mn32(g_raw + ks * 1024u), make_desc(tabb + ks * 32u), a.single_pass ? idy64 : idy128,
                            ks == 0 ? first : 1u);
                if (!a.single_pass) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_ss(dy + 64u, desc_mn32(g_lo + ks * 1024u), make_desc(tabb + ks * 32u), idy64, 1u);
                }
                if (ch == 0) {                          // the previous row's conv accumulator must have been flushed
                    mbar_wait(cv_empty, (it & 1) ^ 1);
                    tc_fence_after();
                }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)          // (main | g_raw x_lo)
                    umma_ss(tmem_base, desc_mn32(g_raw + ks * 1024u), desc_mn32(x_raw + ks * 1024u), a.single_pass ? idc128 : idc256,
                            ks == 0 ? first : 1u);
                if (!a.single_pass) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_ss(tmem_base + 128u, desc_mn32(g_lo + ks * 1024u), desc_mn32(x_raw + ks * 1024u), idc128, 1u);
                    umma_commit(&lo_empty[l]);
                }
                umma_commit(&raw_empty[s]);
                if (ch == nch - 1) { umma_commit(cv_full); umma_commit(&yd_full[ybuf]); }
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        // =========================================================================== converters: lo = raw - trunc(raw), db
        const int ct = (int)threadIdx.x - 64;                        // 0..127
        const int pair = ct & 31, rgrp = ct >> 5;                    // (channel block j, 16-byte chunk c) and row group
        const int j = pair >> 3, c = pair & 7;
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t q = 0; q < total; ++q) {
            const uint32_t s = q % kGzRaw, l = q % kGzLo;
            mbar_wait(&raw_full[s], (q / kGzRaw) & 1);
            const uint32_t src0 = smem_u32(smem) + s * kGzStage, dst0 = smem_u32(lo_ring) + l * kGzLoTile;
            if (!a.single_pass) mbar_wait(&lo_empty[l], ((q / kGzLo) & 1) ^ 1);
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                const uint32_t r = (uint32_t)(rgrp + 4 * rr);
                const uint32_t off = (uint32_t)j * 4096u + r * 128u + ((((uint32_t)c >> 1) ^ (r & 3u)) << 5) + ((uint32_t)c & 1u) * 16u;
                float4 vg, vx;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vg.x), "=f"(vg.y), "=f"(vg.z), "=f"(vg.w) : "r"(src0 + off));
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vx.x), "=f"(vx.y), "=f"(vx.z), "=f"(vx.w) : "r"(src0 + kGzXOff + off));
                sum.x += vg.x; sum.y += vg.y; sum.z += vg.z; sum.w += vg.w;
                if (!a.single_pass) {
                    vg.x -= __uint_as_float(__float_as_uint(vg.x) & 0xffffe000u); vg.y -= __uint_as_float(__float_as_uint(vg.y) & 0xffffe000u);
                    vg.z -= __uint_as_float(__float_as_uint(vg.z) & 0xffffe000u); vg.w -= __uint_as_float(__float_as_uint(vg.w) & 0xffffe000u);
                    vx.x -= __uint_as_float(__float_as_uint(vx.x) & 0xffffe000u); vx.y -= __uint_as_float(__float_as_uint(vx.y) & 0xffffe000u);
                    vx.z -= __uint_as_float(__float_as_uint(vx.z) & 0xffffe000u); vx.w -= __uint_as_float(__float_as_uint(vx.w) & 0xffffe000u);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst0 + off), "f"(vg.x), "f"(vg.y), "f"(vg.z), "f"(vg.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(src0 + 2u * kGzXOff + off), "f"(vx.x), "f"(vx.y), "f"(vx.z), "f"(vx.w) : "memory");
                }
            }
            if (!a.single_pass) {
                // the table chunk into the second half of its stacked slab (layout-preserving copy, 2 KB per pass)
                const uint32_t tsrc = src0 + kGzTabOff + (uint32_t)ct * 16u;
#pragma unroll
                for (uint32_t off = 0; off < 8192u; off += 2048u) {
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(tsrc + off));
                    v.x -= __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); v.y -= __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                    v.z -= __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); v.w -= __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tsrc + off + 8192u), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
                fence_proxy_async();
            }
            __syncwarp();
            if (lane == 0) {
                if (!a.single_pass) mbar_arrive(&lo_full[l]);
                mbar_arrive(&raw_empty[s]);
            }
        }
        // bias gradient: channel 32 j + 4 c + {0..3}; the four row groups meet in shared memory
        atomicAdd(&dbs[32 * j + 4 * c + 0], sum.x); atomicAdd(&dbs[32 * j + 4 * c + 1], sum.y);
        atomicAdd(&dbs[32 * j + 4 * c + 2], sum.z); atomicAdd(&dbs[32 * j + 4 * c + 3], sum.w);
    } else {
        // =========================================================================== epilogue (8 warps)
        const int ew = warp - 6, q4 = warp & 3, half = ew >> 2;      // TMEM lane quarter = warp % 4
        const int o = q4 * 32 + lane;
        const int fr = o >> a.cshift, cc = o & (a.C - 1);            // lane = (row within the group, channel)
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
        float acc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = 0.f;
        for (uint32_t it = 0; it < my_rows; ++it) {
            const long long row = ((long long)blockIdx.x + (long long)it * gridDim.x) * a.F + fr;
            const uint32_t ybuf = it & 1u;
            // conv accumulator of the row -> registers (the issuer waits for this before the next row's conv MMAs)
            mbar_wait(cv_full, it & 1);
            tc_fence_after();
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
                uint32_t r[16], r2[16];
                tmem_ld16(lane_base + (uint32_t)half * 64u + g4 * 16, r);
                if (!a.single_pass) tmem_ld16(lane_base + 128u + (uint32_t)half * 64u + g4 * 16, r2);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    acc[g4 * 16 + i] += __uint_as_float(r[i]) + (a.single_pass ? 0.f : __uint_as_float(r2[i]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(cv_empty);
            // the row's transform: 32 of the 64 table rows per warp, 128-byte coalesced stores (lane = channel)
            mbar_wait(&yd_full[ybuf], (it >> 1) & 1);
            tc_fence_after();
            const int l0 = half * 32;
            float* op = a.T + ((((size_t)row * a.LY + l0) << a.cshift) + cc);
            const int lmax = row < a.rows ? a.LY - l0 : 0;           // valid table rows of this warp's half
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t m0[16], m1[16];
                const uint32_t ty = lane_base + 256u + ybuf * 128u + (uint32_t)(l0 + h * 16);
                tmem_ld16(ty, m0);
                if (!a.single_pass) tmem_ld16(ty + 64u, m1);
                tmem_ld_wait();
                if (h == 1) {                                        // both halves read: the buffer is free for the row after next
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&yd_empty[ybuf]);
                }
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (h * 16 + i < lmax)
                        op[(size_t)(h * 16 + i) << a.cshift] = __uint_as_float(m0[i]) + (a.single_pass ? 0.f : __uint_as_float(m1[i]));
            }
        }
        float* pd = a.part + (size_t)blockIdx.x * cw_part_floats(128, 128) + (size_t)o * 128 + half * 64;
#pragma unroll
        for (int i = 0; i < 64; i += 4) *reinterpret_cast<float4*>(pd + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 128) {
        float* pb = a.part + (size_t)blockIdx.x * cw_part_floats(128, 128) + (size_t)128 * 128;
        pb[threadIdx.x] = dbs[threadIdx.x];
        for (int w = 1; w < kCwLoaderWgs; ++w) pb[w * 128 + threadIdx.x] = 0.f;
    }
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static int cw_grid(int64_t npos, bool for_workspace = false) {
    const int64_t chunks = (npos + 31) / 32;
    const int sms = for_workspace ? sm_count_max() : sm_count();     // the workspace must hold any later SM budget
    // at least 16 chunks per CTA: on the small L2-resident problems every extra CTA is one more 66 KB partial for the
    // reduction kernel, while the main kernel's time is its prologue either way
    const int64_t want = (chunks + 15) / 16;
    return (int)(want < 1 ? 1 : (want < sms ? want : sms));
}

}  // namespace tc

using namespace tc;

bool tc_conv_wgrad_supported(int64_t npos, int Ci, int Co) {
    if (!tc_enabled()) return false;
    return npos >= 4096 && Ci % 16 == 0 && Co % 16 == 0 && Ci >= 16 && Co >= 16 && Ci <= 128 && Co <= 128;
}

static int cw_fold(int64_t npos, int Ci, int Co) {
    if (Ci != Co || (Ci != 32 && Ci != 64)) return 1;
    const int fold = 128 / Ci;
    return npos % fold == 0 ? fold : 1;
}

size_t tc_conv_wgrad_workspace(int64_t npos, int Ci, int Co) {
    const int fold = cw_fold(npos, Ci, Co);
    return align_up((size_t)cw_grid(npos / fold, true) * cw_part_floats(Ci * fold, Co * fold) * sizeof(float), 256);
}

int tc_conv_wgrad(const float* g, const float* x, int act, float* dwc, float* dbc, int64_t npos, int Ci, int Co,
                  float* partial, int mode, cudaStream_t st) {
    CwArgs a{};
    const int fold = cw_fold(npos, Ci, Co);
    a.g = g; a.x = x; a.part = partial; a.npos = npos / fold; a.Ci = Ci * fold; a.Co = Co * fold; a.act = act;
    a.single_pass = mode == FNM_MODE_TF32;
    a.nchunks = (a.npos + 31) / 32;
    const int grid = cw_grid(a.npos);
    a.per = (a.nchunks + grid - 1) / grid;
    // 128 (folded) channels, no activation on load, aligned: operands straight from TMA (MN-major), no register staging
    {
        const char* e = getenv("FNM_WGRAD_NO_TMA");
        const bool off = e && e[0] == '1';
        if (!off && act == 0 && a.Ci == 128 && a.Co == 128 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(x)) & 15) == 0) {
            CUtensorMap mg, mx;
            const cuuint64_t dims[2] = {128, (cuuint64_t)a.npos};
            const cuuint64_t str[1] = {128 * 4};
            const cuuint32_t box[2] = {32, 32};
            if (tc_make_map_swz(&mg, g, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) &&
                tc_make_map_swz(&mx, x, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) {
                static std::atomic<uint64_t> attr2{0};
    int attr2_dev = 0;
                const size_t smem2 = (size_t)kW2Raw * kW2Tile + (size_t)kW2Lo * kW2LoTile + 128 * 4 + 1024 + 512;
                if (once_needed(attr2, attr2_dev)) {
                    const cudaError_t e2 = cudaFuncSetAttribute(conv_wgrad_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
                    if (e2 != cudaSuccess) return fail(FNM_ECUDA, "conv_wgrad: smem attribute: %s", cudaGetErrorString(e2));
                    once_done(attr2, attr2_dev);
                }
                {
                    LaunchScope ls("conv_wgrad", st);
                    conv_wgrad_tma<<<grid, kW2Threads, smem2, st>>>(mg, mx, a);
                }
                FNM_TRY(check_launch("conv_wgrad"));
                const int count2 = Co * Ci + Co;
                {
                    LaunchScope ls("conv_wgrad_reduce", st);
                    conv_wgrad_tc_reduce<<<(count2 * 32 + 255) / 256, 256, 0, st>>>(partial, grid, Ci, Co, fold, dwc, dbc);
                }
                return check_launch("conv_wgrad_reduce");
            }
        }
    }
    static std::atomic<uint64_t> attr_set{0};
    int attr_set_dev = 0;
    const size_t smem = kCwStages * kCwStageBytes + kCwAccBytes + 2048;   // > half the SM: one CTA per SM owns the TMEM
    if (once_needed(attr_set, attr_set_dev)) {
        const cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(FNM_ECUDA, "conv_wgrad: smem attribute: %s", cudaGetErrorString(e));
        once_done(attr_set, attr_set_dev);
    }
    {
        LaunchScope ls("conv_wgrad", st);
        conv_wgrad_tc<<<grid, kCwThreads, smem, st>>>(a);
    }
    FNM_TRY(check_launch("conv_wgrad"));
    const int count = Co * Ci + Co;
    {
        LaunchScope ls("conv_wgrad_reduce", st);
        conv_wgrad_tc_reduce<<<(count * 32 + 255) / 256, 256, 0, st>>>(partial, grid, Ci, Co, fold, dwc, dbc);
    }
    return check_launch("conv_wgrad_reduce");
}


// fused ydft(g) + conv weight gradient (gz_pass_tma): 128 channels on both sides (or 32 / 64 with grid rows folded into the
// lanes), no activation on load, whole row groups as work items -- problems with few long rows (1-D) keep the separate,
// row-segmenting kernels.  The grid never exceeds the partial count tc_conv_wgrad_workspace sized the workspace for.
bool tc_gz_pass_supported(const float* g, const float* x, const float* tab, int ld, int64_t rows, int S2, int LY, int Ci, int Co, int act) {
    // ld: row stride of the table in floats (S2 itself, or the padded copy fnm_plan.byT4 when S2 is no multiple of 4)
    if (!tc_enabled() || act != 0 || Ci != Co || (Ci != 128 && Ci != 64 && Ci != 32) || LY < 8 || LY > 64 || ld % 4 != 0 || ld < S2 || S2 < 4)
        return false;
    const int F = 128 / Ci;
    // (small L2-resident problems are launch bound either way and do better with every SM on the separate kernels)
    if (rows >= (1LL << 30) || (rows + F - 1) / F < 4LL * sm_count() || (rows * S2) % F != 0) return false;
    if ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(tab)) & 15) return false;
    const char* e = getenv("FNM_NO_GZPASS");
    return !(e && e[0] == '1');
}

int tc_gz_pass(const float* g, const float* x, const float* tab, int ld, float* T, float* dwc, float* dbc, int64_t rows, int S2, int LY,
               int C, float* partial, int mode, cudaStream_t st) {
    GzArgs a{};
    a.T = T; a.part = partial; a.rows = rows; a.S2 = S2; a.LY = LY; a.nch = (S2 + 31) / 32;
    a.single_pass = mode == FNM_MODE_TF32;
    a.C = C; a.F = 128 / C; a.cshift = C == 128 ? 7 : (C == 64 ? 6 : 5); a.groups = (rows + a.F - 1) / a.F;
    CUtensorMap mg, mx, mt;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)S2, (cuuint64_t)rows};
        const cuuint64_t str[2] = {(cuuint64_t)C * 4, (cuuint64_t)S2 * C * 4};
        const cuuint32_t box[3] = {32, 32, 1};
        if (!tc_make_map_swz(&mg, g, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
            !tc_make_map_swz(&mx, x, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
            return fail(FNM_ECUDA, "ydft_conv_wgrad: tensor map (g / x) failed");
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)S2, (cuuint64_t)LY};
        const cuuint64_t str[1] = {(cuuint64_t)ld * 4};
        const cuuint32_t box[2] = {32, 64};
        if (!tc_make_map(&mt, tab, 2, dims, str, box, true)) return fail(FNM_ECUDA, "ydft_conv_wgrad: tensor map (table) failed");
    }
    const size_t smem = (size_t)kGzRaw * kGzStage + (size_t)kGzLo * kGzLoTile + 128 * 4 + 1024 + 256;
    static std::atomic<uint64_t> attr_set{0};
    int attr_set_dev = 0;
    if (once_needed(attr_set, attr_set_dev)) {
        const cudaError_t e = cudaFuncSetAttribute(gz_pass_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(FNM_ECUDA, "ydft_conv_wgrad: smem attribute: %s", cudaGetErrorString(e));
        once_done(attr_set, attr_set_dev);
    }
    // as many CTAs as conv_wgrad would use on the same (folded) problem: the workspace holds that many partials, and the small
    // L2-resident problems get few partials for the reduction kernel
    const int cap = cw_grid(rows * S2 / a.F);
    const int grid = a.groups < cap ? (int)a.groups : cap;
    {
        LaunchScope ls("ydft_conv_wgrad", st);
        gz_pass_tma<<<grid, kGzThreads, smem, st>>>(mg, mx, mt, a);
    }
    FNM_TRY(check_launch("ydft_conv_wgrad"));
    {
        LaunchScope ls("conv_wgrad_reduce", st);
        conv_wgrad_tc_reduce<<<((C * C + C) * 32 + 255) / 256, 256, 0, st>>>(partial, grid, C, C, a.F, dwc, dbc);
    }
    return check_launch("conv_wgrad_reduce");
}

}  // namespace fnm
