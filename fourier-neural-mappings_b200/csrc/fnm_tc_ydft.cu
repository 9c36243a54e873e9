// tc_ydft_tma -- the truncated DFT along y of a channels-last activation with TMA-fed operands:
//
//     T[row][l'][c] = sum_y tab[l'][y] * x[row][y][c]          row = (sample, x) grid row, l' < LY <= 64, c < C
//
// C = 128, or C = 32 / 64 with F = 128 / C consecutive grid rows FOLDED into the 128 TMEM lanes (lane = (row in group,
// channel): the 32-channel blocks of the A tile are simply TMA-loaded from F different rows; everything below says "row" for
// such a row group).
//
// x[row] is [y][c] with c contiguous: MN-major for the contraction over y, so it is the A operand (M = c, lanes = c)
// straight from TMA boxes of 32 y x 32 c in the SWIZZLE_128B_BASE32B layout (see fnm_tc_wgrad.cu); the table chunk
// [l'][32 y] is the K-major B operand, also by TMA.  Nothing is staged through registers.  Converter warps derive the
// remainder tiles (x_lo, tab_lo) shared memory -> shared memory; 3xTF32:
//     main  D_m += x_raw * tab_raw          corrections  D_c += x_lo * tab_raw + x_raw * tab_lo
// The table chunk and its remainder are stacked into ONE 128-row operand [tab_raw ; tab_lo], so x_raw * [tab_raw ; tab_lo]
// is a single N = 128 MMA that feeds the main and the correction accumulator (adjacent TMEM columns) at once: two MMAs
// and two reads of the activation slice per k-step instead of three.
// The table chunk is streamed (it is L2 resident) instead of kept in shared memory, which leaves room for a deep raw ring;
// an item is a PAIR of grid rows so that each streamed table chunk serves two activation tiles.
// Accuracy as in fnm_tc_tabgemm.cu: separate correction accumulator, chunks alternate between two main accumulators.
// TMEM (per row of the pair): main0 | corr0 | main1 | corr1, 64 columns each (even / odd chunks).
// Warps: 0 TMA producer | 1 MMA issuer | 2-5 converters | 6-13 epilogue (lane = channel, 128-byte coalesced stores).
#include "fnm_tc_ptx.cuh"

#include <cstdlib>

namespace fnm {
namespace tc {

constexpr int kYdRaw = 3, kYdLo = 2;
constexpr uint32_t kYdATile = 16384;                         // one row: 4 channel blocks x [32 y][128 B]
constexpr uint32_t kYdStage = 2 * kYdATile + 16384;          // two rows + [table chunk ; its remainder] = [128 rows][128 B]
constexpr uint32_t kYdLoStage = 2 * kYdATile;                // remainder tiles of the two rows
constexpr int kYdThreads = 14 * 32;

struct YdArgs {
    float* out;
    long long rows, groups;                 // grid rows; row groups (F rows each)
    int P2, LY, NT, nch, single_pass;
    int C, F, gpi;                          // channels, rows folded per group, groups per work item (2: one table chunk serves both)
    long long pairs;                        // work items
    // K segments (1-D models: few, long rows): an item is (segment, row pair) and covers `nch` chunks starting at chunk
    // seg * nch; segment s writes its PARTIAL transform to out + s * seg_stride (the consumer sums them).  nseg = 1: none.
    int nseg;
    long long pairs0, seg_stride;           // row pairs per segment; floats between the partial outputs
};

__device__ __forceinline__ uint64_t yd_desc_mn32(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)(4096 >> 4) << 16;                      // LBO: next 32-channel block
    d |= (uint64_t)(512 >> 4) << 32;                       // SBO: next group of 4 y
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                                // SWIZZLE_128B_BASE32B
    return d;
}

__global__ void __launch_bounds__(kYdThreads, 1)
ydft_tma(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapT, const YdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* lo_ring = smem + kYdRaw * kYdStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(lo_ring + kYdLo * kYdLoStage);
    uint64_t* raw_full = bars;                     // [kYdRaw] TMA transaction barriers
    uint64_t* raw_empty = bars + kYdRaw;           // [kYdRaw] MMA commit + 4 converter warps
    uint64_t* lo_full = bars + 2 * kYdRaw;         // [kYdLo] 4 converter warps
    uint64_t* lo_empty = lo_full + kYdLo;          // [kYdLo] MMA commit
    uint64_t* d_full = lo_empty + kYdLo;           // [1]
    uint64_t* d_empty = d_full + 1;                // [1] 8 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 1);

    const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int nconv = a.single_pass ? 0 : 4;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kYdRaw; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 1 + nconv); }
        for (int s = 0; s < kYdLo; ++s) { mbar_init(&lo_full[s], 4); mbar_init(&lo_empty[s], 1); }
        mbar_init(d_full, 1); mbar_init(d_empty, 8);
        fence_barrier_init();
        tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapT);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    const long long my_pairs = (a.pairs - (long long)blockIdx.x + (long long)gridDim.x - 1) / (long long)gridDim.x;
    const uint32_t nch = (uint32_t)a.nch;
    const uint32_t total = (uint32_t)my_pairs * nch;

    if (warp == 0) {
        // =========================================================================== TMA producer
        if (lane == 0) {
            for (uint32_t q = 0; q < total; ++q) {
                const uint32_t it = q / nch, s = q % kYdRaw;
                uint32_t ch = q - it * nch;
                long long pair = (long long)blockIdx.x + (long long)it * gridDim.x;
                if (a.nseg > 1) { const long long seg = pair / a.pairs0; pair -= seg * a.pairs0; ch += (uint32_t)seg * nch; }
                mbar_wait(&raw_empty[s], ((q / kYdRaw) & 1) ^ 1);
                const bool two = a.gpi == 2 && pair * 2 + 1 < a.groups;
                mbar_expect_tx(&raw_full[s], (two ? 2u : 1u) * kYdATile + 8192u);
                uint8_t* st = smem + (size_t)s * kYdStage;
                const int bpr = a.C / 32;                                // 32-channel blocks per grid row
                for (int j = 0; j < (two ? 2 : 1); ++j)
                    for (int cb = 0; cb < 4; ++cb)                       // rows past the end read as zeros (TMA bounds)
                        tma_load_3d(&mapX, &raw_full[s], st + j * kYdATile + cb * 4096, (cb % bpr) * 32, (int)ch * 32,
                                    (int)((pair * a.gpi + j) * a.F + cb / bpr));
                tma_load_2d(&mapT, &raw_full[s], st + 2 * kYdATile, (int)ch * 32, 0);
            }
        }
    } else if (warp == 1) {
        // =========================================================================== MMA issuer
        // Descriptors as (constant high word, 32-bit low word) so that the per-MMA arithmetic is one uniform add; the waits
        // of the NEXT chunk are tested (non-blocking) before the last four MMAs of the current one, so their ~90 clk pass
        // while the tensor pipe still has queued work (same scheme as pw2_kernel's issuer).
        const uint32_t sbase = smem_u32(smem), lbase = smem_u32(lo_ring);
        const uint32_t idesc128 = make_idesc(128, 128) | (1u << 15);                // A MN-major, B K-major
        const uint32_t idesc64 = make_idesc(128, 64) | (1u << 15);
        constexpr uint32_t kAHi = (512u >> 4) | (1u << 14) | (1u << 29);            // SBO = 4-row k group, version 1, SWIZZLE_128B_BASE32B
        constexpr uint32_t kALoF = (4096u >> 4) << 16;                              // LBO = next 32-channel block
        constexpr uint32_t kBHi = (1024u >> 4) | (1u << 14) | (2u << 29);           // K-major SWIZZLE_128B
        constexpr uint32_t kBLoF = 1u << 16;
        auto mma = [](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accum) {
            asm volatile(
                "{\n\t.reg .pred p, q;\n\t.reg .b64 ad, bd;\n\tmov.b64 ad, {%1, %5};\n\tmov.b64 bd, {%2, %6};\n\t"
                "elect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accum),
                "r"(kAHi), "r"(kBHi) : "memory");
        };
        uint32_t q = 0, s = 0, sph = 0, l = 0, lph = 0;
        bool ready = false;
        for (uint32_t it = 0; it < (uint32_t)my_pairs; ++it) {
            mbar_wait_warp(d_empty, (it & 1) ^ 1);
            tc_fence_after();
            for (uint32_t ch = 0; ch < nch; ++ch, ++q) {
                if (!ready) {
                    mbar_wait_warp(&raw_full[s], sph);
                    if (!a.single_pass) mbar_wait_warp(&lo_full[l], lph);
                }
                tc_fence_after();
                const uint32_t tb = uniform_u32(tmem_base + (q & 0u));
                const uint32_t st16 = uniform_u32(((sbase + s * kYdStage) >> 4) & 0x3fffu), lt16 = uniform_u32(((lbase + l * kYdLoStage) >> 4) & 0x3fffu);
                const uint32_t t16 = st16 + ((2 * kYdATile) >> 4);
                const uint32_t first = ch < 2u ? 0u : 1u;
                uint32_t s1 = s + 1, sph1 = sph, l1 = l + 1, lph1 = lph;
                if (s1 == (uint32_t)kYdRaw) { s1 = 0; sph1 ^= 1u; }
                if (l1 == (uint32_t)kYdLo) { l1 = 0; lph1 ^= 1u; }
                const bool more = q + 1 < total;
                bool next_ready = false;
                auto test_next = [&]() {
                    next_ready = more && mbar_test_warp(&raw_full[s1], sph1) && (a.single_pass || mbar_test_warp(&lo_full[l1], lph1));
                };
                // x_raw * [tab_raw ; tab_lo] -> (main | corr) of this chunk parity, both rows, then x_lo * tab_raw -> corr
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint32_t d = tb + (uint32_t)j * 256u + (ch & 1u) * 128u;
                    if (j == 1 && a.single_pass) test_next();
                    if (j == 1 && a.gpi == 1) break;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        mma(d, (st16 + (uint32_t)(j * (kYdATile >> 4) + ks * 64)) | kALoF, (t16 + 2u * ks) | kBLoF,
                            a.single_pass ? idesc64 : idesc128, (ks == 0) ? first : 1u);
                }
                if (!a.single_pass) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t dc = tb + (uint32_t)j * 256u + (ch & 1u) * 128u + 64u;
                        if (j == 1) test_next();
                        if (j == 1 && a.gpi == 1) break;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            mma(dc, (lt16 + (uint32_t)(j * (kYdATile >> 4) + ks * 64)) | kALoF, (t16 + 2u * ks) | kBLoF, idesc64, 1u);
                    }
                    umma_commit(&lo_empty[l]);
                }
                umma_commit(&raw_empty[s]);
                s = s1; sph = sph1; l = l1; lph = lph1;
                ready = next_ready;
            }
            umma_commit(d_full);
        }
    } else if (warp < 6) {
        // =========================================================================== converters: lo = raw - trunc(raw)
        if (!a.single_pass) {
            const uint32_t ct = threadIdx.x - 64u;                   // 0..127
            for (uint32_t q = 0; q < total; ++q) {
                const uint32_t s = q % kYdRaw, l = q % kYdLo;
                mbar_wait(&raw_full[s], (q / kYdRaw) & 1);
                mbar_wait(&lo_empty[l], ((q / kYdLo) & 1) ^ 1);
                const uint32_t src = smem_u32(smem) + s * kYdStage + ct * 16u, dst = smem_u32(lo_ring) + l * kYdLoStage + ct * 16u;
                // layout-preserving copies, 2 KB per pass of the 128 threads: the two activation tiles into the lo ring,
                // the table chunk into the second half of its own stacked slab
#pragma unroll 4
                for (uint32_t off = 0; off < 2 * kYdATile + 8192u; off += 2048u) {
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src + off));
                    v.x -= __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); v.y -= __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                    v.z -= __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); v.w -= __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                    const uint32_t to = off < 2 * kYdATile ? dst + off : src + off + 8192u;
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(to), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&lo_full[l]); mbar_arrive(&raw_empty[s]); }
            }
        }
    } else {
        // =========================================================================== epilogue (8 warps)
        const int q4 = warp & 3, half = (warp - 6) >> 2;             // TMEM lane quarter, column half of the 64 table rows
        const int c = q4 * 32 + lane;
        const int fr = c / a.C, cc = c - fr * a.C;                      // lane = (row within the group, channel)
        const int l0 = half * 32;
        for (uint32_t it = 0; it < (uint32_t)my_pairs; ++it) {
            long long pair = (long long)blockIdx.x + (long long)it * gridDim.x;
            float* outp = a.out;
            if (a.nseg > 1) { const long long seg = pair / a.pairs0; pair -= seg * a.pairs0; outp += seg * a.seg_stride; }
            mbar_wait(d_full, it & 1);
            tc_fence_after();
            // accumulators -> registers for BOTH rows first, then the buffer is handed back: the issuer starts the next pair
            // while this warp's 64 global stores are still in flight (the accumulators are single buffered)
            float v[2][32];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j >= a.gpi) break;
                const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)j * 256u + (uint32_t)l0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {                       // 16 columns at a time: 64 results + 32 temporaries live
                    uint32_t m0[16], m1[16];
                    tmem_ld16(taddr + h * 16, m0);                  // main of the even chunks
                    if (nch > 1) tmem_ld16(taddr + 128u + h * 16, m1);      // main of the odd chunks
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[j][h * 16 + i] = __uint_as_float(m0[i]) + (nch > 1 ? __uint_as_float(m1[i]) : 0.f);
                    if (!a.single_pass) {
                        tmem_ld16(taddr + 64u + h * 16, m0);        // corrections, even / odd chunks
                        if (nch > 1) tmem_ld16(taddr + 192u + h * 16, m1);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[j][h * 16 + i] += __uint_as_float(m0[i]) + (nch > 1 ? __uint_as_float(m1[i]) : 0.f);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d_empty);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j >= a.gpi) break;
                const long long row = (pair * a.gpi + j) * a.F + fr;
                if (row < a.rows) {
                    float* op = outp + ((size_t)row * a.LY + l0) * a.C + cc;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (l0 + i < a.LY) op[(size_t)i * a.C] = v[j][i];
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc

using namespace tc;

// K segments for problems with few long rows (the 1-D models): as many as bring the item count up to the SM count, at most 8,
// at least two 32-point chunks each; 1 = no split.  Uses the device's SM count, not the caller's current budget, so that
// workspace sizes do not depend on fnm_set_sm_limit.
int tc_ydft_tma_segments(int64_t rows, int P2, int C) {
    if (C != 128 && C != 64 && C != 32) return 1;
    const long long groups = (rows + 128 / C - 1) / (128 / C);
    const int nch = (P2 + 31) / 32;
    if (groups < 1 || groups * 2 > sm_count_max()) return 1;
    long long n = sm_count_max() / groups;
    if (n > 8) n = 8;
    if (n > nch / 2) n = nch / 2;
    return n < 2 ? 1 : (int)n;
}

bool tc_ydft_tma_supported(const float* x, const float* tab, int ld, int64_t rows, int P2, int LY, int C, int act, int nseg) {
    // ld: row stride of the table in floats (TMA wants 16-byte strides: P2 itself, or the padded copy fnm_plan.ay4 / byT4)
    if (!tc_enabled() || act != 0 || (C != 128 && C != 64 && C != 32) || LY < 8 || LY > 64 || ld % 4 != 0 || ld < P2 || P2 < 4 ||
        rows < 1 || rows >= (1LL << 30))
        return false;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(tab)) & 15) return false;
    // an item is a whole (folded) grid row: problems with few long rows (1-D) stay on the segmenting register-path kernel
    if (C != 128 && rows / (128 / C) < sm_count() && nseg <= 1) return false;
    const char* e = getenv("FNM_YDFT_NO_TMA");
    return !(e && e[0] == '1');
}

int tc_ydft_tma(const float* x, const float* tab, int ld, float* T, int64_t rows, int P2, int LY, int C, int mode, cudaStream_t st,
                int nseg) {
    YdArgs a{};
    a.out = T; a.rows = rows; a.P2 = P2; a.LY = LY; a.NT = (LY + 15) / 16 * 16; a.nch = (P2 + 31) / 32;
    a.single_pass = mode == FNM_MODE_TF32;
    a.C = C; a.F = 128 / C;
    a.groups = (rows + a.F - 1) / a.F;
    a.gpi = a.groups >= 4LL * sm_count() ? 2 : 1;            // pair the groups when there are enough of them to keep every SM busy
    a.pairs = (a.groups + a.gpi - 1) / a.gpi;
    a.nseg = nseg > 1 ? nseg : 1; a.pairs0 = a.pairs; a.seg_stride = (long long)rows * LY * C;
    if (a.nseg > 1) {                                        // T holds nseg partial transforms [seg][row][l'][c]
        a.nch = (a.nch + a.nseg - 1) / a.nseg;               // chunks past the row's end read as zeros (TMA bounds)
        a.pairs = a.pairs0 * a.nseg;
    }
    CUtensorMap mx, mt;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)P2, (cuuint64_t)rows};
        const cuuint64_t str[2] = {(cuuint64_t)C * 4, (cuuint64_t)P2 * C * 4};
        const cuuint32_t box[3] = {32, 32, 1};
        if (!tc_make_map_swz(&mx, x, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return fail(FNM_ECUDA, "ydft: tensor map (x) failed");
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)P2, (cuuint64_t)LY};
        const cuuint64_t str[1] = {(cuuint64_t)ld * 4};
        const cuuint32_t box[2] = {32, 64};
        if (!tc_make_map(&mt, tab, 2, dims, str, box, true)) return fail(FNM_ECUDA, "ydft: tensor map (table) failed");
    }
    const size_t smem = (size_t)kYdRaw * kYdStage + (size_t)kYdLo * kYdLoStage + 1024 + 256;
    static std::atomic<uint64_t> attr_set{0};
    int attr_set_dev = 0;
    if (once_needed(attr_set, attr_set_dev)) {
        const cudaError_t e = cudaFuncSetAttribute(ydft_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(FNM_ECUDA, "ydft: smem attribute: %s", cudaGetErrorString(e));
        once_done(attr_set, attr_set_dev);
    }
    const int grid = a.pairs < sm_count() ? (int)a.pairs : sm_count();
    {
        LaunchScope ls("ydft", st);
        ydft_tma<<<grid, kYdThreads, smem, st>>>(mx, mt, a);
    }
    return check_launch("ydft");
}

}  // namespace fnm

// Measured and dropped: the same operand path for xdft (item = (sample, ky), K = 2 P1, table 2R x 2 P1 streamed per item).
// There the streamed table chunk is as large as the activation chunk, which doubles the L2 -> shared-memory traffic, and
// the 3xTF32 pass ran at 0.17 ms against 0.14 ms for the register-staged table GEMM with its TMA-streamed table ring
// (single-pass TF32: 0.12 vs 0.13 ms).  Sharing one table chunk between several items would need more than the 512 TMEM
// columns that two alternating (main | correction) accumulator pairs of 128 table rows already take.
