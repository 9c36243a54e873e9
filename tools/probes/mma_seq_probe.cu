// Probe: pw2_kernel's exact per-tile tcgen05.mma sequence (config 5: 4 channel slabs + 8 mode k-steps, 3xTF32 = 72 MMAs
// of M = 128, N = 64) issued back to back by one warp with every barrier satisfied, to find which feature of the sequence
// makes an MMA cost ~51 clk in the kernel when the same instruction costs 32 clk in mma_rate_probe.  Variants:
//   bit 0: A operands alternate hi / lo (columns c and c + 128) on the SAME B k-slice (the kernel's order); else hi only x2
//   bit 1: accumulator at column 384 / 448 (kernel) instead of 128 / 192
//   bit 2: 8 epilogue warps drain the other accumulator with tcgen05.ld.x8 + global stores meanwhile
//   bit 3: 4 converter warps copy a stage (ld.shared.v4 / st.shared.v4) meanwhile
#include "fnm_tc_ptx.cuh"
#include <cstdio>
#include <vector>
using namespace fnm::tc;

constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t dlo(uint32_t saddr) { return ((saddr >> 4) & 0x3fffu) | (1u << 16); }
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint32_t b_lo, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 bd;\n\tmov.b64 bd, {%2, %5};\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %3, p;\n\t}" ::"r"(d), "r"(a), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi) : "memory");
}

constexpr int TILES = 64;

__global__ void __launch_bounds__(448, 1) probe(long long* out, float* sink, int variant, uint32_t zero) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    // 2 raw stages + 2 lo buffers of 48 KB (64 positions x (128 channels + 64 modes) x 4 B)
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * 49152);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 6);
    volatile int* stop = reinterpret_cast<volatile int*>(slot + 1);
    const int tid = threadIdx.x, warp = (int)uniform_u32(threadIdx.x >> 5), lane = tid & 31;
    for (int i = tid; i < 4 * 49152 / 16; i += 448) {
        uint32_t h = (uint32_t)i * 2654435761u;
        float4 v;
        v.x = (float)(int)(h & 0xffff) / 32768.f - 1.f; h = h * 1664525u + 1013904223u;
        v.y = (float)(int)(h & 0xffff) / 32768.f - 1.f; h = h * 1664525u + 1013904223u;
        v.z = (float)(int)(h & 0xffff) / 32768.f - 1.f; h = h * 1664525u + 1013904223u;
        v.w = (float)(int)(h & 0xffff) / 32768.f - 1.f;
        reinterpret_cast<float4*>(smem)[i] = v;
    }
    if (tid == 0) { for (int b = 0; b < 6; ++b) mbar_init(&bar[b], 1); fence_barrier_init(); *stop = 0; }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp < 4) {
        uint32_t z[16];
        for (int c = 0; c < 384; c += 16) {
            for (int j = 0; j < 16; ++j) z[j] = __float_as_uint((float)((tid * 31 + j * 7 + c) % 64) / 64.f - 0.5f);
            tmem_st16(((uint32_t)(warp * 32) << 16) + c, z);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        const uint32_t idesc = make_idesc(128, 64);
        const uint32_t sbase = smem_u32(smem);
        const uint32_t slab16 = 64u * 8u, xb16 = (4u * 64u * 128u) >> 4;
        const bool alt = variant & 1;
        const uint32_t dcol = (variant & 2) ? 384u : 128u;
        const long long t0 = clock64();
#pragma unroll 1
        for (int t = 0; t < TILES; ++t) {
            const uint32_t tb = uniform_u32(zero + (t & 0u));
            const uint32_t d = tb + dcol + (t & 1) * 64u;
            const uint32_t x_lo = uniform_u32(dlo(sbase + (t & 1) * 49152u)), by_lo = x_lo + xb16;
            const uint32_t xl_lo = uniform_u32(dlo(sbase + (2 + (t & 1)) * 49152u)), byl_lo = xl_lo + xb16;
            uint32_t accum = 0;
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t b = x_lo + kb * slab16 + 2 * ks;
                    mma(d, tb + kb * 32 + ks * 8, b, idesc, accum);
                    accum = 1;
                    mma(d, tb + (alt ? 128u : 0u) + kb * 32 + ks * 8, b, idesc, 1);
                }
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8) {
                const uint32_t b = by_lo + (k8 >> 2) * slab16 + 2 * (k8 & 3);
                mma(d, tb + 256 + k8 * 8, b, idesc, 1);
                mma(d, tb + (alt ? 320u : 256u) + k8 * 8, b, idesc, 1);
            }
            if (variant & 16) umma_commit(&bar[1]);
            if (variant & 32) tc_fence_after();
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma(d, tb + kb * 32 + ks * 8, xl_lo + kb * slab16 + 2 * ks, idesc, 1);
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8) mma(d, tb + 256 + k8 * 8, byl_lo + (k8 >> 2) * slab16 + 2 * (k8 & 3), idesc, 1);
            if (variant & 16) { umma_commit(&bar[2]); umma_commit(&bar[3]); }
            if (variant & 32) tc_fence_after();
            if (variant & 64) __syncwarp();
        }
        const long long t1 = clock64();
        umma_commit(&bar[0]);
        mbar_wait_warp(&bar[0], 0);
        const long long t2 = clock64();
        if (lane == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = t2 - t0; }
        __syncwarp();
        if (lane == 0) { *stop = 1; mbar_arrive(&bar[4]); }
    } else if (variant & 128) {
        // every other warp polls an mbarrier that completes only at the end (what 17 warps of pw2_kernel do while they wait)
        mbar_wait(&bar[4], 0);
    } else if (warp >= 2 && warp < 6) {
        // converter stand-in
        const uint32_t src0 = smem_u32(smem) + (uint32_t)(tid - 64) * 16u, dst0 = src0 + 2 * 49152u;
        while (*stop == 0) {
            if (variant & 8) {
                for (uint32_t off = 0; off + 4096u <= 49152u; off += 4096u) {
                    float4 v0, v1;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w) : "r"(src0 + off));
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w) : "r"(src0 + off + 2048u));
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst0 + off), "f"(v0.x), "f"(v0.y), "f"(v0.z), "f"(v0.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst0 + off + 2048u), "f"(v1.x), "f"(v1.y), "f"(v1.z), "f"(v1.w) : "memory");
                }
            } else __nanosleep(100);
        }
    } else if (warp >= 6) {
        // epilogue stand-in: 8 warps, lane = channel, x8 loads + coalesced stores
        const int q4 = warp & 3;
        const uint32_t lane_base = ((uint32_t)(q4 * 32) << 16) + ((variant & 2) ? 448u : 192u) + ((warp - 6) >> 2) * 32u;
        float* op = sink + (size_t)blockIdx.x * 64 * 128 + q4 * 32 + lane;
        while (*stop == 0) {
            if (variant & 4) {
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[8];
                    tmem_ld8(lane_base + c * 8, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) op[(size_t)(c * 8 + j) * 128] = __uint_as_float(r[j]) + 1.f;
                }
            } else __nanosleep(100);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(0, 512);
}

int main() {
    long long* d; float* sink;
    cudaMalloc(&d, 148 * 2 * sizeof(long long));
    cudaMalloc(&sink, (size_t)148 * 64 * 128 * 4);
    const size_t smem = 4 * 49152 + 1024 + 128;
    if (cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { printf("attr failed\n"); return 1; }
    printf("variant bits: 1 = hi/lo A alternation, 2 = D at col 384, 4 = epilogue warps draining, 8 = converter warps copying, 16 = 3 commits per tile, 32 = 2 fence::after per tile, 64 = syncwarp per tile\n");
    for (int variant : {0, 128, 255}) {
        for (int it = 0; it < 2; ++it) {
            probe<<<148, 448, smem>>>(d, sink, variant, 0u);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error: %s (variant %d)\n", cudaGetErrorString(e), variant); return 1; }
        }
        std::vector<long long> h(2 * 148);
        cudaMemcpy(h.data(), d, sizeof(long long) * 2 * 148, cudaMemcpyDeviceToHost);
        double issue = 0, total = 0;
        for (int b = 0; b < 148; ++b) { issue += (double)h[2 * b]; total += (double)h[2 * b + 1]; }
        printf("variant %2d: issue %7.1f clk/tile (%.1f per MMA), issue+retire %7.1f clk/tile (%.1f per MMA)\n", variant, issue / 148 / TILES,
               issue / 148 / TILES / 72, total / 148 / TILES, total / 148 / TILES / 72);
    }
    return 0;
}
