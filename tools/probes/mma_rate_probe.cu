// Probe: cost in SM clocks of one tcgen05.mma (M = 128, cta_group::1) as a function of N, operand source
// (A in TMEM vs A in shared memory) and kind (tf32 K = 8, bf16 K = 16), measured as the issue-to-retire
// time of a chain of NMMA instructions on one accumulator / on two alternating accumulators.
// It answers what DESIGN.md section 4 left open: is the ~50 clk / MMA of pw2_kernel a per-instruction floor
// (then wider N is nearly free) or a throughput limit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../fourier-neural-mappings_b200/csrc -I../../include \
//        mma_rate_probe.cu -o mma_rate_probe
#include "fnm_tc_ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace fnm::tc;

constexpr int NMMA = 512;

__device__ __forceinline__ uint32_t idesc_kind(int M, int N, int bf16) {
    const uint32_t fmt = bf16 ? 1u : 2u;                    // a/b format: 1 = BF16 (kind::f16), 2 = TF32 (kind::tf32)
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t at, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(at), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t at, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(at), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}


constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo32(uint32_t saddr) { return ((saddr >> 4) & 0x3fffu) | (1u << 16); }
#define MMA32(NAME, KIND, AOP, ACONS)                                                                                   \
__device__ __forceinline__ void NAME(uint32_t d, uint32_t a, uint32_t b_lo, uint32_t idesc, uint32_t accum) {           \
    asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 bd, ad;\n\tmov.b64 bd, {%2, %5};\n\tmov.b64 ad, {%1, %5};\n\t"       \
                 "elect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                            \
                 "@q tcgen05.mma.cta_group::1.kind::" KIND " [%0], " AOP ", bd, %3, p;\n\t}" ::"r"(d), "r"(a), "r"(b_lo), \
                 "r"(idesc), "r"(accum), "r"(kDescHi) : "memory");                                                      \
}
MMA32(mma32_tf32_ts, "tf32", "[%1]", 0)
MMA32(mma32_tf32_ss, "tf32", "ad", 0)
MMA32(mma32_f16_ts, "f16", "[%1]", 0)
MMA32(mma32_f16_ss, "f16", "ad", 0)

// cfg: N, a_tmem (1 = TS), bf16, alt (number of accumulators the chain rotates over: 1 or 2)
template <int a_tmem, int bf16>
__global__ void __launch_bounds__(384, 1) probe(long long* out, int N, int alt, int nwarps, int stress) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                    // A: 128 rows x 128 B x 4 k-slabs = 64 KB
    uint8_t* sb = smem + 65536;            // B: 256 rows x 128 B x 4 k-slabs = 128 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536 + 131072);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    for (int i = tid; i < (65536 + 131072) / 16; i += 384) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 97u;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (stress & 16) {                                   // random operands in [-1, 1) (bf16 kind reads them as pairs of bf16: still finite)
            v.x = (float)(int)(h & 0xffff) / 32768.f - 1.f; h = h * 1664525u + 1013904223u;
            v.y = (float)(int)(h & 0xffff) / 32768.f - 1.f; h = h * 1664525u + 1013904223u;
            v.z = (float)(int)(h & 0xffff) / 32768.f - 1.f; h = h * 1664525u + 1013904223u;
            v.w = (float)(int)(h & 0xffff) / 32768.f - 1.f;
        }
        reinterpret_cast<float4*>(smem)[i] = v;
    }
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = __shfl_sync(0xffffffffu, *slot, 0);
    volatile int* stop = reinterpret_cast<volatile int*>(slot + 1);
    if (tid == 0) { *stop = 0; slot[2] = 0; slot[3] = 0; }
    if (warp < 4) {   // zero the A region of TMEM (columns 0..127)
        uint32_t z[16];
        for (int j = 0; j < 16; ++j) z[j] = (stress & 16) ? __float_as_uint((float)((tid * 31 + j * 7) % 64) / 64.f - 0.5f) : 0u;
        for (int c = 0; c < 128; c += 16) tmem_st16(tb + ((uint32_t)(warp * 32) << 16) + c, z);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // issuers: warp 0 (and warp 1 when nwarps == 2, on its own accumulator and barrier)
    if (warp < nwarps) {
        const uint32_t idesc = idesc_kind(128, N, bf16);
        // stress bit 5 (32): descriptor bases of unknown uniformity (a per-lane load) -> the compiler keeps them in vector
        // registers and moves every descriptor to the uniform file with R2UR, as in pw2_kernel's issue loop
        const uint32_t vz = (stress & 32) ? (uint32_t)reinterpret_cast<volatile int*>(slot)[2 + (tid & 1)] : 0u;
        const uint32_t a_lo = desc_lo32(smem_u32(sa)) + vz, b_lo = desc_lo32(smem_u32(sb)) + vz;
        const uint32_t slabA = (128 * 128) >> 4, slabB = (256 * 128) >> 4;
        const uint32_t dbase = tb + 128 + (uint32_t)warp * (nwarps == 2 ? 128u : 0u);
        const uint32_t d0 = dbase, d1 = dbase + (alt == 2 ? (nwarps == 1 && N <= 128 ? 128u : (N <= 64 ? 64u : 0u)) : 0u);
        const long long t0 = clock64();
#pragma unroll 1
        for (int rep = 0; rep < NMMA / 16; ++rep) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t d = (k & 1) ? d1 : d0;
                const uint32_t bd = b_lo + (uint32_t)((k >> 2) * slabB + 2 * (k & 3));
                const uint32_t acc = (rep | (k >> 1)) ? 1u : 0u;
                if (a_tmem) {
                    const uint32_t at = tb + (uint32_t)(k * 8);
                    if (bf16) mma32_f16_ts(d, at, bd, idesc, acc); else mma32_tf32_ts(d, at, bd, idesc, acc);
                } else {
                    const uint32_t ad = a_lo + (uint32_t)((k >> 2) * slabA + 2 * (k & 3));
                    if (bf16) mma32_f16_ss(d, ad, bd, idesc, acc); else mma32_tf32_ss(d, ad, bd, idesc, acc);
                }
            }
        }
        const long long t1 = clock64();
        umma_commit(&bar[warp]);
        mbar_wait(&bar[warp], 0);
        const long long t2 = clock64();
        if ((tid & 31) == 0) { out[4 * blockIdx.x + 2 * warp] = t1 - t0; out[4 * blockIdx.x + 2 * warp + 1] = t2 - t0; }
        __syncwarp();
        if ((tid & 31) == 0) *stop = 1;
    } else if (warp >= 4 && warp < 8) {
        // stress warps 4..7 (TMEM lanes 32 * (warp & 3)); bit 0: tcgen05.ld of a 64-column accumulator region,
        // bit 1: shared-memory copy (ld.shared.v4 + st.shared.v4) inside the operand region, bit 2: tcgen05.st
        const uint32_t lane_base = tb + ((uint32_t)((warp & 3) * 32) << 16);
        float sink = 0.f;
        uint32_t off = (uint32_t)(tid - 128) * 16u;
        while (*stop == 0) {
            if (stress & 1) {
                uint32_t r[32];
                tmem_ld32(lane_base + 384, r);
                tmem_ld_wait();
                sink += __uint_as_float(r[0] ^ r[31]);
            }
            if (stress & 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 v;
                    const uint32_t src = smem_u32(sb) + ((off + j * 2048u) & 0xffffu), dst = smem_u32(sa) + ((off + j * 2048u) & 0xffffu);
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src));
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
                off += 16384u;
            }
            if (stress & 4) {
                uint32_t z[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) z[j] = 0;
                tmem_st16(lane_base + 448, z);
                tmem_st_wait();
            }
        }
        if (sink == 123.f) out[0] = 0;
    } else if (warp >= 8) {
        // bit 3: fp32 math spin on every sub-partition (issue-slot contention with the MMA warp)
        float a = (float)tid, b = 1.0001f;
        while (*stop == 0) {
            if (stress & 8) {
#pragma unroll
                for (int j = 0; j < 64; ++j) a = fmaf(a, b, 0.5f);
            } else {
                __nanosleep(200);
            }
        }
        if (a == 123.f) out[0] = 0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * 4 * sizeof(long long));
    const size_t smem = 65536 + 131072 + 1024 + 64;
    cudaFuncSetAttribute(probe<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(probe<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(probe<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaFuncSetAttribute(probe<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { printf("attr failed\n"); return 1; }
    const int Ns[] = {32, 64, 96, 128, 192, 256};
    printf("%-6s %-4s %-5s %-4s %-5s %-5s %14s %14s %8s\n", "kind", "A", "N", "acc", "warps", "grid", "issue clk/MMA", "total clk/MMA", "floor");
    for (int grid : {148})
        for (int bf16 = 0; bf16 < 2; ++bf16)
            for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
                for (int nwarps = 1; nwarps <= 2; ++nwarps)
                    for (int alt = 1; alt <= 2; ++alt)
                        for (int N : Ns) for (int stress : {0, 32}) {
                            if (alt == 2 && (N > 128 || (nwarps == 2 && N > 64))) continue;
                            if (nwarps == 2 && N > 128) continue;
                            if (N > 128 || alt == 2 || nwarps == 2 || !a_tmem || bf16) { if (stress) continue; }
                            for (int it = 0; it < 2; ++it) {                     // second run is the warm one
                                if (a_tmem && bf16) probe<1, 1><<<grid, 384, smem>>>(d, N, alt, nwarps, stress);
                                else if (a_tmem) probe<1, 0><<<grid, 384, smem>>>(d, N, alt, nwarps, stress);
                                else if (bf16) probe<0, 1><<<grid, 384, smem>>>(d, N, alt, nwarps, stress);
                                else probe<0, 0><<<grid, 384, smem>>>(d, N, alt, nwarps, stress);
                                cudaError_t e = cudaDeviceSynchronize();
                                if (e != cudaSuccess) { printf("CUDA error: %s (N=%d a_tmem=%d bf16=%d)\n", cudaGetErrorString(e), N, a_tmem, bf16); return 1; }
                            }
                            std::vector<long long> h(4 * grid);
                            cudaMemcpy(h.data(), d, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost);
                            double issue = 0, total = 0;
                            for (int b = 0; b < grid; ++b) { issue += (double)h[4 * b]; total += (double)(nwarps == 2 ? (h[4 * b + 1] > h[4 * b + 3] ? h[4 * b + 1] : h[4 * b + 3]) : h[4 * b + 1]); }
                            // per-MMA cost over ALL MMAs issued by the CTA (nwarps * NMMA)
                            printf("%-6s %-4s %-5d %-4d %-5d %-5d s%-3d %14.1f %14.1f %8.1f\n", bf16 ? "bf16" : "tf32", a_tmem ? "TMEM" : "SMEM", N, alt, nwarps, grid, stress,
                                   issue / grid / NMMA, total / grid / (NMMA * nwarps), 128.0 * N / 256.0);
                        }
    return 0;
}
