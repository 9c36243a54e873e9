// Probe: which shared-memory layouts / descriptor fields does tcgen05.mma kind::tf32 accept for an MN-major operand?
// One CTA computes D[128][64] = A[128][K] * B[64][K]^T (K = 32) with A K-major (known good) and B stored MN-major
// (n contiguous: rows = k, 128-byte rows of 32 n values, SWIZZLE_128B), under several (LBO, SBO) choices, and compares
// with the host.  Also A MN-major (SS form).  Build: nvcc -arch=sm_100a -I../../fourier-neural-mappings_b200/csrc ...
#include "fnm_tc_ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

using namespace fnm::tc;

__device__ __forceinline__ uint64_t desc_generic(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}

// variant: 0 = B K-major reference; 1..: B MN-major with different descriptor choices
__global__ void __launch_bounds__(128, 1) probe(const float* A, const float* B, float* D, int variant) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                 // A K-major: 4 chunks (32 k) -> slab [128 rows][128 B] per 32 k: K = 32 -> one slab, 16 KB
    uint8_t* sb = smem + 16384;         // B region: 16 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    // A[m][k], K = 32: row m at m*128 B, 16-byte chunk c swizzled
    for (int i = tid; i < 128 * 8; i += 128) {
        const int m = i >> 3, c = i & 7;
        float4 v = *reinterpret_cast<const float4*>(A + (size_t)m * 32 + c * 4);
        *reinterpret_cast<float4*>(sa + swz(m, c)) = v;
    }
    if (variant == 0) {
        for (int i = tid; i < 64 * 8; i += 128) {      // B K-major: row n, k contiguous
            const int n = i >> 3, c = i & 7;
            float4 v = *reinterpret_cast<const float4*>(B + (size_t)n * 32 + c * 4);
            *reinterpret_cast<float4*>(sb + swz(n, c)) = v;
        }
    } else {
        // B MN-major: n-block j (32 n) is a slab of 32 k rows x 128 B at sb + j*4096; row k holds n = 32j .. 32j+31
        for (int i = tid; i < 2 * 32 * 8; i += 128) {
            const int j = i / 256, r = (i % 256) >> 3, c = i & 7;     // r = k
            float4 v;
            v.x = B[(size_t)(32 * j + c * 4 + 0) * 32 + r]; v.y = B[(size_t)(32 * j + c * 4 + 1) * 32 + r];
            v.z = B[(size_t)(32 * j + c * 4 + 2) * 32 + r]; v.w = B[(size_t)(32 * j + c * 4 + 3) * 32 + r];
            // variants >= 6: SWIZZLE_128B_BASE32B (cute Swizzle<2,5,2>): 32-byte chunks of a 128-byte row XORed with (row & 3)
            const uint32_t off = variant >= 6 ? (uint32_t)r * 128u + ((((uint32_t)c >> 1) ^ ((uint32_t)r & 3u)) << 5) + ((uint32_t)c & 1u) * 16u
                                              : swz(r, c);
            *reinterpret_cast<float4*>(sb + j * 4096 + off) = v;
        }
    }
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *slot;
    if (warp == 0) {
        uint32_t idesc = make_idesc(128, 64);
        if (variant != 0) idesc |= (1u << 16);
        for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = make_desc(smem_u32(sa)) + 2 * ks;
            uint64_t bd;
            switch (variant) {
                case 0: bd = make_desc(smem_u32(sb)) + 2 * ks; break;
                case 1: bd = desc_generic(smem_u32(sb) + ks * 1024, 4096, 1024); break;     // LBO = n-block stride, SBO = k-group stride
                case 2: bd = desc_generic(smem_u32(sb) + ks * 1024, 1024, 4096); break;     // swapped
                case 3: bd = desc_generic(smem_u32(sb) + ks * 1024, 4096, 4096); break;
                case 4: bd = desc_generic(smem_u32(sb) + ks * 1024, 1024, 1024); break;
                case 5: bd = desc_generic(smem_u32(sb) + ks * 1024, 16, 4096); break;
                case 6: bd = desc_generic(smem_u32(sb) + ks * 1024, 4096, 512, 1); break;    // BASE32B: LBO = n-block stride, SBO = 4-row k atom stride
                case 7: bd = desc_generic(smem_u32(sb) + ks * 1024, 512, 4096, 1); break;    // swapped
                case 8: bd = desc_generic(smem_u32(sb) + ks * 1024, 4096, 1024, 1); break;
                default: bd = desc_generic(smem_u32(sb) + ks * 1024, 4096, 512, 2); break;   // BASE32B data, plain 128B descriptor
            }
            umma_ss(tb, ad, bd, idesc, ks ? 1u : 0u);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    uint32_t r[32];
    const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16);
    for (int g = 0; g < 2; ++g) {
        tmem_ld32(taddr + g * 32, r);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[(size_t)tid * 64 + g * 32 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 64);
}

int main() {
    std::vector<float> A(128 * 32), B(64 * 32), D(128 * 64), R(128 * 64);
    srand(1);
    for (auto& v : A) v = (rand() % 17 - 8) / 8.0f;       // exactly representable in TF32
    for (auto& v : B) v = (rand() % 13 - 6) / 4.0f;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            float s = 0;
            for (int k = 0; k < 32; ++k) s += A[m * 32 + k] * B[n * 32 + k];
            R[m * 64 + n] = s;
        }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
    for (int variant = 0; variant <= 9; ++variant) {
        cudaMemset(dD, 0, D.size() * 4);
        probe<<<1, 128, 36 * 1024>>>(dA, dB, dD, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0, ref = 0;
        int bad_cols = 0;
        for (int n = 0; n < 64; ++n) {
            double ce = 0;
            for (int m = 0; m < 128; ++m) { const double d = D[m * 64 + n] - R[m * 64 + n]; ce += d * d; ref += (double)R[m * 64 + n] * R[m * 64 + n]; }
            err += ce;
            if (ce > 1e-6) ++bad_cols;
        }
        printf("variant %d: rel err %.3e, wrong columns %d / 64   D[0][0..3] = %g %g %g %g (ref %g %g %g %g)\n", variant, sqrt(err / ref), bad_cols,
               D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
    }
    return 0;
}
