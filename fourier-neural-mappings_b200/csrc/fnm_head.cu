// The two ends of the trunk, fused so that no activation-sized temporary is made:
//
//   lift      fc0 on [data channels, grid coordinates] + zero padding, written channels-last padded
//             (models/func_to_func2d.py:79-86, shared.py:111-118,150-159) and its weight gradient
//   mlp head  fc2(GELU(h1)) of the projection MLP (shared.py:26-45; func_to_func2d.py:104-107) and its
//             backward "gate"  g_h1 = (g_out w2) * GELU'(h1)  together with dW2 / db2
//
// fc1 itself (a 1x1 convolution C -> width_final over every position) runs on the TMA/tcgen05 fused
// pointwise kernel with no spectral term (fnm_pointwise_ysyn, LY = 0), its weight gradient on
// conv_wgrad.  These kernels are plain bandwidth-bound SIMT passes.
#include "fnm_tc_ptx.cuh"

namespace fnm {

constexpr int kLiftMaxJ = 8;          // data channels + grid coordinates

struct LiftArgs {
    const float* x; const float* w; const float* b; float* out;
    const float* g; float* partial;
    int B, Din, N1, N2, P1, P2, C, ngrid, J;
    float step1, step2;                   // torch.linspace(0, 1, n) step 1 / (n - 1) in fp32
};

// torch.linspace(0, 1, n) as ATen evaluates it in fp32: start + i * step in the first half, end - (n - 1 - i) * step in
// the second (no division per element)
__device__ __forceinline__ float linspace01(int i, int n, float step) {
    return i < n / 2 ? (float)i * step : 1.0f - (float)(n - 1 - i) * step;
}

__device__ __forceinline__ float lift_feat(const LiftArgs& a, int j, int b, int xx, int yy) {
    if (j < a.Din) return __ldg(a.x + (((size_t)b * a.Din + j) * a.N1 + xx) * a.N2 + yy);
    const int gi = j - a.Din;
    // 2-D: first coordinate along x, second along y; 1-D: along y
    if (a.ngrid == 2 && gi == 0) return a.N1 > 1 ? linspace01(xx, a.N1, a.step1) : 0.f;
    return a.N2 > 1 ? linspace01(yy, a.N2, a.step2) : 0.f;
}

// thread = (position slot, 4-channel group): its weight columns live in registers, the features of a position are
// evaluated once per thread (no per-element index division), one 16-byte store per position
template <int MAXJ>
__global__ void __launch_bounds__(256) lift_fwd_kernel(const LiftArgs a) {
    const int C4 = a.C / 4;
    const int slots = blockDim.x / C4;
    const int c4 = threadIdx.x % C4, slot = threadIdx.x / C4;
    if (slot >= slots) return;
    float4 w[MAXJ], bias = __ldg(reinterpret_cast<const float4*>(a.b) + c4);
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
        w[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < a.J) {
            w[j].x = __ldg(a.w + (size_t)(c4 * 4 + 0) * a.J + j); w[j].y = __ldg(a.w + (size_t)(c4 * 4 + 1) * a.J + j);
            w[j].z = __ldg(a.w + (size_t)(c4 * 4 + 2) * a.J + j); w[j].w = __ldg(a.w + (size_t)(c4 * 4 + 3) * a.J + j);
        }
    }
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const int nrows = a.B * a.P1;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {          // row = (b, x) of the padded grid
        const int b = row / a.P1, xx = row - b * a.P1;
        float4* orow = reinterpret_cast<float4*>(a.out + (size_t)row * a.P2 * a.C) + c4;
        const bool x_in = xx < a.N1;
        for (int yy = slot; yy < a.P2; yy += slots) {
            float4 v = zero;
            if (x_in && yy < a.N2) {
                v = bias;
#pragma unroll
                for (int j = 0; j < MAXJ; ++j) {
                    if (j < a.J) {
                        const float f = lift_feat(a, j, b, xx, yy);
                        v.x = fmaf(f, w[j].x, v.x); v.y = fmaf(f, w[j].y, v.y); v.z = fmaf(f, w[j].z, v.z); v.w = fmaf(f, w[j].w, v.w);
                    }
                }
            }
            orow[(size_t)yy * C4] = v;
        }
    }
}

// partial[block][j][c] = sum over the block's positions of g[pos][c] * feat_j(pos)   (j == J: bias)
template <int MAXJ>
__global__ void __launch_bounds__(256, 3) lift_bwd_kernel(const LiftArgs a) {
    extern __shared__ float red[];                        // [slots][(J+1)][C]
    const int C4 = a.C / 4;
    const int slots = blockDim.x / C4;
    const int c4 = threadIdx.x % C4, slot = threadIdx.x / C4;
    float4 acc[MAXJ + 1];
#pragma unroll
    for (int j = 0; j <= MAXJ; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int nrows = a.B * a.N1;                         // un-padded rows (b, x < N1)
    if (slot < slots) {
        for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
            const int b = row / a.N1, xx = row - b * a.N1;
            const float4* grow = reinterpret_cast<const float4*>(a.g + ((size_t)b * a.P1 + xx) * a.P2 * a.C) + c4;
            for (int y0 = slot; y0 < a.N2; y0 += 4 * slots) {           // four independent loads in flight (eight spill at 128 registers)
                float4 g[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int yy = y0 + u * slots;
                    g[u] = yy < a.N2 ? __ldg(grow + (size_t)yy * C4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int yy = y0 + u * slots;
                    if (yy < a.N2) {
#pragma unroll
                        for (int j = 0; j < MAXJ; ++j) {
                            if (j < a.J) {
                                const float f = lift_feat(a, j, b, xx, yy);
                                acc[j].x = fmaf(f, g[u].x, acc[j].x); acc[j].y = fmaf(f, g[u].y, acc[j].y);
                                acc[j].z = fmaf(f, g[u].z, acc[j].z); acc[j].w = fmaf(f, g[u].w, acc[j].w);
                            }
                        }
                        acc[MAXJ].x += g[u].x; acc[MAXJ].y += g[u].y; acc[MAXJ].z += g[u].z; acc[MAXJ].w += g[u].w;
                    }
                }
            }
        }
        for (int j = 0; j <= a.J; ++j) {
            const float4 v = j < a.J ? acc[j] : acc[MAXJ];
            *reinterpret_cast<float4*>(red + ((size_t)slot * (a.J + 1) + j) * a.C + c4 * 4) = v;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (a.J + 1) * a.C; i += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < slots; ++k) s += red[(size_t)k * (a.J + 1) * a.C + i];
        a.partial[(size_t)blockIdx.x * (a.J + 1) * a.C + i] = s;
    }
}

// sum of the per-block partials: block = 32 consecutive outputs (threadIdx.x) x 32 groups of partials (threadIdx.y) that
// meet in shared memory in a fixed order -- a single thread per output walking all 592 partials took 40-55 us
__device__ __forceinline__ float partial_sum_32x32(const float* __restrict__ partial, int nblocks, int per, int i) {
    __shared__ float sm[32][33];
    float s = 0.f;
    if (i < per)
        for (int k = threadIdx.y; k < nblocks; k += 32) s += partial[(size_t)k * per + i];
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.y == 0) {
#pragma unroll
        for (int g = 0; g < 32; ++g) t += sm[g][threadIdx.x];
    }
    return t;
}

__global__ void __launch_bounds__(1024) lift_bwd_reduce(const float* __restrict__ partial, int nblocks, int J, int C,
                                                       float* __restrict__ dw, float* __restrict__ db) {
    const int i = blockIdx.x * 32 + threadIdx.x;
    const float s = partial_sum_32x32(partial, nblocks, (J + 1) * C, i);
    if (threadIdx.y != 0 || i >= (J + 1) * C) return;
    const int j = i / C, c = i - j * C;
    if (j < J) { if (dw) dw[(size_t)c * J + j] = s; }
    else if (db) db[c] = s;
}

// ---- projection head: out[pos][d] = sum_h w2[d][h] GELU(h1[pos][h]) + b2[d]
constexpr int kHeadMaxD = 4;
constexpr int kHeadMaxH4 = 1;         // H <= 128: one float4 group per lane (wider heads fall back to torch)

struct HeadArgs {
    const float* h1; const float* g_out; const float* w2; const float* b2;
    float* out; float* g_h1; float* partial;
    long long npos;
    int H, D;
};

// exact-erf GELU to ~1e-7 (Abramowitz-Stegun 7.1.26 on the MUFU pipe), as in the fused tcgen05 epilogues
__device__ __forceinline__ float gelu_erf(float x) { return tc::gelu_fast(x); }
__device__ __forceinline__ float gelu_erf_grad(float x) { return tc::gelu_grad_fast(x); }

// One warp per 8 positions: lane = 4 hidden channels, eight 16-byte loads in flight per lane.  The 8 (x D) per-lane
// partial dot products are summed over the warp with a halving butterfly -- lanes trade half of their values at each of
// the first three steps -- 9 shuffles instead of 40; lane 4*v then holds the total of position v.
template <int D>
__global__ void __launch_bounds__(256) head_fwd_kernel(const HeadArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int H4 = a.H / 4;
    const bool lane_ok = lane < H4;
    float4 w[D];
    float bias[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        w[d] = lane_ok ? __ldg(reinterpret_cast<const float4*>(a.w2 + (size_t)d * a.H) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        bias[d] = a.b2 ? __ldg(a.b2 + d) : 0.f;
    }
    const int vsel = ((lane >> 2) & 1) + 2 * ((lane >> 3) & 1) + 4 * ((lane >> 4) & 1);   // position this lane ends up owning
    for (long long p0 = ((long long)blockIdx.x * nw + warp) * 8; p0 < a.npos; p0 += (long long)gridDim.x * nw * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            v[u] = (lane_ok && p0 + u < a.npos) ? tc::ld_stream4(a.h1 + (p0 + u) * a.H + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        float s[D][8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 t = make_float4(gelu_erf(v[u].x), gelu_erf(v[u].y), gelu_erf(v[u].z), gelu_erf(v[u].w));
#pragma unroll
            for (int d = 0; d < D; ++d) s[d][u] = t.x * w[d].x + t.y * w[d].y + t.z * w[d].z + t.w * w[d].w;
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {                       // lanes with bit 4 set keep positions 4..7
                const bool hi = lane & 16;
                const float recv = __shfl_xor_sync(0xffffffffu, hi ? s[d][i] : s[d][i + 4], 16);
                s[d][i] = (hi ? s[d][i + 4] : s[d][i]) + recv;
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const bool hi = lane & 8;
                const float recv = __shfl_xor_sync(0xffffffffu, hi ? s[d][i] : s[d][i + 2], 8);
                s[d][i] = (hi ? s[d][i + 2] : s[d][i]) + recv;
            }
            {
                const bool hi = lane & 4;
                const float recv = __shfl_xor_sync(0xffffffffu, hi ? s[d][0] : s[d][1], 4);
                s[d][0] = (hi ? s[d][1] : s[d][0]) + recv;
            }
            s[d][0] += __shfl_xor_sync(0xffffffffu, s[d][0], 2);
            s[d][0] += __shfl_xor_sync(0xffffffffu, s[d][0], 1);
        }
        if ((lane & 3) == 0 && p0 + vsel < a.npos) {
#pragma unroll
            for (int d = 0; d < D; ++d) a.out[(p0 + vsel) * a.D + d] = s[d][0] + bias[d];
        }
    }
}

// g_h1[pos][h] = (sum_d g_out[pos][d] w2[d][h]) * GELU'(h1[pos][h]);  partial sums of
// dW2[d][h] = sum_pos g_out[pos][d] GELU(h1[pos][h]) and db2[d] = sum_pos g_out[pos][d]
template <int D>
__global__ void __launch_bounds__(256) head_bwd_kernel(const HeadArgs a) {
    extern __shared__ float red[];                        // [warps][D][H] + [warps][D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int H4 = a.H / 4;
    float4 w[D][kHeadMaxH4], dw[D][kHeadMaxH4];
    float dbs[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        dbs[d] = 0.f;
#pragma unroll
        for (int k = 0; k < kHeadMaxH4; ++k) {
            const int h4 = lane + 32 * k;
            w[d][k] = h4 < H4 ? __ldg(reinterpret_cast<const float4*>(a.w2 + (size_t)d * a.H) + h4) : make_float4(0.f, 0.f, 0.f, 0.f);
            dw[d][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const bool lane_ok = lane < H4;
    for (long long p0 = ((long long)blockIdx.x * nw + warp) * 8; p0 < a.npos; p0 += (long long)gridDim.x * nw * 8) {
        float4 v[8];
        float g[8][D];
#pragma unroll
        for (int u = 0; u < 8; ++u) {                     // eight 16-byte loads in flight per lane
            const bool ok = p0 + u < a.npos;
            v[u] = (lane_ok && ok) ? tc::ld_stream4(a.h1 + (p0 + u) * a.H + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int d = 0; d < D; ++d) g[u][d] = ok ? __ldg(a.g_out + (p0 + u) * D + d) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (p0 + u < a.npos) {
                float4 act, dact, gg = make_float4(0.f, 0.f, 0.f, 0.f);
                tc::gelu_both_fast(v[u].x, act.x, dact.x); tc::gelu_both_fast(v[u].y, act.y, dact.y);
                tc::gelu_both_fast(v[u].z, act.z, dact.z); tc::gelu_both_fast(v[u].w, act.w, dact.w);
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const float gd = g[u][d];
                    dbs[d] += gd;
                    gg.x = fmaf(gd, w[d][0].x, gg.x); gg.y = fmaf(gd, w[d][0].y, gg.y);
                    gg.z = fmaf(gd, w[d][0].z, gg.z); gg.w = fmaf(gd, w[d][0].w, gg.w);
                    dw[d][0].x = fmaf(gd, act.x, dw[d][0].x); dw[d][0].y = fmaf(gd, act.y, dw[d][0].y);
                    dw[d][0].z = fmaf(gd, act.z, dw[d][0].z); dw[d][0].w = fmaf(gd, act.w, dw[d][0].w);
                }
                gg.x *= dact.x; gg.y *= dact.y; gg.z *= dact.z; gg.w *= dact.w;
                if (lane_ok) *(reinterpret_cast<float4*>(a.g_h1 + (p0 + u) * a.H) + lane) = gg;
            }
        }
    }
    // block reduction of dW2 / db2 partials
    float* rw = red;                                       // [nw][D][H]
    float* rb = red + (size_t)nw * a.D * a.H;              // [nw][D]
#pragma unroll
    for (int d = 0; d < D; ++d) {
#pragma unroll
        for (int k = 0; k < kHeadMaxH4; ++k) {
            const int h4 = lane + 32 * k;
            if (h4 < H4) *reinterpret_cast<float4*>(rw + ((size_t)warp * a.D + d) * a.H + h4 * 4) = dw[d][k];
        }
        if (lane == 0) rb[warp * a.D + d] = dbs[d];
    }
    __syncthreads();
    const int per = a.D * a.H + a.D;
    for (int i = threadIdx.x; i < per; i += blockDim.x) {
        float s = 0.f;
        if (i < a.D * a.H) { for (int k = 0; k < nw; ++k) s += rw[(size_t)k * a.D * a.H + i]; }
        else { for (int k = 0; k < nw; ++k) s += rb[k * a.D + (i - a.D * a.H)]; }
        a.partial[(size_t)blockIdx.x * per + i] = s;
    }
}

__global__ void __launch_bounds__(1024) head_bwd_reduce(const float* __restrict__ partial, int nblocks, int H, int D,
                                                       float* __restrict__ dw2, float* __restrict__ db2) {
    const int per = D * H + D;
    const int i = blockIdx.x * 32 + threadIdx.x;
    const float s = partial_sum_32x32(partial, nblocks, per, i);
    if (threadIdx.y != 0 || i >= per) return;
    if (i < D * H) { if (dw2) dw2[i] = s; }
    else if (db2) db2[i - D * H] = s;
}


// ---- F2V local branch: the reference applies mlpfunc0 pointwise and integrates with
// two torch.trapz calls (func_to_vec2d.py:100-104).  fc2 is linear, so trapz(fc2(GELU(h1))) =
// fc2(sum_pos wt[pos] GELU(h1[pos])): only the weighted GELU sum S[b][h] touches the field.
//   S[b][h] = sum_{x,y} wx[x] wy[y] GELU(h1[b][x][y][h])          g_h1 = gS[b][h] wx[x] wy[y] GELU'(h1)
struct WsumArgs {
    const float* h1; const float* wx; const float* wy; const float* gS;
    float* partial; float* g_h1;
    int B, P1, P2, H, nchunk;
};

__global__ void wsum_fwd_kernel(const WsumArgs a) {
    extern __shared__ float red[];                        // [warps][H]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int H4 = a.H / 4, b = blockIdx.y, chunk = blockIdx.x;
    const int npos = a.P1 * a.P2;
    const int per = (npos + a.nchunk - 1) / a.nchunk;
    const int p_end = min(npos, (chunk + 1) * per);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* base = a.h1 + (size_t)b * npos * a.H;
    if (lane < H4) {
        for (int p0 = chunk * per + warp * 4; p0 < p_end; p0 += nw * 4) {
            float4 v[4];
            float w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int p = p0 + u;
                const bool ok = p < p_end;
                v[u] = ok ? __ldg(reinterpret_cast<const float4*>(base + (size_t)p * a.H) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                w[u] = ok ? __ldg(a.wx + p / a.P2) * __ldg(a.wy + p % a.P2) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc.x = fmaf(w[u], gelu_erf(v[u].x), acc.x); acc.y = fmaf(w[u], gelu_erf(v[u].y), acc.y);
                acc.z = fmaf(w[u], gelu_erf(v[u].z), acc.z); acc.w = fmaf(w[u], gelu_erf(v[u].w), acc.w);
            }
        }
        *reinterpret_cast<float4*>(red + (size_t)warp * a.H + lane * 4) = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.H; i += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < nw; ++k) s += red[(size_t)k * a.H + i];
        a.partial[((size_t)b * a.nchunk + chunk) * a.H + i] = s;
    }
}

__global__ void wsum_reduce_kernel(const float* __restrict__ partial, int nchunk, int BH, int H, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BH) return;
    const int b = i / H, h = i - b * H;
    float s = 0.f;
    for (int k = 0; k < nchunk; ++k) s += partial[((size_t)b * nchunk + k) * H + h];
    out[i] = s;
}

__global__ void wsum_bwd_kernel(const WsumArgs a) {
    const int H4 = a.H / 4;
    const int npos = a.P1 * a.P2;
    const int b = blockIdx.y;
    const float* base = a.h1 + (size_t)b * npos * a.H;
    float* gbase = a.g_h1 + (size_t)b * npos * a.H;
    const int total = npos * H4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int p = i / H4, h4 = i - p * H4;
        const float w = __ldg(a.wx + p / a.P2) * __ldg(a.wy + p % a.P2);
        const float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)p * a.H) + h4);
        const float4 g = __ldg(reinterpret_cast<const float4*>(a.gS + (size_t)b * a.H) + h4);
        float4 o;
        o.x = w * g.x * gelu_erf_grad(v.x); o.y = w * g.y * gelu_erf_grad(v.y);
        o.z = w * g.z * gelu_erf_grad(v.z); o.w = w * g.w * gelu_erf_grad(v.w);
        *(reinterpret_cast<float4*>(gbase + (size_t)p * a.H) + h4) = o;
    }
}

constexpr int kWsumChunks = 64;

constexpr int kEndBlocks = 148 * 4;

static bool lift_ok(int Din, int C, int ngrid) { return Din >= 0 && Din + ngrid >= 1 && Din + ngrid <= kLiftMaxJ && C % 4 == 0 && C >= 4 && C <= 1024; }

// ---- training loss of the reference harness: LpLoss(size_average=False) = sum_b ||out_b - y_b||_2 / (||y_b||_2 + eps)
// (util/utilities_module.py:188-204; train_fno.py:98).  Three launches instead of the ~12 elementwise / reduction kernels the
// torch statement and its autograd graph take -- what the launch-bound small configurations notice.
constexpr int kLossChunks = 64;                     // partial sums per sample

__global__ void __launch_bounds__(256) rel_l2_partial_kernel(const float* __restrict__ out, const float* __restrict__ y,
                                                            float* __restrict__ part, long long n, int vec) {
    // grid (chunk, sample): part[b][chunk] = (sum (out - y)^2, sum y^2) over the chunk's strided float4 groups
    // (vec == 0: sample rows are not 16-byte aligned, everything goes through the scalar tail on chunk 0)
    const long long base = (long long)blockIdx.y * n;
    float sn = 0.f, sd = 0.f;
    const long long n4 = vec ? n / 4 : 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float4 a = *reinterpret_cast<const float4*>(out + base + 4 * i), b = *reinterpret_cast<const float4*>(y + base + 4 * i);
        const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
        sn += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        sd += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    }
    if (blockIdx.x == 0 || !vec)
        for (long long i = 4 * n4 + (vec ? 0 : (long long)blockIdx.x * 256) + threadIdx.x; i < n; i += (vec ? 256 : (long long)gridDim.x * 256)) {
            const float a = out[base + i], b = y[base + i];
            sn += (a - b) * (a - b);
            sd += b * b;
        }
    __shared__ float s0[8], s1[8];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { sn += __shfl_xor_sync(0xffffffffu, sn, d); sd += __shfl_xor_sync(0xffffffffu, sd, d); }
    if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = sn; s1[threadIdx.x >> 5] = sd; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tn = 0.f, td = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { tn += s0[w]; td += s1[w]; }
        part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 + 0] = tn;
        part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 + 1] = td;
    }
}

// one block: per-sample norms, coef[b] = 1 / (||out_b - y_b|| (||y_b|| + eps)) (0 where the numerator vanishes: the subgradient
// torch.norm's backward uses), loss = sum_b ||.|| / (||y_b|| + eps) in sample order (deterministic)
__global__ void __launch_bounds__(256) rel_l2_finish_kernel(const float* __restrict__ part, int nchunks, int B, float eps,
                                                           float* __restrict__ loss, float* __restrict__ coef) {
    __shared__ float terms[256];
    float acc = 0.f;
    for (int b0 = 0; b0 < B; b0 += 256) {
        const int b = b0 + threadIdx.x;
        float t = 0.f;
        if (b < B) {
            float sn = 0.f, sd = 0.f;
            for (int c = 0; c < nchunks; ++c) { sn += part[((size_t)b * nchunks + c) * 2]; sd += part[((size_t)b * nchunks + c) * 2 + 1]; }
            const float num = sqrtf(sn), den = sqrtf(sd) + eps;
            t = num / den;
            coef[b] = num > 0.f ? 1.f / (num * den) : 0.f;
        }
        terms[threadIdx.x] = t;
        __syncthreads();
        if (threadIdx.x == 0)
            for (int i = 0; i < 256 && b0 + i < B; ++i) acc += terms[i];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = acc;
}

__global__ void __launch_bounds__(256) rel_l2_bwd_kernel(const float* __restrict__ out, const float* __restrict__ y,
                                                        const float* __restrict__ coef, const float* __restrict__ gloss,
                                                        float* __restrict__ gout, long long n, int vec) {
    const long long base = (long long)blockIdx.y * n;
    const float c = __ldg(coef + blockIdx.y) * __ldg(gloss);
    const long long n4 = vec ? n / 4 : 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float4 a = *reinterpret_cast<const float4*>(out + base + 4 * i), b = *reinterpret_cast<const float4*>(y + base + 4 * i);
        *reinterpret_cast<float4*>(gout + base + 4 * i) = make_float4(c * (a.x - b.x), c * (a.y - b.y), c * (a.z - b.z), c * (a.w - b.w));
    }
    if (blockIdx.x == 0 || !vec)
        for (long long i = 4 * n4 + (vec ? 0 : (long long)blockIdx.x * 256) + threadIdx.x; i < n; i += (vec ? 256 : (long long)gridDim.x * 256))
            gout[base + i] = c * (out[base + i] - y[base + i]);
}

}  // namespace fnm

using namespace fnm;

extern "C" {

int fnm_lift_fwd(const float* x, const float* w, const float* b, float* out, int B, int Din, int N1, int N2, int P1,
                 int P2, int C, int ngrid, fnm_stream_t stream) {
    FNM_REQUIRE(x && w && b && out, "lift_fwd: NULL pointer");
    FNM_REQUIRE(B > 0 && N1 > 0 && N2 > 0 && P1 >= N1 && P2 >= N2 && ngrid >= 0 && ngrid <= 2, "lift_fwd: bad extents");
    if (!lift_ok(Din, C, ngrid)) return fail(FNM_ENOTSUP, "lift_fwd: unsupported channel counts");
    LiftArgs a{};
    a.x = x; a.w = w; a.b = b; a.out = out; a.B = B; a.Din = Din; a.N1 = N1; a.N2 = N2; a.P1 = P1; a.P2 = P2; a.C = C;
    a.ngrid = ngrid; a.J = Din + ngrid;
    a.step1 = N1 > 1 ? 1.0f / (float)(N1 - 1) : 0.f; a.step2 = N2 > 1 ? 1.0f / (float)(N2 - 1) : 0.f;
    const int nrows = B * P1;
    const int grid = nrows < kEndBlocks * 8 ? nrows : kEndBlocks * 8;
    {
        LaunchScope ls("lift_fwd", as_stream(stream));
        int threads = 256;
        if (threads < C / 4) threads = (C / 4 + 31) / 32 * 32;              // C <= 512 (lift_ok): at most 128 groups
        if (a.J <= 4) lift_fwd_kernel<4><<<grid, threads, 0, as_stream(stream)>>>(a);
        else lift_fwd_kernel<kLiftMaxJ><<<grid, threads, 0, as_stream(stream)>>>(a);
    }
    return check_launch("lift_fwd");
}

size_t fnm_lift_bwd_workspace_bytes(int C, int J) { return align_up((size_t)kEndBlocks * (J + 1) * C * sizeof(float), 256); }

int fnm_lift_bwd(const float* g, const float* x, float* dw, float* db, int B, int Din, int N1, int N2, int P1, int P2,
                 int C, int ngrid, void* workspace, size_t workspace_bytes, fnm_stream_t stream) {
    FNM_REQUIRE(g && x && workspace, "lift_bwd: NULL pointer");
    if (!lift_ok(Din, C, ngrid) || C > 512) return fail(FNM_ENOTSUP, "lift_bwd: unsupported channel counts");
    const int J = Din + ngrid;
    if (workspace_bytes < fnm_lift_bwd_workspace_bytes(C, J)) return fail(FNM_EWORKSPACE, "lift_bwd: workspace too small");
    LiftArgs a{};
    a.g = g; a.x = x; a.partial = static_cast<float*>(workspace);
    a.B = B; a.Din = Din; a.N1 = N1; a.N2 = N2; a.P1 = P1; a.P2 = P2; a.C = C; a.ngrid = ngrid; a.J = J;
    a.step1 = N1 > 1 ? 1.0f / (float)(N1 - 1) : 0.f; a.step2 = N2 > 1 ? 1.0f / (float)(N2 - 1) : 0.f;
    const int C4 = C / 4;
    int threads = 256;
    if (threads < C4) threads = (C4 + 31) / 32 * 32;
    const int slots = threads / C4;
    const size_t smem = (size_t)slots * (J + 1) * C * sizeof(float);
    if (smem > 100 * 1024 || threads > 256) return fail(FNM_ENOTSUP, "lift_bwd: reduction tile too large");
    cudaStream_t st = as_stream(stream);
    static std::atomic<uint64_t> attr_set{0};
    int attr_set_dev = 0;
    if (once_needed(attr_set, attr_set_dev)) {
        cudaFuncSetAttribute(lift_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(lift_bwd_kernel<kLiftMaxJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        once_done(attr_set, attr_set_dev);
    }
    {
        LaunchScope ls("lift_bwd", st);
        if (J <= 4) lift_bwd_kernel<4><<<kEndBlocks, threads, smem, st>>>(a);
        else lift_bwd_kernel<kLiftMaxJ><<<kEndBlocks, threads, smem, st>>>(a);
    }
    FNM_TRY(check_launch("lift_bwd"));
    {
        LaunchScope ls("lift_bwd_reduce", st);
        lift_bwd_reduce<<<((J + 1) * C + 31) / 32, dim3(32, 32), 0, st>>>(a.partial, kEndBlocks, J, C, dw, db);
    }
    return check_launch("lift_bwd_reduce");
}

int fnm_mlp_head_fwd(const float* h1, const float* w2, const float* b2, float* out, int64_t npos, int H, int D,
                     fnm_stream_t stream) {
    FNM_REQUIRE(h1 && w2 && out, "mlp_head_fwd: NULL pointer");
    if (H % 4 != 0 || H < 4 || H > 128 * kHeadMaxH4 || D < 1 || D > kHeadMaxD) return fail(FNM_ENOTSUP, "mlp_head_fwd: unsupported widths");
    HeadArgs a{};
    a.h1 = h1; a.w2 = w2; a.b2 = b2; a.out = out; a.npos = npos; a.H = H; a.D = D;
    if (npos <= 0) return FNM_OK;
    const long long want = (npos + 63) / 64;                          // 8 warps x 8 positions per block iteration
    const int grid = (int)(want < kEndBlocks * 4 ? want : kEndBlocks * 4);
    {
        LaunchScope ls("mlp_head_fwd", as_stream(stream));
        switch (D) {
            case 1: head_fwd_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(a); break;
            case 2: head_fwd_kernel<2><<<grid, 256, 0, as_stream(stream)>>>(a); break;
            case 3: head_fwd_kernel<3><<<grid, 256, 0, as_stream(stream)>>>(a); break;
            default: head_fwd_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(a); break;
        }
    }
    return check_launch("mlp_head_fwd");
}

size_t fnm_mlp_head_bwd_workspace_bytes(int H, int D) { return align_up((size_t)kEndBlocks * (D * H + D) * sizeof(float), 256); }

int fnm_mlp_head_bwd(const float* h1, const float* g_out, const float* w2, float* g_h1, float* dw2, float* db2,
                     int64_t npos, int H, int D, void* workspace, size_t workspace_bytes, fnm_stream_t stream) {
    FNM_REQUIRE(h1 && g_out && w2 && g_h1 && workspace, "mlp_head_bwd: NULL pointer");
    if (H % 4 != 0 || H < 4 || H > 128 * kHeadMaxH4 || D < 1 || D > kHeadMaxD) return fail(FNM_ENOTSUP, "mlp_head_bwd: unsupported widths");
    if (workspace_bytes < fnm_mlp_head_bwd_workspace_bytes(H, D)) return fail(FNM_EWORKSPACE, "mlp_head_bwd: workspace too small");
    HeadArgs a{};
    a.h1 = h1; a.g_out = g_out; a.w2 = w2; a.g_h1 = g_h1; a.partial = static_cast<float*>(workspace);
    a.npos = npos; a.H = H; a.D = D;
    cudaStream_t st = as_stream(stream);
    const size_t smem = (size_t)8 * (D * H + D) * sizeof(float);
    static std::atomic<uint64_t> attr_set{0};
    int attr_set_dev = 0;
    if (once_needed(attr_set, attr_set_dev)) {
        cudaFuncSetAttribute(head_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(head_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(head_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(head_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        once_done(attr_set, attr_set_dev);
    }
    {
        LaunchScope ls("mlp_head_bwd", st);
        switch (D) {
            case 1: head_bwd_kernel<1><<<kEndBlocks, 256, smem, st>>>(a); break;
            case 2: head_bwd_kernel<2><<<kEndBlocks, 256, smem, st>>>(a); break;
            case 3: head_bwd_kernel<3><<<kEndBlocks, 256, smem, st>>>(a); break;
            default: head_bwd_kernel<4><<<kEndBlocks, 256, smem, st>>>(a); break;
        }
    }
    FNM_TRY(check_launch("mlp_head_bwd"));
    {
        LaunchScope ls("mlp_head_bwd_reduce", st);
        head_bwd_reduce<<<(D * H + D + 31) / 32, dim3(32, 32), 0, st>>>(a.partial, kEndBlocks, H, D, dw2, db2);
    }
    return check_launch("mlp_head_bwd_reduce");
}

size_t fnm_wsum_gelu_workspace_bytes(int B, int H) { return align_up((size_t)B * kWsumChunks * H * sizeof(float), 256); }

int fnm_wsum_gelu_fwd(const float* h1, const float* wx, const float* wy, float* S, int B, int P1, int P2, int H,
                      void* workspace, size_t workspace_bytes, fnm_stream_t stream) {
    FNM_REQUIRE(h1 && wx && wy && S && workspace, "wsum_gelu_fwd: NULL pointer");
    if (H % 4 != 0 || H < 4 || H > 128 || B < 1 || B > 65535) return fail(FNM_ENOTSUP, "wsum_gelu_fwd: unsupported widths");
    if (workspace_bytes < fnm_wsum_gelu_workspace_bytes(B, H)) return fail(FNM_EWORKSPACE, "wsum_gelu_fwd: workspace too small");
    WsumArgs a{};
    a.h1 = h1; a.wx = wx; a.wy = wy; a.partial = static_cast<float*>(workspace);
    a.B = B; a.P1 = P1; a.P2 = P2; a.H = H; a.nchunk = kWsumChunks;
    cudaStream_t st = as_stream(stream);
    {
        LaunchScope ls("wsum_gelu_fwd", st);
        wsum_fwd_kernel<<<dim3(kWsumChunks, B), 256, 8 * H * sizeof(float), st>>>(a);
    }
    FNM_TRY(check_launch("wsum_gelu_fwd"));
    {
        LaunchScope ls("wsum_gelu_reduce", st);
        wsum_reduce_kernel<<<(B * H + 255) / 256, 256, 0, st>>>(a.partial, kWsumChunks, B * H, H, S);
    }
    return check_launch("wsum_gelu_reduce");
}

int fnm_wsum_gelu_bwd(const float* h1, const float* wx, const float* wy, const float* gS, float* g_h1, int B, int P1,
                      int P2, int H, fnm_stream_t stream) {
    FNM_REQUIRE(h1 && wx && wy && gS && g_h1, "wsum_gelu_bwd: NULL pointer");
    if (H % 4 != 0 || H < 4 || B < 1 || B > 65535) return fail(FNM_ENOTSUP, "wsum_gelu_bwd: unsupported widths");
    WsumArgs a{};
    a.h1 = h1; a.wx = wx; a.wy = wy; a.gS = gS; a.g_h1 = g_h1; a.B = B; a.P1 = P1; a.P2 = P2; a.H = H;
    const long long per_b = (long long)P1 * P2 * (H / 4);
    int gx = (int)((per_b + 255) / 256);
    const int cap = (kEndBlocks * 8 + B - 1) / B;
    if (gx > cap) gx = cap < 1 ? 1 : cap;
    {
        LaunchScope ls("wsum_gelu_bwd", as_stream(stream));
        wsum_bwd_kernel<<<dim3(gx, B), 256, 0, as_stream(stream)>>>(a);
    }
    return check_launch("wsum_gelu_bwd");
}

size_t fnm_rel_l2_workspace_bytes(int B) { return align_up((size_t)(B > 0 ? B : 0) * kLossChunks * 2 * sizeof(float), 256); }

int fnm_rel_l2_fwd(const float* out, const float* y, float* loss, float* coef, int B, int64_t n, float eps, void* workspace,
                   size_t workspace_bytes, fnm_stream_t stream) {
    FNM_REQUIRE(out && y && loss && coef && workspace, "rel_l2_fwd: NULL pointer");
    FNM_REQUIRE(B > 0 && B <= 65535 && n > 0, "rel_l2_fwd: bad extents");
    if (workspace_bytes < fnm_rel_l2_workspace_bytes(B)) return fail(FNM_EWORKSPACE, "rel_l2_fwd: workspace too small");
    if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(y)) & 3) return fail(FNM_ENOTSUP, "rel_l2_fwd: unaligned tensors");
    // float4 path when every sample row starts on a 16-byte boundary, scalar path otherwise
    const int vec = !(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(y)) & 15) || (n % 4 != 0 && B > 1));
    cudaStream_t st = as_stream(stream);
    const long long want = ((vec ? n / 4 : n) + 255) / 256;
    const int chunks = (int)(want < 1 ? 1 : (want < kLossChunks ? want : kLossChunks));
    float* part = static_cast<float*>(workspace);
    {
        LaunchScope ls("rel_l2_partial", st);
        rel_l2_partial_kernel<<<dim3(chunks, B), 256, 0, st>>>(out, y, part, n, vec);
    }
    FNM_TRY(check_launch("rel_l2_partial"));
    {
        LaunchScope ls("rel_l2_finish", st);
        rel_l2_finish_kernel<<<1, 256, 0, st>>>(part, chunks, B, eps, loss, coef);
    }
    return check_launch("rel_l2_finish");
}

int fnm_rel_l2_bwd(const float* out, const float* y, const float* coef, const float* gloss, float* gout, int B, int64_t n,
                   fnm_stream_t stream) {
    FNM_REQUIRE(out && y && coef && gloss && gout, "rel_l2_bwd: NULL pointer");
    FNM_REQUIRE(B > 0 && B <= 65535 && n > 0, "rel_l2_bwd: bad extents");
    if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gout)) & 3)
        return fail(FNM_ENOTSUP, "rel_l2_bwd: unaligned tensors");
    const int vec = !(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gout)) & 15) ||
                      (n % 4 != 0 && B > 1));
    cudaStream_t st = as_stream(stream);
    const long long want = ((vec ? n / 4 : n) + 255) / 256;
    const int gx = (int)(want < 1 ? 1 : (want < 148 * 8 ? want : 148 * 8));
    {
        LaunchScope ls("rel_l2_bwd", st);
        rel_l2_bwd_kernel<<<dim3(gx, B), 256, 0, st>>>(out, y, coef, gloss, gout, n, vec);
    }
    return check_launch("rel_l2_bwd");
}

}  // extern "C"
