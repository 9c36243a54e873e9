// fp32 SIMT path of libfnm_b200: one tiled batched-GEMM engine, instantiated with small "problem"
// functors that describe where each operand element lives.  Every stage of the Fourier layer is a
// (batched) contraction against either a twiddle table, a weight, or an activation tile:
//
//   ydft            T[row]     = tab(LY x P2)          . f(X[row])(P2 x C)          (rfft along y, truncated)
//   xdft / xsyn     per sample : table(2R x 2P1) . T   /  table(2S1 x 2R) . Oh     (C2C along x, truncated)
//   mix*            per mode   : complex (B x Ci).(Ci x Co) as a real-embedded GEMM  (compl_mul)
//   pointwise_ysyn  per row    : [f(X) | by](S2 x (K+LY)) . [Wk ; Z[row]]           (1x1 conv + irfft along y)
//   conv_wgrad      split-K    : g^T . [f(X) | 1]                                    (conv dW, db)
//   lfunc* / ldec*  per mode or split over modes (F2V / V2F ends)
//
// This is the general path (any C, any grid, exact fp32 FMA).  The tcgen05 kernels in fnm_tc.cu
// take over the heavy spatial passes for the shapes they support.
#include "fnm_common.cuh"

#include <cstring>

namespace fnm {

// ------------------------------------------------------------------------------------------------
// engine: C[batch](M x N) = A[batch](M x K) . B[batch](K x N), 4x4 register tiles, BK = 16,
// register-prefetched double-buffered shared tiles.  Problem functor P supplies
//   int M, N;  kbeg(batch), kend(batch);  a(batch,m,k);  b(batch,k,n);  store(batch,m,n,v)
//   static constexpr bool A_MFAST / B_NFAST : which index is contiguous in memory (coalescing)
// ------------------------------------------------------------------------------------------------
constexpr int BK = 16;

template <int BM, int BN, class P>
__global__ void __launch_bounds__((BM / 4) * (BN / 4)) gemm_engine(const P p, int tiles_m, int tiles_n) {
    constexpr int NT = (BM / 4) * (BN / 4);
    constexpr int TXN = BN / 4;
    constexpr int A_PER = (BM * BK + NT - 1) / NT;
    constexpr int B_PER = (BK * BN + NT - 1) / NT;
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int tiles = tiles_m * tiles_n;
    const int batch = blockIdx.x / tiles;
    const int t = blockIdx.x - batch * tiles;
    const int m0 = (t / tiles_n) * BM;
    const int n0 = (t % tiles_n) * BN;
    const int M = p.M, N = p.N;
    const int k_begin = p.kbeg(batch), k_end = p.kend(batch);

    const int tx = tid % TXN, ty = tid / TXN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float ra[A_PER], rb[B_PER];

    auto fetch = [&](int k0) {
#pragma unroll
        for (int j = 0; j < A_PER; ++j) {
            const int e = tid + j * NT;
            int m, k;
            if (P::A_MFAST) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
            float v = 0.f;
            if (e < BM * BK && (m0 + m) < M && (k0 + k) < k_end) v = p.a(batch, m0 + m, k0 + k);
            ra[j] = v;
        }
#pragma unroll
        for (int j = 0; j < B_PER; ++j) {
            const int e = tid + j * NT;
            int n, k;
            if (P::B_NFAST) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
            float v = 0.f;
            if (e < BK * BN && (n0 + n) < N && (k0 + k) < k_end) v = p.b(batch, k0 + k, n0 + n);
            rb[j] = v;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int j = 0; j < A_PER; ++j) {
            const int e = tid + j * NT;
            int m, k;
            if (P::A_MFAST) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
            if (e < BM * BK) As[buf][k][m] = ra[j];
        }
#pragma unroll
        for (int j = 0; j < B_PER; ++j) {
            const int e = tid + j * NT;
            int n, k;
            if (P::B_NFAST) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
            if (e < BK * BN) Bs[buf][k][n] = rb[j];
        }
    };

    int buf = 0;
    if (k_begin < k_end) {
        fetch(k_begin);
        stash(0);
    }
    __syncthreads();
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool more = (k0 + BK) < k_end;
        if (more) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            stash(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) p.store(batch, m, n, acc[i][j]);
        }
    }
}

// Profile name of a stage when the fp32 SIMT engine serves it: stages that also have a tcgen05 kernel get a ".simt" suffix, so
// that a fallback shows up as its own line in bench.py's per-kernel list instead of hiding behind the stage name.
static const char* simt_profile_name(const char* what) {
    static const char* const names[][2] = {
        {"ydft", "ydft.simt"}, {"xdft", "xdft.simt"}, {"xsyn", "xsyn.simt"}, {"mix_fwd", "mix_fwd.simt"},
        {"mix_bwd_x", "mix_bwd_x.simt"}, {"mix_bwd_w", "mix_bwd_w.simt"}, {"pointwise_ysyn", "pointwise_ysyn.simt"},
        {"conv_wgrad", "conv_wgrad.simt"}};
    for (const auto& n : names)
        if (strcmp(what, n[0]) == 0) return n[1];
    return what;
}

template <int BM, int BN, class P>
static int launch_engine(const P& p, int64_t nbatch, cudaStream_t st, const char* what) {
    const int tm = (p.M + BM - 1) / BM, tn = (p.N + BN - 1) / BN;
    const int64_t blocks = nbatch * tm * tn;
    if (blocks <= 0) return FNM_OK;
    if (blocks > 0x7fffffffLL) return fail(FNM_EINVAL, "%s: grid too large (%lld blocks)", what, (long long)blocks);
    {
        LaunchScope ls(simt_profile_name(what), st);
        gemm_engine<BM, BN, P><<<(unsigned)blocks, (BM / 4) * (BN / 4), 0, st>>>(p, tm, tn);
    }
    return check_launch(what);
}

// pick a tile shape from the runtime extents (keeps padding waste small for C = 32 / S2 = 72 etc.)
template <class P>
static int launch_auto(const P& p, int64_t nbatch, cudaStream_t st, const char* what) {
    const bool m72 = (p.M % 72 == 0) || (p.M > 64 && p.M <= 72);
    const bool small_m = p.M <= 32;
    const bool small_n = p.N <= 32;
    if (small_m && small_n) return launch_engine<32, 32, P>(p, nbatch, st, what);
    if (small_m) return launch_engine<32, 64, P>(p, nbatch, st, what);
    if (m72 && small_n) return launch_engine<72, 32, P>(p, nbatch, st, what);
    if (m72) return launch_engine<72, 64, P>(p, nbatch, st, what);
    if (small_n) return launch_engine<64, 32, P>(p, nbatch, st, what);
    return launch_engine<64, 64, P>(p, nbatch, st, what);
}

struct FullK {
    int K;
    __device__ __forceinline__ int kbeg(int) const { return 0; }
    __device__ __forceinline__ int kend(int) const { return K; }
};

// complex weight element W[i][o][r][l] of a (Ci,Co,rows_per_w,NY) tensor pair (reference layout)
struct WeightRef {
    const float2* w0;
    const float2* w1;
    int rpw, NY, Co, Ci, mm;                   // mm: the tensors are held mode-major, (rows, NY, Ci, Co)
    __device__ __forceinline__ size_t index(int i, int o, int mode, bool& second) const {
        const int r = mode / NY, l = mode - r * NY;
        second = r >= rpw;
        const int rr = second ? r - rpw : r;
        if (mm) return (((size_t)rr * NY + l) * Ci + i) * Co + o;
        return (((size_t)i * Co + o) * rpw + rr) * NY + l;
    }
    __device__ __forceinline__ float2 load(int i, int o, int mode) const {
        bool second;
        const size_t idx = index(i, o, mode, second);
        return __ldg((second ? w1 : w0) + idx);
    }
};

// ------------------------------------------------------------------------------------------------
// stage functors
// ------------------------------------------------------------------------------------------------
struct YdftP : FullK {            // batch = row (b,x)
    const float* x; const float* tab; float* T;
    int M, N, P2, act;
    static constexpr bool A_MFAST = false, B_NFAST = true;
    __device__ __forceinline__ float a(int, int m, int k) const { return __ldg(tab + (size_t)m * P2 + k); }
    __device__ __forceinline__ float b(int row, int k, int n) const {
        const float v = __ldg(x + ((size_t)row * P2 + k) * N + n);
        return act ? gelu_f(v) : v;
    }
    __device__ __forceinline__ void store(int row, int m, int n, float v) const { T[((size_t)row * M + m) * N + n] = v; }
};

struct XdftP : FullK {            // batch = sample b;  M = 2R, K = 2P1, N = NY*C
    const float* T; const float* tab; float* Xh;
    int M, N, R, NY, C, B;
    static constexpr bool A_MFAST = false, B_NFAST = true;
    __device__ __forceinline__ float a(int, int m, int k) const { return __ldg(tab + (size_t)m * K + k); }
    __device__ __forceinline__ float b(int bb, int k, int n) const { return __ldg(T + ((size_t)bb * K + k) * N + n); }
    __device__ __forceinline__ void store(int bb, int m, int n, float v) const {
        const int ro = m / R, r = m - ro * R, l = n / C, c = n - l * C;
        Xh[((((size_t)r * NY + l) * 2 + ro) * B + bb) * C + c] = v;
    }
};

struct XsynP : FullK {            // batch = sample b;  M = 2S1, K = 2R, N = NY*C
    const float* Oh; const float* tab; float* Z;
    int M, N, R, NY, C, B;
    static constexpr bool A_MFAST = false, B_NFAST = true;
    __device__ __forceinline__ float a(int, int m, int k) const { return __ldg(tab + (size_t)m * K + k); }
    __device__ __forceinline__ float b(int bb, int k, int n) const {
        const int ri = k / R, r = k - ri * R, l = n / C, c = n - l * C;
        return __ldg(Oh + ((((size_t)r * NY + l) * 2 + ri) * B + bb) * C + c);
    }
    __device__ __forceinline__ void store(int bb, int m, int n, float v) const { Z[((size_t)bb * M + m) * N + n] = v; }
};

// compl_mul forward: [Or Oi] = [Xr Xi] . [[Wr Wi], [-Wi Wr]]      batch = mode; M = B, K = 2Ci, N = 2Co
struct MixFwdP : FullK {
    const float* Xh; float* Oh; WeightRef w;
    int M, N, Ci, Co;
    static constexpr bool A_MFAST = false, B_NFAST = true;
    __device__ __forceinline__ float a(int mode, int b, int k) const {
        const int ri = k / Ci, i = k - ri * Ci;
        return __ldg(Xh + (((size_t)mode * 2 + ri) * M + b) * Ci + i);
    }
    __device__ __forceinline__ float b(int mode, int k, int n) const {
        const int ri = k / Ci, i = k - ri * Ci, ro = n / Co, o = n - ro * Co;
        const float2 v = w.load(i, o, mode);
        return ri == ro ? v.x : (ri == 0 ? v.y : -v.y);
    }
    __device__ __forceinline__ void store(int mode, int b, int n, float v) const {
        const int ro = n / Co, o = n - ro * Co;
        Oh[(((size_t)mode * 2 + ro) * M + b) * Co + o] = v;
    }
};

// gxh = sum_o gh conj(W):  [Gxr Gxi] = [Gr Gi] . [[Wr -Wi], [Wi Wr]]    M = B, K = 2Co, N = 2Ci
struct MixBwdXP : FullK {
    const float* Gh; float* Gxh; WeightRef w;
    int M, N, Ci, Co;
    static constexpr bool A_MFAST = false, B_NFAST = true;
    __device__ __forceinline__ float a(int mode, int b, int k) const {
        const int ro = k / Co, o = k - ro * Co;
        return __ldg(Gh + (((size_t)mode * 2 + ro) * M + b) * Co + o);
    }
    __device__ __forceinline__ float b(int mode, int k, int n) const {
        const int ro = k / Co, o = k - ro * Co, ri = n / Ci, i = n - ri * Ci;
        const float2 v = w.load(i, o, mode);
        return ri == ro ? v.x : (ro == 1 ? v.y : -v.y);
    }
    __device__ __forceinline__ void store(int mode, int b, int n, float v) const {
        const int ri = n / Ci, i = n - ri * Ci;
        Gxh[(((size_t)mode * 2 + ri) * M + b) * Ci + i] = v;
    }
};

// dW = sum_b conj(xh) gh:  dWr = xr gr + xi gi, dWi = xr gi - xi gr      M = Ci, K = 2B, N = 2Co
struct MixBwdWP : FullK {
    const float* Xh; const float* Gh; float2* dw0; float2* dw1;
    int M, N, B, Co, rpw, NY, mm;
    static constexpr bool A_MFAST = true, B_NFAST = true;
    __device__ __forceinline__ float a(int mode, int i, int k) const {
        const int ri = k / B, b = k - ri * B;
        return __ldg(Xh + (((size_t)mode * 2 + ri) * B + b) * M + i);
    }
    __device__ __forceinline__ float b(int mode, int k, int n) const {
        const int ri = k / B, b = k - ri * B, ro = n / Co, o = n - ro * Co;
        const float v = __ldg(Gh + (((size_t)mode * 2 + (ri ^ ro)) * B + b) * Co + o);
        return (ri == 1 && ro == 1) ? -v : v;
    }
    __device__ __forceinline__ void store(int mode, int i, int n, float v) const {
        const int ro = n / Co, o = n - ro * Co;
        const int r = mode / NY, l = mode - r * NY;
        const bool second = r >= rpw;
        const int rr = second ? r - rpw : r;
        const size_t idx = mm ? (((size_t)rr * NY + l) * M + i) * Co + o : (((size_t)i * Co + o) * rpw + rr) * NY + l;
        reinterpret_cast<float*>((second ? dw1 : dw0) + idx)[ro] = v;
    }
};

// fused pointwise + y-synthesis (+ bias, + GELU' multiply).  batch = row; M = S2, K = Kc + LY
struct PointwiseYsynP : FullK {
    const float* x; const float* wc; const float* bias; const float* Z; const float* by; const float* mul;
    float* out; float* out_act;
    int M, N, Kc, LY, act, wc_is_kn, flags;
    static constexpr bool A_MFAST = false, B_NFAST = true;
    __device__ __forceinline__ float a(int row, int y, int k) const {
        if (k < Kc) {
            const float v = __ldg(x + ((size_t)row * M + y) * Kc + k);
            return act ? gelu_f(v) : v;
        }
        return __ldg(by + (size_t)y * LY + (k - Kc));
    }
    __device__ __forceinline__ float b(int row, int k, int n) const {
        if (k < Kc) return wc_is_kn ? __ldg(wc + (size_t)k * N + n) : __ldg(wc + (size_t)n * Kc + k);
        return __ldg(Z + ((size_t)row * LY + (k - Kc)) * N + n);
    }
    __device__ __forceinline__ void store(int row, int y, int n, float v) const {
        const size_t idx = ((size_t)row * M + y) * N + n;
        if (bias) v += __ldg(bias + n);
        if (mul) v *= (flags & FNM_PW_MUL_IS_GRAD) ? __ldg(mul + idx) : gelu_grad_f(__ldg(mul + idx));
        out[idx] = (out_act && (flags & FNM_PW_OUT_IS_GRAD)) ? gelu_grad_f(v) : v;
        if (out_act) out_act[idx] = gelu_f(v);
    }
};

// conv weight/bias gradient, split over positions: partial[s][o][n], n < Ci -> dW, n == Ci -> db
struct ConvWgradP {
    const float* g; const float* x; float* partial;
    int M, N, Ci, act;
    int64_t npos; int chunk;
    static constexpr bool A_MFAST = true, B_NFAST = true;
    __device__ __forceinline__ int kbeg(int) const { return 0; }
    __device__ __forceinline__ int kend(int s) const {
        const int64_t rem = npos - (int64_t)s * chunk;
        return rem < chunk ? (int)rem : chunk;
    }
    __device__ __forceinline__ float a(int s, int o, int k) const { return __ldg(g + ((size_t)s * chunk + k) * M + o); }
    __device__ __forceinline__ float b(int s, int k, int n) const {
        if (n >= Ci) return 1.f;
        const float v = __ldg(x + ((size_t)s * chunk + k) * Ci + n);
        return act ? gelu_f(v) : v;
    }
    __device__ __forceinline__ void store(int s, int o, int n, float v) const { partial[((size_t)s * M + o) * N + n] = v; }
};

__global__ void conv_wgrad_reduce(const float* __restrict__ partial, int nsplit, int Co, int Ci,
                                  float* __restrict__ dwc, float* __restrict__ dbc) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = Ci + 1;
    if (j >= Co * N) return;
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += partial[(size_t)k * Co * N + j];
    const int o = j / N, n = j - o * N;
    if (n < Ci) { if (dwc) dwc[o * Ci + n] = s; }
    else if (dbc) dbc[o] = s;
}

__global__ void sum_partials(const float* __restrict__ partial, int nsplit, int count, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += partial[(size_t)k * count + j];
    out[j] = s;
}

// ---- F2V contraction: out[b,o] = sum_{mode,i} cc[mode] (xr wr - xi wi); split over groups of modes
struct LfuncFwdP {
    const float* Xh; const float* cc; float* partial; WeightRef w;
    int M, N, Ci, nmodes, per;           // M = B, N = Co
    static constexpr bool A_MFAST = false, B_NFAST = true;
    __device__ __forceinline__ int kbeg(int s) const { return s * per * 2 * Ci; }
    __device__ __forceinline__ int kend(int s) const {
        const int m1 = (s + 1) * per < nmodes ? (s + 1) * per : nmodes;
        return m1 * 2 * Ci;
    }
    __device__ __forceinline__ float a(int, int b, int k) const {
        const int mode = k / (2 * Ci), rem = k - mode * 2 * Ci, ri = rem / Ci, i = rem - ri * Ci;
        return __ldg(Xh + (((size_t)mode * 2 + ri) * M + b) * Ci + i);
    }
    __device__ __forceinline__ float b(int, int k, int o) const {
        const int mode = k / (2 * Ci), rem = k - mode * 2 * Ci, ri = rem / Ci, i = rem - ri * Ci;
        const float2 v = w.load(i, o, mode);
        return __ldg(cc + mode) * (ri == 0 ? v.x : -v.y);
    }
    __device__ __forceinline__ void store(int s, int b, int o, float v) const { partial[((size_t)s * M + b) * N + o] = v; }
};

// gxh[mode][ri][b][i] = cc * sum_o g[b,o] * (ri ? -wi : wr)            batch = mode; M = B, K = Co, N = 2Ci
struct LfuncBwdXP : FullK {
    const float* g; const float* cc; float* Gxh; WeightRef w;
    int M, N, Ci;
    static constexpr bool A_MFAST = false, B_NFAST = true;
    __device__ __forceinline__ float a(int, int b, int o) const { return __ldg(g + (size_t)b * K + o); }
    __device__ __forceinline__ float b(int mode, int o, int n) const {
        const int ri = n / Ci, i = n - ri * Ci;
        const float2 v = w.load(i, o, mode);
        return __ldg(cc + mode) * (ri == 0 ? v.x : -v.y);
    }
    __device__ __forceinline__ void store(int mode, int b, int n, float v) const {
        const int ri = n / Ci, i = n - ri * Ci;
        Gxh[(((size_t)mode * 2 + ri) * M + b) * Ci + i] = v;
    }
};

// dW[i,o,mode].ro = cc * sum_b g[b,o] * (ro ? -xi : xr)[b,i]           batch = mode; M = 2Ci, K = B, N = Co
struct LfuncBwdWP : FullK {
    const float* g; const float* Xh; const float* cc; float2* dw;
    int M, N, Ci, NYR, mm;               // NYR = R*NY (modes per (i,o)); mm: dw is held mode-major
    static constexpr bool A_MFAST = true, B_NFAST = true;
    __device__ __forceinline__ float a(int mode, int m, int b) const {
        const int ro = m / Ci, i = m - ro * Ci;
        const float v = __ldg(Xh + (((size_t)mode * 2 + ro) * K + b) * Ci + i);
        return ro ? -v : v;
    }
    __device__ __forceinline__ float b(int, int b, int o) const { return __ldg(g + (size_t)b * N + o); }
    __device__ __forceinline__ void store(int mode, int m, int o, float v) const {
        const int ro = m / Ci, i = m - ro * Ci;
        float* dst = reinterpret_cast<float*>(dw + (mm ? ((size_t)mode * Ci + i) * N + o : ((size_t)i * N + o) * NYR + mode));
        dst[ro] = __ldg(cc + mode) * v;
    }
};

// ---- V2F mixing: Oh[mode][ro][b][o] = sum_i v[b,i] W{ro}[i,o,mode]     batch = mode; M = B, K = Ci, N = 2Co
struct LdecFwdP : FullK {
    const float* v; float* Oh; WeightRef w;
    int M, N, Co;
    static constexpr bool A_MFAST = false, B_NFAST = true;
    __device__ __forceinline__ float a(int, int b, int i) const { return __ldg(v + (size_t)b * K + i); }
    __device__ __forceinline__ float b(int mode, int i, int n) const {
        const int ro = n / Co, o = n - ro * Co;
        const float2 e = w.load(i, o, mode);
        return ro ? e.y : e.x;
    }
    __device__ __forceinline__ void store(int mode, int b, int n, float val) const {
        const int ro = n / Co, o = n - ro * Co;
        Oh[(((size_t)mode * 2 + ro) * M + b) * Co + o] = val;
    }
};

// dW[i,o,mode].ro = sum_b v[b,i] gh{ro}[b,o]                            batch = mode; M = Ci, K = B, N = 2Co
struct LdecBwdWP : FullK {
    const float* v; const float* Gh; float2* dw;
    int M, N, Co, NYR, mm;
    static constexpr bool A_MFAST = true, B_NFAST = true;
    __device__ __forceinline__ float a(int, int i, int b) const { return __ldg(v + (size_t)b * M + i); }
    __device__ __forceinline__ float b(int mode, int b, int n) const {
        const int ro = n / Co, o = n - ro * Co;
        return __ldg(Gh + (((size_t)mode * 2 + ro) * K + b) * Co + o);
    }
    __device__ __forceinline__ void store(int mode, int i, int n, float val) const {
        const int ro = n / Co, o = n - ro * Co;
        float* dst = reinterpret_cast<float*>(dw + (mm ? ((size_t)mode * M + i) * Co + o : ((size_t)i * Co + o) * NYR + mode));
        dst[ro] = val;
    }
};

// dv[b,i] = sum_{mode,o} (gr wr + gi wi)     split over groups of modes; M = B, N = Ci, k = (mode, ro, o)
struct LdecBwdVP {
    const float* Gh; float* partial; WeightRef w;
    int M, N, Co, nmodes, per;
    static constexpr bool A_MFAST = false, B_NFAST = false;
    __device__ __forceinline__ int kbeg(int s) const { return s * per * 2 * Co; }
    __device__ __forceinline__ int kend(int s) const {
        const int m1 = (s + 1) * per < nmodes ? (s + 1) * per : nmodes;
        return m1 * 2 * Co;
    }
    __device__ __forceinline__ float a(int, int b, int k) const {
        const int mode = k / (2 * Co), rem = k - mode * 2 * Co, ro = rem / Co, o = rem - ro * Co;
        return __ldg(Gh + (((size_t)mode * 2 + ro) * M + b) * Co + o);
    }
    __device__ __forceinline__ float b(int, int k, int i) const {
        const int mode = k / (2 * Co), rem = k - mode * 2 * Co, ro = rem / Co, o = rem - ro * Co;
        const float2 e = w.load(i, o, mode);
        return ro ? e.y : e.x;
    }
    __device__ __forceinline__ void store(int s, int b, int i, float v) const { partial[((size_t)s * M + b) * N + i] = v; }
};

// ------------------------------------------------------------------------------------------------
// host-side stage launchers (internal C++ API; the extern "C" wrappers live in fnm_api.cu)
// ------------------------------------------------------------------------------------------------
int simt_ydft(const float* x, const float* tab, float* T, int64_t rows, int P2, int LY, int C, int act, cudaStream_t st) {
    YdftP p;
    p.K = P2; p.x = x; p.tab = tab; p.T = T; p.M = LY; p.N = C; p.P2 = P2; p.act = act;
    return launch_auto(p, rows, st, "ydft");
}

int simt_xdft(const float* T, const float* ex, float* Xh, int B, int P1, int R, int NY, int C, cudaStream_t st) {
    XdftP p;
    p.K = 2 * P1; p.T = T; p.tab = ex; p.Xh = Xh; p.M = 2 * R; p.N = NY * C; p.R = R; p.NY = NY; p.C = C; p.B = B;
    return launch_auto(p, B, st, "xdft");
}

int simt_xsyn(const float* Oh, const float* fx, float* Z, int B, int S1, int R, int NY, int C, cudaStream_t st) {
    XsynP p;
    p.K = 2 * R; p.Oh = Oh; p.tab = fx; p.Z = Z; p.M = 2 * S1; p.N = NY * C; p.R = R; p.NY = NY; p.C = C; p.B = B;
    return launch_auto(p, B, st, "xsyn");
}

// ---- thin x-transforms: 1-D models have P1 = S1 = R = 1, so xdft / xsyn contract over TWO values (re, im) with a 2 x 2
// table and otherwise only change the layout between T / Z ([b][x, ri][ky][c]) and the mode-major spectra
// ([kx][ky][ri][b][c]).  On the table-GEMM kernel such a launch costs its fixed 12-20 us (TMEM, barriers, table
// staging) for half a megabyte of data -- 16 launches, a fifth of a config-2 step.  Here: one thread per (sample, four
// channels of one mode column), K <= 8 loads and M <= 8 stores of 16 bytes, fp32 FMA (exact to rounding).
constexpr int kThinMax = 8;
struct ThinArgs {
    const float* in; const float* tab; float* out;
    int B, K, M, R, NY, C;                     // K / M: contracted / produced table extent (2 P1 or 2 R, 2 R or 2 S1)
    int nsum; long long sum_stride;            // xdft: the input is the sum of nsum partial transforms (segmented 1-D y-DFT)
};

template <bool SYN>
__global__ void __launch_bounds__(256) thin_xform_kernel(const ThinArgs a) {
    const int C4 = a.C >> 2, n4 = a.NY * C4;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)a.B * n4) return;
    const int b = (int)(i / n4), r4 = (int)(i - (long long)b * n4), l = r4 / C4, c = (r4 - l * C4) * 4;
    const size_t N = (size_t)a.NY * a.C;
    // spectra element (k or m = ri * R + r) of this thread: Xh / Oh[r][l][ri][b][c]
    auto spec = [&](int q) { const int ri = q / a.R, r = q - ri * a.R; return ((((size_t)r * a.NY + l) * 2 + ri) * a.B + b) * a.C + c; };
    float4 v[kThinMax];
#pragma unroll
    for (int k = 0; k < kThinMax; ++k)
        if (k < a.K) {
            const float* src = a.in + (SYN ? spec(k) : ((size_t)b * a.K + k) * N + (size_t)l * a.C + c);
            v[k] = __ldg(reinterpret_cast<const float4*>(src));
            for (int s = 1; s < a.nsum; ++s) {                      // fixed order: deterministic
                const float4 w = __ldg(reinterpret_cast<const float4*>(src + (size_t)s * a.sum_stride));
                v[k].x += w.x; v[k].y += w.y; v[k].z += w.z; v[k].w += w.w;
            }
        }
#pragma unroll
    for (int m = 0; m < kThinMax; ++m) {
        if (m < a.M) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < kThinMax; ++k)
                if (k < a.K) {
                    const float t = __ldg(a.tab + (size_t)m * a.K + k);
                    s.x = fmaf(t, v[k].x, s.x); s.y = fmaf(t, v[k].y, s.y); s.z = fmaf(t, v[k].z, s.z); s.w = fmaf(t, v[k].w, s.w);
                }
            *reinterpret_cast<float4*>(a.out + (SYN ? ((size_t)b * a.M + m) * N + (size_t)l * a.C + c : spec(m))) = s;
        }
    }
}

bool thin_xform_supported(const float* in, const float* out, int K, int M, int C) {
    return K >= 1 && M >= 1 && K <= kThinMax && M <= kThinMax && C % 4 == 0 &&
           ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
}

static int thin_launch(const ThinArgs& a, bool syn, const char* what, cudaStream_t st) {
    const long long n = (long long)a.B * a.NY * (a.C / 4);
    if (n <= 0) return FNM_OK;
    {
        LaunchScope ls(what, st);
        if (syn) thin_xform_kernel<true><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a);
        else thin_xform_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a);
    }
    return check_launch(what);
}

int thin_xdft(const float* T, const float* ex, float* Xh, int B, int P1, int R, int NY, int C, cudaStream_t st, int nsum) {
    ThinArgs a{T, ex, Xh, B, 2 * P1, 2 * R, R, NY, C, nsum > 1 ? nsum : 1, (long long)B * 2 * P1 * NY * C};
    return thin_launch(a, false, "xdft", st);
}

int thin_xsyn(const float* Oh, const float* fx, float* Z, int B, int S1, int R, int NY, int C, cudaStream_t st) {
    ThinArgs a{Oh, fx, Z, B, 2 * R, 2 * S1, R, NY, C, 1, 0};
    return thin_launch(a, true, "xsyn", st);
}

static WeightRef make_w(const float* w0, const float* w1, int rpw, int NY, int Co, int Ci = 0, int mm = 0) {
    WeightRef w;
    w.w0 = reinterpret_cast<const float2*>(w0);
    w.w1 = reinterpret_cast<const float2*>(w1 ? w1 : w0);
    w.rpw = rpw; w.NY = NY; w.Co = Co; w.Ci = Ci; w.mm = mm;
    return w;
}

int simt_mix_fwd(const float* Xh, const float* w0, const float* w1, int rpw, float* Oh, int B, int R, int NY, int Ci,
                 int Co, cudaStream_t st, int mode_major) {
    MixFwdP p;
    p.K = 2 * Ci; p.Xh = Xh; p.Oh = Oh; p.w = make_w(w0, w1, rpw, NY, Co, Ci, mode_major); p.M = B; p.N = 2 * Co; p.Ci = Ci; p.Co = Co;
    return launch_auto(p, (int64_t)R * NY, st, "mix_fwd");
}

int simt_mix_bwd_x(const float* Gh, const float* w0, const float* w1, int rpw, float* Gxh, int B, int R, int NY, int Ci,
                   int Co, cudaStream_t st, int mode_major) {
    MixBwdXP p;
    p.K = 2 * Co; p.Gh = Gh; p.Gxh = Gxh; p.w = make_w(w0, w1, rpw, NY, Co, Ci, mode_major); p.M = B; p.N = 2 * Ci; p.Ci = Ci; p.Co = Co;
    return launch_auto(p, (int64_t)R * NY, st, "mix_bwd_x");
}

int simt_mix_bwd_w(const float* Xh, const float* Gh, float* dw0, float* dw1, int rpw, int B, int R, int NY, int Ci,
                   int Co, cudaStream_t st, int mode_major) {
    MixBwdWP p;
    p.mm = mode_major;
    p.K = 2 * B; p.Xh = Xh; p.Gh = Gh;
    p.dw0 = reinterpret_cast<float2*>(dw0); p.dw1 = reinterpret_cast<float2*>(dw1 ? dw1 : dw0);
    p.M = Ci; p.N = 2 * Co; p.B = B; p.Co = Co; p.rpw = rpw; p.NY = NY;
    return launch_auto(p, (int64_t)R * NY, st, "mix_bwd_w");
}

int simt_pointwise_ysyn(const float* x, int act, const float* wc, int wc_is_kn, const float* bias, const float* Z,
                        const float* by, const float* mul, float* out, float* out_act, int64_t rows, int S2, int LY,
                        int K, int N, cudaStream_t st, int flags) {
    PointwiseYsynP p;
    p.flags = flags;
    const int Kc = (x && wc) ? K : 0;
    p.K = Kc + LY; p.x = x; p.wc = wc; p.bias = bias; p.Z = Z; p.by = by; p.mul = mul; p.out = out; p.out_act = out_act;
    p.M = S2; p.N = N; p.Kc = Kc; p.LY = LY; p.act = act; p.wc_is_kn = wc_is_kn;
    return launch_auto(p, rows, st, "pointwise_ysyn");
}

static void conv_split(int64_t npos, int& chunk, int& nsplit) {
    int64_t c = (npos + 591) / 592;              // ~4 waves of 148 SMs
    if (c < 256) c = 256;
    c = (c + 15) / 16 * 16;
    chunk = (int)c;
    nsplit = (int)((npos + c - 1) / c);
}

size_t simt_conv_wgrad_workspace(int64_t npos, int Ci, int Co) {
    int chunk, nsplit;
    conv_split(npos, chunk, nsplit);
    return align_up((size_t)nsplit * Co * (Ci + 1) * sizeof(float), 256);
}

int simt_conv_wgrad(const float* g, const float* x, int act, float* dwc, float* dbc, int64_t npos, int Ci, int Co,
                    float* partial, cudaStream_t st) {
    int chunk, nsplit;
    conv_split(npos, chunk, nsplit);
    ConvWgradP p;
    p.g = g; p.x = x; p.partial = partial; p.M = Co; p.N = Ci + 1; p.Ci = Ci; p.act = act; p.npos = npos; p.chunk = chunk;
    FNM_TRY(launch_auto(p, nsplit, st, "conv_wgrad"));
    const int count = Co * (Ci + 1);
    {
        LaunchScope ls("conv_wgrad_reduce", st);
        conv_wgrad_reduce<<<(count + 255) / 256, 256, 0, st>>>(partial, nsplit, Co, Ci, dwc, dbc);
    }
    return check_launch("conv_wgrad_reduce");
}

static void mode_split(int nmodes, int& per, int& nsplit) {
    per = (nmodes + 147) / 148;
    if (per < 1) per = 1;
    nsplit = (nmodes + per - 1) / per;
}

size_t simt_mode_split_workspace(int nmodes, int B, int C) {
    int per, nsplit;
    mode_split(nmodes, per, nsplit);
    return align_up((size_t)nsplit * B * C * sizeof(float), 256);
}

int simt_lfunc_contract(const float* Xh, const float* w, const float* cc, float* out, float* partial, int B, int R,
                        int NY, int Ci, int Co, cudaStream_t st, int mm) {
    int per, nsplit;
    mode_split(R * NY, per, nsplit);
    LfuncFwdP p;
    p.Xh = Xh; p.cc = cc; p.partial = partial; p.w = make_w(w, nullptr, R, NY, Co, Ci, mm);
    p.M = B; p.N = Co; p.Ci = Ci; p.nmodes = R * NY; p.per = per;
    FNM_TRY(launch_auto(p, nsplit, st, "lfunc_contract"));
    {
        LaunchScope ls("lfunc_reduce", st);
        sum_partials<<<(B * Co + 255) / 256, 256, 0, st>>>(partial, nsplit, B * Co, out);
    }
    return check_launch("lfunc_reduce");
}

int simt_lfunc_bwd_x(const float* g, const float* w, const float* cc, float* Gxh, int B, int R, int NY, int Ci, int Co,
                     cudaStream_t st, int mm) {
    LfuncBwdXP p;
    p.K = Co; p.g = g; p.cc = cc; p.Gxh = Gxh; p.w = make_w(w, nullptr, R, NY, Co, Ci, mm); p.M = B; p.N = 2 * Ci; p.Ci = Ci;
    return launch_auto(p, (int64_t)R * NY, st, "lfunc_bwd_x");
}

int simt_lfunc_bwd_w(const float* g, const float* Xh, const float* cc, float* dw, int B, int R, int NY, int Ci, int Co,
                     cudaStream_t st, int mm) {
    LfuncBwdWP p;
    p.mm = mm;
    p.K = B; p.g = g; p.Xh = Xh; p.cc = cc; p.dw = reinterpret_cast<float2*>(dw);
    p.M = 2 * Ci; p.N = Co; p.Ci = Ci; p.NYR = R * NY;
    return launch_auto(p, (int64_t)R * NY, st, "lfunc_bwd_w");
}

int simt_ldec_mix(const float* v, const float* w, float* Oh, int B, int R, int NY, int Ci, int Co, cudaStream_t st, int mm) {
    LdecFwdP p;
    p.K = Ci; p.v = v; p.Oh = Oh; p.w = make_w(w, nullptr, R, NY, Co, Ci, mm); p.M = B; p.N = 2 * Co; p.Co = Co;
    return launch_auto(p, (int64_t)R * NY, st, "ldec_mix");
}

int simt_ldec_bwd_w(const float* v, const float* Gh, float* dw, int B, int R, int NY, int Ci, int Co, cudaStream_t st, int mm) {
    LdecBwdWP p;
    p.mm = mm;
    p.K = B; p.v = v; p.Gh = Gh; p.dw = reinterpret_cast<float2*>(dw); p.M = Ci; p.N = 2 * Co; p.Co = Co; p.NYR = R * NY;
    return launch_auto(p, (int64_t)R * NY, st, "ldec_bwd_w");
}

int simt_ldec_bwd_v(const float* Gh, const float* w, float* dv, float* partial, int B, int R, int NY, int Ci, int Co,
                    cudaStream_t st, int mm) {
    int per, nsplit;
    mode_split(R * NY, per, nsplit);
    LdecBwdVP p;
    p.Gh = Gh; p.partial = partial; p.w = make_w(w, nullptr, R, NY, Co, Ci, mm);
    p.M = B; p.N = Ci; p.Co = Co; p.nmodes = R * NY; p.per = per;
    FNM_TRY(launch_auto(p, nsplit, st, "ldec_bwd_v"));
    {
        LaunchScope ls("ldec_reduce", st);
        sum_partials<<<(B * Ci + 255) / 256, 256, 0, st>>>(partial, nsplit, B * Ci, dv);
    }
    return check_launch("ldec_reduce");
}

// ---- F2V / V2F ends on the mixing kernels: a real (B, C) matrix becomes a planar spectrum that is the same in every mode
// (times cc[mode]), and the real part of a spectrum is summed over modes (times cc[mode]).
//   spread: dst[mode][0][j] = cc[mode] * src[j], dst[mode][1][j] = 0          (j < n = B * C)
__global__ void mode_spread_kernel(const float* __restrict__ src, const float* __restrict__ cc, float* __restrict__ dst,
                                   int nmodes, int n) {
    const int j4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (j4 >= n) return;
    const float4 v = *reinterpret_cast<const float4*>(src + j4);
    for (int m = blockIdx.y; m < nmodes; m += gridDim.y) {
        const float c = cc ? __ldg(cc + m) : 1.f;
        float* d = dst + (size_t)m * 2 * n + j4;
        *reinterpret_cast<float4*>(d) = make_float4(c * v.x, c * v.y, c * v.z, c * v.w);
        *reinterpret_cast<float4*>(d + n) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

//   reduce: out[j] = sum_mode cc[mode] * src[mode][0][j]; 32 outputs x 8 mode groups per CTA, fixed summation order
__global__ void __launch_bounds__(256) mode_reduce_kernel(const float* __restrict__ src, const float* __restrict__ cc,
                                                          float* __restrict__ out, int nmodes, int n) {
    __shared__ float part[8][33];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    float acc = 0.f;
    if (j < n)
        for (int m = grp; m < nmodes; m += 8) acc = fmaf(cc ? __ldg(cc + m) : 1.f, __ldg(src + (size_t)m * 2 * n + j), acc);
    part[grp][lane] = acc;
    __syncthreads();
    if (grp == 0 && j < n) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) s += part[g][lane];
        out[j] = s;
    }
}

int mode_spread(const float* src, const float* cc, float* dst, int nmodes, int n, const char* what, cudaStream_t st) {
    if (n % 4 != 0) return fail(FNM_ENOTSUP, "%s: extent %d is no multiple of 4", what, n);
    if (nmodes <= 0 || n <= 0) return FNM_OK;
    const int bx = (n / 4 + 127) / 128;
    int by = (4 * 148 + bx - 1) / bx;
    if (by > nmodes) by = nmodes;
    {
        LaunchScope ls(what, st);
        mode_spread_kernel<<<dim3(bx, by), 128, 0, st>>>(src, cc, dst, nmodes, n);
    }
    return check_launch(what);
}

int mode_reduce(const float* src, const float* cc, float* out, int nmodes, int n, const char* what, cudaStream_t st) {
    if (n <= 0) return FNM_OK;
    {
        LaunchScope ls(what, st);
        mode_reduce_kernel<<<(n + 31) / 32, 256, 0, st>>>(src, cc, out, nmodes, n);
    }
    return check_launch(what);
}

//   batch reduce: out[b][n] = sum_{j < J} in[b][j][n]; 32 outputs x 8 groups of j per CTA, fixed summation order
__global__ void __launch_bounds__(256) batch_reduce_kernel(const float* __restrict__ in, float* __restrict__ out, int J, int N) {
    __shared__ float part[8][33];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + lane;
    const float* src = in + (size_t)blockIdx.y * J * N;
    float acc = 0.f;
    if (n < N)
        for (int j = grp; j < J; j += 8) acc += __ldg(src + (size_t)j * N + n);
    part[grp][lane] = acc;
    __syncthreads();
    if (grp == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) s += part[g][lane];
        out[(size_t)blockIdx.y * N + n] = s;
    }
}

int batch_reduce(const float* in, float* out, int B, int J, int N, const char* what, cudaStream_t st) {
    if (B <= 0 || N <= 0) return FNM_OK;
    if (B > 65535) return fail(FNM_ENOTSUP, "%s: batch extent %d exceeds the grid limit", what, B);
    {
        LaunchScope ls(what, st);
        batch_reduce_kernel<<<dim3((N + 31) / 32, B), 256, 0, st>>>(in, out, J, N);
    }
    return check_launch(what);
}

}  // namespace fnm
