// Fused multi-tensor complex-aware Adam (reference: util/Adam.py:25-51) and the stand-alone
// GELU / GELU' elementwise kernels.  Pure HBM streaming: 16-byte vector accesses, one launch for
// up to FNM_ADAM_MAX_TENSORS parameter tensors (the reference issues ~9 launches per tensor).
#include "fnm_common.cuh"

namespace fnm {

constexpr int ADAM_CHUNK = 4096;          // floats per block
constexpr int ADAM_THREADS = 256;

struct AdamArgs {
    float* p[FNM_ADAM_MAX_TENSORS];
    const float* g[FNM_ADAM_MAX_TENSORS];
    float* m[FNM_ADAM_MAX_TENSORS];
    float* v[FNM_ADAM_MAX_TENSORS];
    float* vmax[FNM_ADAM_MAX_TENSORS];
    long long numel[FNM_ADAM_MAX_TENSORS];
    int block_start[FNM_ADAM_MAX_TENSORS + 1];
    unsigned char cplx[FNM_ADAM_MAX_TENSORS];
    int n;
    double lr, beta1, beta2, eps, wd;
    int amsgrad, step;
    const int* step_dev;
};

struct AdamScalars {
    float b1, one_m_b1, b2, one_m_b2, wd, eps, neg_step_size, sqrt_bc2;
};

__device__ __forceinline__ void adam_real(float& p, float g, float& m, float& v, float* vmax, const AdamScalars& s) {
    if (s.wd != 0.f) g = fmaf(s.wd, p, g);
    m = m * s.b1 + s.one_m_b1 * g;
    v = v * s.b2 + s.one_m_b2 * (g * g);
    float second = v;
    if (vmax) { *vmax = fmaxf(*vmax, v); second = *vmax; }
    const float denom = sqrtf(second) / s.sqrt_bc2 + s.eps;
    p = p + s.neg_step_size * m / denom;
}

// one complex element: both moments of m, ONE shared |g|^2 second moment stored in v.re (v.im = 0 decays)
__device__ __forceinline__ void adam_cplx(float& pr, float& pi, float gr, float gi, float& mr, float& mi, float& vr,
                                          float& vi, const AdamScalars& s) {
    if (s.wd != 0.f) { gr = fmaf(s.wd, pr, gr); gi = fmaf(s.wd, pi, gi); }
    mr = mr * s.b1 + s.one_m_b1 * gr;
    mi = mi * s.b1 + s.one_m_b1 * gi;
    vr = vr * s.b2 + s.one_m_b2 * (gr * gr + gi * gi);
    vi = vi * s.b2;
    const float denom = sqrtf(vr) / s.sqrt_bc2 + s.eps;
    pr = pr + s.neg_step_size * mr / denom;
    pi = pi + s.neg_step_size * mi / denom;
}

__global__ void __launch_bounds__(ADAM_THREADS) adam_kernel(const AdamArgs a) {
    __shared__ AdamScalars sc;
    __shared__ int s_tensor;
    if (threadIdx.x == 0) {
        const int step = a.step_dev ? *a.step_dev : a.step;
        const double bc1 = 1.0 - pow(a.beta1, (double)step);
        const double bc2 = 1.0 - pow(a.beta2, (double)step);
        sc.b1 = (float)a.beta1; sc.one_m_b1 = (float)(1.0 - a.beta1);
        sc.b2 = (float)a.beta2; sc.one_m_b2 = (float)(1.0 - a.beta2);
        sc.wd = (float)a.wd; sc.eps = (float)a.eps;
        sc.neg_step_size = (float)(-(a.lr / bc1));
        sc.sqrt_bc2 = (float)sqrt(bc2);
        int t = 0;
        while (t + 1 < a.n && (int)blockIdx.x >= a.block_start[t + 1]) ++t;
        s_tensor = t;
    }
    __syncthreads();
    const AdamScalars s = sc;
    const int t = s_tensor;
    const long long base = (long long)(blockIdx.x - a.block_start[t]) * ADAM_CHUNK;
    const long long n = a.numel[t];
    float* __restrict__ p = a.p[t];
    const float* __restrict__ g = a.g[t];
    float* __restrict__ m = a.m[t];
    float* __restrict__ v = a.v[t];
    float* __restrict__ vmax = a.amsgrad ? a.vmax[t] : nullptr;
    const bool cplx = a.cplx[t] != 0;
    const bool vec_ok = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)vmax) & 15) == 0);

#pragma unroll
    for (int it = 0; it < ADAM_CHUNK / (ADAM_THREADS * 4); ++it) {
        const long long i = base + ((long long)it * ADAM_THREADS + threadIdx.x) * 4;
        if (i >= n) break;
        if (vec_ok && i + 4 <= n) {
            float4 p4 = *reinterpret_cast<float4*>(p + i);
            const float4 g4 = *reinterpret_cast<const float4*>(g + i);
            float4 m4 = *reinterpret_cast<float4*>(m + i);
            float4 v4 = *reinterpret_cast<float4*>(v + i);
            if (cplx) {
                adam_cplx(p4.x, p4.y, g4.x, g4.y, m4.x, m4.y, v4.x, v4.y, s);
                adam_cplx(p4.z, p4.w, g4.z, g4.w, m4.z, m4.w, v4.z, v4.w, s);
            } else if (vmax) {
                float4 x4 = *reinterpret_cast<float4*>(vmax + i);
                adam_real(p4.x, g4.x, m4.x, v4.x, &x4.x, s);
                adam_real(p4.y, g4.y, m4.y, v4.y, &x4.y, s);
                adam_real(p4.z, g4.z, m4.z, v4.z, &x4.z, s);
                adam_real(p4.w, g4.w, m4.w, v4.w, &x4.w, s);
                *reinterpret_cast<float4*>(vmax + i) = x4;
            } else {
                adam_real(p4.x, g4.x, m4.x, v4.x, nullptr, s);
                adam_real(p4.y, g4.y, m4.y, v4.y, nullptr, s);
                adam_real(p4.z, g4.z, m4.z, v4.z, nullptr, s);
                adam_real(p4.w, g4.w, m4.w, v4.w, nullptr, s);
            }
            *reinterpret_cast<float4*>(p + i) = p4;
            *reinterpret_cast<float4*>(m + i) = m4;
            *reinterpret_cast<float4*>(v + i) = v4;
        } else {
            // unaligned tensor or tail: complex elements are (re,im) pairs, numel is even for them
            const long long end = (i + 4 < n) ? i + 4 : n;
            if (cplx) {
                for (long long j = i; j + 1 < end; j += 2)
                    adam_cplx(p[j], p[j + 1], g[j], g[j + 1], m[j], m[j + 1], v[j], v[j + 1], s);
            } else {
                for (long long j = i; j < end; ++j) adam_real(p[j], g[j], m[j], v[j], vmax ? vmax + j : nullptr, s);
            }
        }
    }
}

int adam_launch(int n, float* const* p, const float* const* g, float* const* m, float* const* v, float* const* vmax,
                const int64_t* numel, const uint8_t* is_complex, double lr, double beta1, double beta2, double eps,
                double wd, int amsgrad, int step, const int32_t* step_dev, cudaStream_t st) {
    for (int t0 = 0; t0 < n; t0 += FNM_ADAM_MAX_TENSORS) {
        const int cnt = (n - t0 < FNM_ADAM_MAX_TENSORS) ? n - t0 : FNM_ADAM_MAX_TENSORS;
        AdamArgs a;
        int blocks = 0, used = 0;
        for (int t = 0; t < cnt; ++t) {
            const int64_t ne = numel[t0 + t];
            if (ne <= 0) continue;
            if (is_complex[t0 + t] && (ne & 1)) return fail(FNM_EINVAL, "adam: complex tensor %d has odd float count", t0 + t);
            if (is_complex[t0 + t] && amsgrad)
                return fail(FNM_EINVAL, "adam: amsgrad is not defined for complex parameters (util/Adam.py:43 raises too)");
            a.p[used] = p[t0 + t]; a.g[used] = g[t0 + t]; a.m[used] = m[t0 + t]; a.v[used] = v[t0 + t];
            a.vmax[used] = (amsgrad && vmax) ? vmax[t0 + t] : nullptr;
            if (amsgrad && !a.vmax[used]) return fail(FNM_EINVAL, "adam: amsgrad needs vmax for tensor %d", t0 + t);
            a.numel[used] = ne; a.cplx[used] = is_complex[t0 + t];
            a.block_start[used] = blocks;
            blocks += (int)((ne + ADAM_CHUNK - 1) / ADAM_CHUNK);
            ++used;
        }
        if (used == 0) continue;
        a.block_start[used] = blocks;
        a.n = used; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = wd;
        a.amsgrad = amsgrad; a.step = step; a.step_dev = step_dev;
        {
            LaunchScope ls("adam", st);
            adam_kernel<<<blocks, ADAM_THREADS, 0, st>>>(a);
        }
        FNM_TRY(check_launch("adam"));
    }
    return FNM_OK;
}

__global__ void counter_inc_kernel(int* c) { *c += 1; }

int counter_inc(int32_t* c, cudaStream_t st) {
    {
        LaunchScope ls("counter_inc", st);
        counter_inc_kernel<<<1, 1, 0, st>>>(c);
    }
    return check_launch("counter_inc");
}

// ---- stand-alone activation --------------------------------------------------------------------
template <bool GRAD>
__global__ void gelu_kernel(const float* __restrict__ a, const float* __restrict__ z, float* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n && ((((uintptr_t)a | (uintptr_t)z | (uintptr_t)out) & 15) == 0)) {
            const float4 z4 = *reinterpret_cast<const float4*>(z + i);
            float4 o;
            if (GRAD) {
                const float4 g4 = *reinterpret_cast<const float4*>(a + i);
                o.x = g4.x * gelu_grad_f(z4.x); o.y = g4.y * gelu_grad_f(z4.y);
                o.z = g4.z * gelu_grad_f(z4.z); o.w = g4.w * gelu_grad_f(z4.w);
            } else {
                o.x = gelu_f(z4.x); o.y = gelu_f(z4.y); o.z = gelu_f(z4.z); o.w = gelu_f(z4.w);
            }
            *reinterpret_cast<float4*>(out + i) = o;
        } else {
            for (long long j = i; j < n && j < i + 4; ++j) out[j] = GRAD ? a[j] * gelu_grad_f(z[j]) : gelu_f(z[j]);
        }
    }
}

int gelu_launch(const float* g, const float* z, float* out, int64_t n, bool grad, cudaStream_t st) {
    if (n <= 0) return FNM_OK;
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    {
        LaunchScope ls(grad ? "gelu_bwd" : "gelu_fwd", st);
        if (grad) gelu_kernel<true><<<(int)blocks, 256, 0, st>>>(g, z, out, n);
        else gelu_kernel<false><<<(int)blocks, 256, 0, st>>>(nullptr, z, out, n);
    }
    return check_launch("gelu");
}

}  // namespace fnm
