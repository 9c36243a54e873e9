// Shared device/host helpers for libfnm_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/fnm_b200.h"

namespace fnm {

// ---- error reporting (thread-local: forward runs on the main thread, backward on autograd's worker)
char* err_buf();
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);

// Brackets one kernel launch: counts it and, while profiling is on, records a CUDA-event pair on the
// launching stream.  Usage:  { LaunchScope ls("name", st); kernel<<<...>>>(...); }  then check_launch.
struct LaunchScope {
    const char* name;
    cudaStream_t st;
    int slot;
    LaunchScope(const char* name_, cudaStream_t st_);
    ~LaunchScope();
};

#define FNM_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) return ::fnm::fail(FNM_EINVAL, __VA_ARGS__); \
    } while (0)

#define FNM_TRY(expr)                   \
    do {                                \
        int _rc = (expr);               \
        if (_rc != FNM_OK) return _rc;  \
    } while (0)

static inline cudaStream_t as_stream(fnm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- exact (erf) GELU, the F.gelu default the reference uses (models/shared.py:13-14)
__device__ __forceinline__ float gelu_f(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// Once-per-DEVICE guard for cudaFuncSetAttribute: function attributes belong to the device's context, so a process that
// touches a second GPU has to set them there too (a plain `static bool` left every kernel with more than 48 KB of dynamic
// shared memory unlaunchable on the second device).  One bit per device ordinal; a race sets the attribute twice.
static inline bool once_needed(std::atomic<uint64_t>& mask, int& dev) {
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    return ((mask.load(std::memory_order_acquire) >> (dev & 63)) & 1ull) == 0;
}
static inline void once_done(std::atomic<uint64_t>& mask, int dev) { mask.fetch_or(1ull << (dev & 63), std::memory_order_release); }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// carve consecutive 256-byte aligned float buffers out of a caller-owned workspace
struct Carver {
    char* base;
    size_t off, cap;
    Carver(void* p, size_t bytes) : base(static_cast<char*>(p)), off(0), cap(bytes) {}
    float* take(size_t nfloats) {
        size_t o = align_up(off, 256);
        off = o + nfloats * sizeof(float);
        return reinterpret_cast<float*>(base + o);
    }
    bool ok() const { return off <= cap; }
};

}  // namespace fnm
