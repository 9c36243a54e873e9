// extern "C" surface of libfnm_b200 (see include/fnm_b200.h): argument checks, workspace carving,
// and the stage chains that make up a fused Fourier layer / F2V / V2F, forward and backward.
#include "fnm_common.cuh"
#include "fnm_stages.cuh"

#include <string.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace fnm {

static thread_local char g_err[512] = "";

char* err_buf() { return g_err; }

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    const cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(FNM_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    }
    return FNM_OK;
}

// ---- instrumentation ---------------------------------------------------------------------------
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_profiling{0};
static std::mutex g_prof_mu;
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
static std::vector<ProfRec> g_prof;

LaunchScope::LaunchScope(const char* name_, cudaStream_t st_) : name(name_), st(st_), slot(-1) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (g_profiling.load(std::memory_order_relaxed)) {
        ProfRec r;
        r.name = name;
        if (cudaEventCreate(&r.e0) == cudaSuccess && cudaEventCreate(&r.e1) == cudaSuccess) {
            cudaEventRecord(r.e0, st);
            std::lock_guard<std::mutex> lk(g_prof_mu);
            g_prof.push_back(r);
            slot = (int)g_prof.size() - 1;
        }
    }
}

LaunchScope::~LaunchScope() {
    if (slot >= 0) {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        cudaEventRecord(g_prof[slot].e1, st);
    }
}

static int check_plan(const fnm_plan* p, bool need_in, bool need_out) {
    FNM_REQUIRE(p != nullptr, "plan is NULL");
    FNM_REQUIRE(p->B > 0 && p->R > 0 && p->NY > 0 && p->Ci > 0 && p->Co > 0, "plan: non-positive extent");
    FNM_REQUIRE(p->rows_per_w > 0 && p->rows_per_w <= p->R, "plan: rows_per_w out of range");
    if (need_in) {
        FNM_REQUIRE(p->P1 > 0 && p->P2 > 0, "plan: bad input grid");
        FNM_REQUIRE(p->ay && p->ayT && p->ex && p->fxb, "plan: analysis tables missing");
    }
    if (need_out) {
        FNM_REQUIRE(p->S1 > 0 && p->S2 > 0, "plan: bad output grid");
        FNM_REQUIRE(p->by && p->byT && p->fx && p->exb, "plan: synthesis tables missing");
    }
    return FNM_OK;
}

static size_t conv_wgrad_workspace(int64_t npos, int Ci, int Co) {
    const size_t a = simt_conv_wgrad_workspace(npos, Ci, Co);
    const size_t b = tc_conv_wgrad_supported(npos, Ci, Co) ? tc_conv_wgrad_workspace(npos, Ci, Co) : 0;
    return a > b ? a : b;
}

static int conv_wgrad(const float* g, const float* x, int act, float* dwc, float* dbc, int64_t npos, int Ci, int Co,
                      float* partial, int mode, cudaStream_t st) {
    if (tc_conv_wgrad_supported(npos, Ci, Co)) return tc_conv_wgrad(g, x, act, dwc, dbc, npos, Ci, Co, partial, mode, st);
    return simt_conv_wgrad(g, x, act, dwc, dbc, npos, Ci, Co, partial, st);
}

struct LayerWs {
    size_t T, Oh, Z, Gh, partial, shadow;      // float counts (shadow: one mode-major weight copy, 0 = SIMT mixing)
};

static LayerWs layer_ws(const fnm_plan* p) {
    LayerWs w;
    const size_t LY = 2 * (size_t)p->NY;
    const size_t cmax = p->Ci > p->Co ? p->Ci : p->Co;
    const size_t rows = (size_t)p->B * (p->P1 > p->S1 ? p->P1 : p->S1);
    w.T = rows * LY * cmax;                               // T / Tg, then reused as Z / Zg
    if (p->P1 == 1 && p->S1 == 1) {                       // 1-D: room for the partial transforms of the segmented y-DFT
        const int si = tc_ydft_tma_segments(p->B, p->P2, p->Ci), so = tc_ydft_tma_segments(p->B, p->S2, p->Co);
        w.T *= (size_t)(si > so ? si : so);
    }
    w.Oh = (size_t)p->R * p->NY * 2 * p->B * cmax;        // Oh / Gxh
    w.Z = rows * LY * cmax;
    w.Gh = (size_t)p->R * p->NY * 2 * p->B * cmax;
    w.partial = conv_wgrad_workspace((int64_t)p->B * p->S1 * p->S2, p->Ci, p->Co) / sizeof(float);
    w.shadow = tc_mix_supported(p->B, p->Ci, p->Co) ? tc_mix_shadow_floats(p->R * p->NY, p->Ci, p->Co) : 0;
    return w;
}

}  // namespace fnm

using namespace fnm;

extern "C" {

const char* fnm_version(void) { return "fnm_b200 0.2 (sm_100a)"; }
size_t fnm_plan_sizeof(void) { return sizeof(fnm_plan); }
const char* fnm_last_error(void) { return err_buf(); }
int fnm_has_tcgen05(void) { return tc_available(); }

int64_t fnm_launch_count(void) { return g_launches.load(); }

int fnm_set_sm_limit(int n) { return tc_set_sm_limit(n); }

int fnm_profile_begin(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof.clear();
    g_profiling.store(1);
    return FNM_OK;
}

size_t fnm_profile_end(char* buf, size_t n) {
    g_profiling.store(0);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<long long, double>> agg;
    for (auto& r : g_prof) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            auto& a = agg[r.name];
            a.first += 1;
            a.second += ms;
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_prof.clear();
    std::string out;
    char line[160];
    for (auto& kv : agg) {
        snprintf(line, sizeof(line), "%s:%lld:%.6f;", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (buf && n > 0) {
        const size_t k = out.size() < n - 1 ? out.size() : n - 1;
        memcpy(buf, out.data(), k);
        buf[k] = 0;
    }
    return out.size() + 1;
}

// ------------------------------------------------------------------------------------------------
// individual stages
// ------------------------------------------------------------------------------------------------
int fnm_ydft(const float* x, const float* tab, float* T, int64_t rows, int P2, int LY, int C, int act_in, int mode,
             fnm_stream_t stream) {
    FNM_REQUIRE(x && tab && T, "ydft: NULL pointer");
    FNM_REQUIRE(rows >= 0 && P2 > 0 && LY > 0 && C > 0, "ydft: bad extents");
    if (tc_ydft_tma_supported(x, tab, P2, rows, P2, LY, C, act_in)) return tc_ydft_tma(x, tab, P2, T, rows, P2, LY, C, mode, as_stream(stream));
    if (tc_ydft_supported(rows, P2, LY, C)) return tc_ydft(x, tab, T, rows, P2, LY, C, act_in, mode, as_stream(stream));
    return simt_ydft(x, tab, T, rows, P2, LY, C, act_in, as_stream(stream));
}

static int plan_ydft(const float* x, const float* tab, const float* tab4, float* T, int64_t rows, int P2, int LY, int C, int act_in,
                     int mode, fnm_stream_t stream);

// y-DFT followed by the x-DFT of the fused layer.  1-D models (P1 = 1) have few, long rows -- config 2: 64 rows of 1152 points,
// 32 work items for 148 SMs -- so the TMA-fed y-DFT splits each row into K segments, every (segment, row pair) an item, and the
// thin x-transform (which only re-lays the two values of a 1-D "x spectrum" out) adds the partial transforms up.  T must hold
// tc_ydft_tma_segments() partials (layer_ws).
static int plan_ydft_xdft(const float* x, const float* tab, const float* tab4, float* T, const float* ex, float* Xh, int B, int P1,
                          int P2, int R, int NY, int C, int act_in, int mode, fnm_stream_t stream) {
    const int LY = 2 * NY;
    if (P1 == 1 && act_in == 0 && thin_xform_supported(T, Xh, 2, 2 * R, C)) {
        const int nseg = tc_ydft_tma_segments(B, P2, C);
        const bool pad = P2 % 4 != 0;
        const float* t = pad ? tab4 : tab;
        const int ld = pad ? ((P2 + 3) & ~3) : P2;
        if (nseg > 1 && t && tc_ydft_tma_supported(x, t, ld, B, P2, LY, C, 0, nseg)) {
            FNM_TRY(tc_ydft_tma(x, t, ld, T, B, P2, LY, C, mode, as_stream(stream), nseg));
            return thin_xdft(T, ex, Xh, B, 1, R, NY, C, as_stream(stream), nseg);
        }
    }
    FNM_TRY(plan_ydft(x, tab, tab4, T, (int64_t)B * P1, P2, LY, C, act_in, mode, stream));
    return fnm_xdft(T, ex, Xh, B, P1, R, NY, C, mode, stream);
}

// y-DFT of a plan's entry points: `tab4` is the plan's copy of `tab` with rows padded to a multiple of 4 floats (fnm_plan.ay4 /
// byT4, may be NULL) -- with it the TMA-fed kernel also serves row lengths that are no multiple of 4
static int plan_ydft(const float* x, const float* tab, const float* tab4, float* T, int64_t rows, int P2, int LY, int C, int act_in,
                     int mode, fnm_stream_t stream) {
    const int ld4 = (P2 + 3) & ~3;
    if (P2 % 4 != 0 && tab4 && tc_ydft_tma_supported(x, tab4, ld4, rows, P2, LY, C, act_in))
        return tc_ydft_tma(x, tab4, ld4, T, rows, P2, LY, C, mode, as_stream(stream));
    return fnm_ydft(x, tab, T, rows, P2, LY, C, act_in, mode, stream);
}

int fnm_xdft(const float* T, const float* ex, float* Xh, int B, int P1, int R, int NY, int C, int mode,
             fnm_stream_t stream) {
    FNM_REQUIRE(T && ex && Xh, "xdft: NULL pointer");
    FNM_REQUIRE(B > 0 && P1 > 0 && R > 0 && NY > 0 && C > 0, "xdft: bad extents");
    if (thin_xform_supported(T, Xh, 2 * P1, 2 * R, C)) return thin_xdft(T, ex, Xh, B, P1, R, NY, C, as_stream(stream));
    if (tc_xdft_supported(B, P1, R, NY, C)) return tc_xdft(T, ex, Xh, B, P1, R, NY, C, mode, as_stream(stream));
    return simt_xdft(T, ex, Xh, B, P1, R, NY, C, as_stream(stream));
}

int fnm_xsyn(const float* Oh, const float* fx, float* Z, int B, int S1, int R, int NY, int C, int mode,
             fnm_stream_t stream) {
    FNM_REQUIRE(Oh && fx && Z, "xsyn: NULL pointer");
    FNM_REQUIRE(B > 0 && S1 > 0 && R > 0 && NY > 0 && C > 0, "xsyn: bad extents");
    if (thin_xform_supported(Oh, Z, 2 * R, 2 * S1, C)) return thin_xsyn(Oh, fx, Z, B, S1, R, NY, C, as_stream(stream));
    if (tc_xsyn_supported(B, S1, R, NY, C)) return tc_xsyn(Oh, fx, Z, B, S1, R, NY, C, mode, as_stream(stream));
    return simt_xsyn(Oh, fx, Z, B, S1, R, NY, C, as_stream(stream));
}

int fnm_resample(const float* in, const float* tab, float* out, int64_t G, int K, int M, int inner, int C, int mode,
                 fnm_stream_t stream) {
    FNM_REQUIRE(in && tab && out, "resample: NULL pointer");
    FNM_REQUIRE(G >= 0 && K > 0 && M > 0 && inner > 0 && C > 0, "resample: bad extents");
    if (G == 0) return FNM_OK;
    if (tc_resample_supported(G, K, M, inner, C))
        return tc_resample(in, tab, out, G, K, M, inner, C, mode, as_stream(stream));
    // same contraction with the trailing (inner, C) extent taken as one channel axis
    FNM_REQUIRE((int64_t)inner * C < (1LL << 30), "resample: trailing extent too large");
    return simt_ydft(in, tab, out, G, K, M, inner * C, 0, as_stream(stream));
}

size_t fnm_mix_workspace_bytes(int B, int R, int NY, int Ci, int Co) {
    return tc_mix_supported(B, Ci, Co) ? tc_mix_shadow_floats(R * NY, Ci, Co) * sizeof(float) + 256 : 0;
}

// the tcgen05 path needs scratch for the mode-major weight copy; without it (or for shapes it does not
// serve) the fp32 SIMT kernels run on the reference layout directly
static bool mix_use_tc(int B, int R, int NY, int Ci, int Co, void* ws, size_t ws_bytes) {
    return ws && tc_mix_supported(B, Ci, Co) && ws_bytes >= fnm_mix_workspace_bytes(B, R, NY, Ci, Co);
}

int fnm_mix_fwd(const float* Xh, const float* w0, const float* w1, int rows_per_w, float* Oh, int B, int R, int NY,
                int Ci, int Co, void* workspace, size_t workspace_bytes, int mode, fnm_stream_t stream) {
    FNM_REQUIRE(Xh && w0 && Oh, "mix_fwd: NULL pointer");
    FNM_REQUIRE(rows_per_w == R || (w1 && 2 * rows_per_w == R), "mix_fwd: weight rows do not cover R");
    if (mix_use_tc(B, R, NY, Ci, Co, workspace, workspace_bytes)) {
        float* W1 = static_cast<float*>(workspace);
        FNM_TRY(tc_mix_pack(w0, w1, rows_per_w, R, NY, Ci, Co, 0, W1, as_stream(stream)));
        return tc_mix_apply(W1, nullptr, 0, Xh, Oh, R * NY, B, Ci, Co, 0, 0, mode, "mix_fwd", as_stream(stream));
    }
    return simt_mix_fwd(Xh, w0, w1, rows_per_w, Oh, B, R, NY, Ci, Co, as_stream(stream));
}

int fnm_mix_bwd_x(const float* Gh, const float* w0, const float* w1, int rows_per_w, float* Gxh, int B, int R, int NY,
                  int Ci, int Co, void* workspace, size_t workspace_bytes, int mode, fnm_stream_t stream) {
    FNM_REQUIRE(Gh && w0 && Gxh, "mix_bwd_x: NULL pointer");
    FNM_REQUIRE(rows_per_w == R || (w1 && 2 * rows_per_w == R), "mix_bwd_x: weight rows do not cover R");
    if (mix_use_tc(B, R, NY, Ci, Co, workspace, workspace_bytes)) {
        // the same mode-major copy W1[m][i][o] as the forward when its rows can be read with 16-byte loads
        const int w1_layout = Co % 2 == 0 ? 1 : 0;
        float* Wm = static_cast<float*>(workspace);
        FNM_TRY(tc_mix_pack(w0, w1, rows_per_w, R, NY, Ci, Co, w1_layout ? 0 : 1, Wm, as_stream(stream)));
        return tc_mix_apply(Wm, nullptr, 0, Gh, Gxh, R * NY, B, Co, Ci, 1, w1_layout, mode, "mix_bwd_x", as_stream(stream));
    }
    return simt_mix_bwd_x(Gh, w0, w1, rows_per_w, Gxh, B, R, NY, Ci, Co, as_stream(stream));
}

int fnm_mix_bwd_w(const float* Xh, const float* Gh, float* dw0, float* dw1, int rows_per_w, int B, int R, int NY,
                  int Ci, int Co, void* workspace, size_t workspace_bytes, int mode, fnm_stream_t stream) {
    FNM_REQUIRE(Xh && Gh && dw0, "mix_bwd_w: NULL pointer");
    FNM_REQUIRE(rows_per_w == R || (dw1 && 2 * rows_per_w == R), "mix_bwd_w: weight rows do not cover R");
    if (mix_use_tc(B, R, NY, Ci, Co, workspace, workspace_bytes)) {
        float* dW1 = static_cast<float*>(workspace);
        FNM_TRY(tc_mix_wgrad(Xh, Gh, dW1, nullptr, 0, R * NY, B, Ci, Co, mode, as_stream(stream)));
        return tc_mix_unpack(dW1, dw0, dw1, rows_per_w, R, NY, Ci, Co, as_stream(stream));
    }
    return simt_mix_bwd_w(Xh, Gh, dw0, dw1, rows_per_w, B, R, NY, Ci, Co, as_stream(stream));
}

int fnm_pointwise_ysyn_ex(const float* x, int act_in, const float* wc, int wc_is_kn, const float* bias, const float* Z,
                          const float* by, const float* mul, float* out, float* out_act, int64_t rows, int S2,
                          int LY, int K, int N, int mode, int flags, const void* by16, fnm_stream_t stream) {
    FNM_REQUIRE(out && (LY == 0 || (Z && by)), "pointwise_ysyn: NULL pointer");
    FNM_REQUIRE(rows >= 0 && S2 > 0 && LY >= 0 && N > 0 && K >= 0, "pointwise_ysyn: bad extents");
    FNM_REQUIRE(LY > 0 || (x && wc && K > 0), "pointwise_ysyn: nothing to compute (no spectral and no pointwise term)");
    FNM_REQUIRE((x == nullptr) == (wc == nullptr) || K == 0, "pointwise_ysyn: x and wc must be given together");
    FNM_REQUIRE(!(flags & FNM_PW_OUT_IS_GRAD) || (out_act && !mul), "pointwise_ysyn: FNM_PW_OUT_IS_GRAD needs out_act and no multiplier");
    FNM_REQUIRE(!(flags & FNM_PW_MUL_IS_GRAD) || mul, "pointwise_ysyn: FNM_PW_MUL_IS_GRAD without a multiplier");
    if (tc_pw2_supported(rows, S2, LY, (x && wc) ? K : 0, N, (x && wc) ? act_in : 0, (x && wc) ? x : nullptr, Z, by))
        return tc_pw2(x, wc, wc_is_kn, bias, Z, by, mul, out, out_act, rows, S2, LY, K, N, mode, as_stream(stream), 0, flags, by16);
    if (flags == 0 && LY > 0 && tc_pointwise_ysyn_supported(rows, S2, LY, (x && wc) ? K : 0, N))
        return tc_pointwise_ysyn(x, act_in, wc, wc_is_kn, bias, Z, by, mul, out, out_act, rows, S2, LY, K, N,
                                 mode, as_stream(stream));
    return simt_pointwise_ysyn(x, act_in, wc, wc_is_kn, bias, Z, by, mul, out, out_act, rows, S2, LY, K, N,
                               as_stream(stream), flags);
}

int fnm_pointwise_ysyn(const float* x, int act_in, const float* wc, int wc_is_kn, const float* bias, const float* Z,
                       const float* by, const float* mul_gelu_grad_of, float* out, float* out_act, int64_t rows, int S2,
                       int LY, int K, int N, int mode, fnm_stream_t stream) {
    return fnm_pointwise_ysyn_ex(x, act_in, wc, wc_is_kn, bias, Z, by, mul_gelu_grad_of, out, out_act, rows, S2, LY, K, N, mode,
                                 0, nullptr, stream);
}

// The F2V / V2F ends keep m2 + 1 columns, so their y-synthesis has LY = 2 (m2 + 1) table columns -- usually no multiple
// of 4, which the TMA kernel's table map (16-byte row stride) cannot address.  With scratch at hand the table is copied
// once per call into a zero-padded [S2][round4(LY)] buffer (a few KB) so that these launches run on the tcgen05 kernel
// too instead of the fp32 SIMT engine.
__global__ void pad_table_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int ld, int ld_pad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ld_pad) return;
    const int r = i / ld_pad, c = i - r * ld_pad;
    dst[i] = c < ld ? src[(size_t)r * ld + c] : 0.f;
}

static size_t padded_table_bytes(int S2, int LY) { return LY % 4 == 0 ? 0 : align_up((size_t)S2 * ((LY + 3) / 4 * 4) * sizeof(float), 256); }

static int ysyn_only(const float* Z, const float* by, const float* mul, float* out, int64_t rows, int S2, int LY, int N, int mode,
                     float* tab_scratch, fnm_stream_t stream, int flags = 0) {
    const int ld = (LY + 3) / 4 * 4;
    if (LY % 4 != 0 && tab_scratch && tc_pw2_supported(rows, S2, LY, 0, N, 0, nullptr, Z, tab_scratch, ld)) {
        cudaStream_t st = as_stream(stream);
        {
            LaunchScope ls("pad_table", st);
            pad_table_kernel<<<(S2 * ld + 255) / 256, 256, 0, st>>>(by, tab_scratch, S2, LY, ld);
        }
        FNM_TRY(check_launch("pad_table"));
        return tc_pw2(nullptr, nullptr, 0, nullptr, Z, tab_scratch, mul, out, nullptr, rows, S2, LY, 0, N, mode, st, ld, flags);
    }
    return fnm_pointwise_ysyn_ex(nullptr, 0, nullptr, 0, nullptr, Z, by, mul, out, nullptr, rows, S2, LY, 0, N, mode, flags, nullptr, stream);
}

size_t fnm_conv_wgrad_workspace_bytes(int64_t npos, int Ci, int Co) { return conv_wgrad_workspace(npos, Ci, Co); }

int fnm_conv_wgrad(const float* g, const float* x, int act_in, float* dwc, float* dbc, int64_t npos, int Ci, int Co,
                   void* workspace, size_t workspace_bytes, int mode, fnm_stream_t stream) {
    FNM_REQUIRE(g && x && workspace, "conv_wgrad: NULL pointer");
    if (workspace_bytes < conv_wgrad_workspace(npos, Ci, Co)) return fail(FNM_EWORKSPACE, "conv_wgrad: workspace too small");
    return conv_wgrad(g, x, act_in, dwc, dbc, npos, Ci, Co, static_cast<float*>(workspace), mode, as_stream(stream));
}

int fnm_ydft_conv_wgrad(const float* g, const float* x, const float* tab, float* T, float* dwc, float* dbc, int64_t rows, int S2,
                        int LY, int C, void* workspace, size_t workspace_bytes, int mode, fnm_stream_t stream) {
    FNM_REQUIRE(g && x && tab && T && workspace, "ydft_conv_wgrad: NULL pointer");
    FNM_REQUIRE(rows > 0 && S2 > 0 && LY > 0 && C > 0, "ydft_conv_wgrad: bad extents");
    if (workspace_bytes < conv_wgrad_workspace(rows * S2, C, C)) return fail(FNM_EWORKSPACE, "ydft_conv_wgrad: workspace too small");
    cudaStream_t st = as_stream(stream);
    if (tc_gz_pass_supported(g, x, tab, S2, rows, S2, LY, C, C, 0))
        return tc_gz_pass(g, x, tab, S2, T, dwc, dbc, rows, S2, LY, C, static_cast<float*>(workspace), mode, st);
    FNM_TRY(fnm_ydft(g, tab, T, rows, S2, LY, C, 0, mode, stream));
    return conv_wgrad(g, x, 0, dwc, dbc, rows * S2, C, C, static_cast<float*>(workspace), mode, st);
}

int fnm_gelu_fwd(const float* z, float* y, int64_t n, fnm_stream_t stream) {
    FNM_REQUIRE(z && y, "gelu_fwd: NULL pointer");
    return gelu_launch(nullptr, z, y, n, false, as_stream(stream));
}

int fnm_gelu_bwd(const float* g_y, const float* z, float* g_z, int64_t n, fnm_stream_t stream) {
    FNM_REQUIRE(g_y && z && g_z, "gelu_bwd: NULL pointer");
    return gelu_launch(g_y, z, g_z, n, true, as_stream(stream));
}

int fnm_adam_multi_tensor(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                          float* const* vmax, const int64_t* numel, const uint8_t* is_complex, double lr, double beta1,
                          double beta2, double eps, double weight_decay, int amsgrad, int step, const int32_t* step_dev,
                          fnm_stream_t stream) {
    FNM_REQUIRE(n >= 0, "adam: negative tensor count");
    if (n == 0) return FNM_OK;
    FNM_REQUIRE(p && g && m && v && numel && is_complex, "adam: NULL table");
    FNM_REQUIRE(step >= 1 || step_dev, "adam: step must be >= 1");
    return adam_launch(n, p, g, m, v, vmax, numel, is_complex, lr, beta1, beta2, eps, weight_decay, amsgrad, step,
                       step_dev, as_stream(stream));
}

int fnm_counter_inc(int32_t* counter, fnm_stream_t stream) {
    FNM_REQUIRE(counter, "counter_inc: NULL pointer");
    return counter_inc(counter, as_stream(stream));
}

// ------------------------------------------------------------------------------------------------
// F2V local branch in one pass (fc1 + GELU + trapezoid-weighted sum)
// ------------------------------------------------------------------------------------------------
size_t fnm_local_functional_workspace_bytes(int B, int P1, int P2, int K, int N) {
    const int parts = tc_local_functional_parts((int64_t)B * P1, P2, K, N);
    return parts > 0 ? align_up((size_t)B * P1 * parts * N * sizeof(float), 256) : 0;
}

int fnm_local_functional_fwd(const float* x, const float* w1, const float* b1, const float* wx, const float* wy, float* S,
                             int B, int P1, int P2, int K, int N, void* workspace, size_t workspace_bytes, int mode,
                             fnm_stream_t stream) {
    FNM_REQUIRE(x && w1 && wx && wy && S && workspace, "local_functional_fwd: NULL pointer");
    FNM_REQUIRE(B > 0 && P1 > 0 && P2 > 0 && K > 0 && N > 0, "local_functional_fwd: bad extents");
    const int64_t rows = (int64_t)B * P1;
    if (!tc_local_functional_supported(rows, P2, K, N, x)) return fail(FNM_ENOTSUP, "local_functional_fwd: shape not served");
    const int parts = tc_local_functional_parts(rows, P2, K, N);
    if (workspace_bytes < fnm_local_functional_workspace_bytes(B, P1, P2, K, N)) return fail(FNM_EWORKSPACE, "local_functional_fwd: workspace too small");
    float* partial = static_cast<float*>(workspace);
    FNM_TRY(tc_local_functional(x, w1, b1, wx, wy, nullptr, partial, rows, P1, P2, K, N, mode, as_stream(stream)));
    return batch_reduce(partial, S, B, P1 * parts, N, "local_functional_reduce", as_stream(stream));
}

int fnm_local_functional_bwd(const float* x, const float* w1, const float* b1, const float* wx, const float* wy, const float* gS,
                             float* g_h1, int B, int P1, int P2, int K, int N, int mode, fnm_stream_t stream) {
    FNM_REQUIRE(x && w1 && wx && wy && gS && g_h1, "local_functional_bwd: NULL pointer");
    FNM_REQUIRE(B > 0 && P1 > 0 && P2 > 0 && K > 0 && N > 0, "local_functional_bwd: bad extents");
    const int64_t rows = (int64_t)B * P1;
    if (!tc_local_functional_supported(rows, P2, K, N, x)) return fail(FNM_ENOTSUP, "local_functional_bwd: shape not served");
    return tc_local_functional(x, w1, b1, wx, wy, gS, g_h1, rows, P1, P2, K, N, mode, as_stream(stream));
}

// ------------------------------------------------------------------------------------------------
// fused Fourier layer
// ------------------------------------------------------------------------------------------------
size_t fnm_fourier_layer_shadow_bytes(const fnm_plan* plan) {
    if (!plan || plan->Co % 2 != 0 || plan->w_layout == FNM_W_MODE_MAJOR) return 0;
    return layer_ws(plan).shadow * sizeof(float);
}

size_t fnm_fourier_layer_workspace_bytes(const fnm_plan* plan) {
    if (!plan) return 0;
    const LayerWs w = layer_ws(plan);
    return (w.T + w.Oh + w.Z + w.Gh + w.partial + 2 * w.shadow) * sizeof(float) + 7 * 256;
}

int fnm_fourier_layer_fwd(const fnm_plan* p, const float* xin, int act_in, const float* w0, const float* w1,
                          const float* wc, const float* bc, float* z_out, float* x_act_out, float* xh_save,
                          void* workspace, size_t workspace_bytes, fnm_stream_t stream) {
    FNM_TRY(check_plan(p, true, true));
    FNM_REQUIRE(xin && w0 && z_out && xh_save && workspace, "layer_fwd: NULL pointer");
    FNM_REQUIRE(p->rows_per_w == p->R || (w1 && 2 * p->rows_per_w == p->R), "layer_fwd: weight rows do not cover R");
    if (wc && (p->S1 != p->P1 || p->S2 != p->P2))
        return fail(FNM_ENOTSUP, "layer_fwd: the pointwise branch cannot be fused with a resolution change");
    if (workspace_bytes < fnm_fourier_layer_workspace_bytes(p)) return fail(FNM_EWORKSPACE, "layer_fwd: workspace too small");
    cudaStream_t st = as_stream(stream);
    const LayerWs w = layer_ws(p);
    Carver cv(workspace, workspace_bytes);
    float* T = cv.take(w.T);
    float* Oh = cv.take(w.Oh);
    float* Z = cv.take(w.Z);
    const int LY = 2 * p->NY;
    FNM_TRY(plan_ydft_xdft(xin, p->ay, p->ay4, T, p->ex, xh_save, p->B, p->P1, p->P2, p->R, p->NY, p->Ci, act_in, p->mode, stream));
    const int mm = p->w_layout == FNM_W_MODE_MAJOR;
    const int nmodes = p->R * p->NY, m_split = p->rows_per_w * p->NY;
    if (w.shadow && mm) {
        // parameters already mode-major: the kernel reads them in place (weights1 | weights2 = two arrays)
        FNM_TRY(tc_mix_apply(w0, w1, m_split, xh_save, Oh, nmodes, p->B, p->Ci, p->Co, 0, 0, p->mode, "mix_fwd", st));
    } else if (w.shadow) {
        // mode-major weight copy: into the caller's persistent buffer when given (backward then reuses it)
        float* W1 = p->w_shadow ? p->w_shadow : cv.take(w.shadow);
        FNM_TRY(tc_mix_pack(w0, w1, p->rows_per_w, p->R, p->NY, p->Ci, p->Co, 0, W1, st));
        FNM_TRY(tc_mix_apply(W1, nullptr, 0, xh_save, Oh, nmodes, p->B, p->Ci, p->Co, 0, 0, p->mode, "mix_fwd", st));
    } else {
        FNM_TRY(simt_mix_fwd(xh_save, w0, w1, p->rows_per_w, Oh, p->B, p->R, p->NY, p->Ci, p->Co, st, mm));
    }
    FNM_TRY(fnm_xsyn(Oh, p->fx, Z, p->B, p->S1, p->R, p->NY, p->Co, p->mode, stream));
    const int zgrad = (p->flags & FNM_PLAN_ZOUT_IS_GELU_GRAD) ? 1 : 0;
    FNM_REQUIRE(!zgrad || x_act_out, "layer_fwd: FNM_PLAN_ZOUT_IS_GELU_GRAD needs x_act_out");
    return fnm_pointwise_ysyn_ex(wc ? xin : nullptr, act_in, wc, 0, bc, Z, p->by, nullptr, z_out, x_act_out,
                                 (int64_t)p->B * p->S1, p->S2, LY, p->Ci, p->Co, p->mode, zgrad ? FNM_PW_OUT_IS_GRAD : 0, p->by16, stream);
}

int fnm_fourier_layer_bwd(const fnm_plan* p, const float* g_z, const float* xin, int act_in, const float* zin_pre,
                          const float* w0, const float* w1, const float* wc, const float* xh_save, float* g_in, float* dw0, float* dw1,
                          float* dwc, float* dbc, void* workspace, size_t workspace_bytes, fnm_stream_t stream) {
    FNM_TRY(check_plan(p, true, true));
    FNM_REQUIRE(g_z && xin && w0 && xh_save && workspace, "layer_bwd: NULL pointer");
    FNM_REQUIRE(p->rows_per_w == p->R || (w1 && 2 * p->rows_per_w == p->R), "layer_bwd: weight rows do not cover R");
    if (wc && (p->S1 != p->P1 || p->S2 != p->P2))
        return fail(FNM_ENOTSUP, "layer_bwd: the pointwise branch cannot be fused with a resolution change");
    if (workspace_bytes < fnm_fourier_layer_workspace_bytes(p)) return fail(FNM_EWORKSPACE, "layer_bwd: workspace too small");
    cudaStream_t st = as_stream(stream);
    const LayerWs w = layer_ws(p);
    Carver cv(workspace, workspace_bytes);
    float* Tg = cv.take(w.T);
    float* Gxh = cv.take(w.Oh);
    float* Zg = cv.take(w.Z);
    float* Gh = cv.take(w.Gh);
    float* partial = cv.take(w.partial);
    const int LY = 2 * p->NY;
    // analysis of g_z (c_l/N folded into byT), A.2 first line
    // ... fused with the conv weight gradient when both can share one read of g_z (gz_pass_tma, fnm_tc_wgrad.cu)
    // two-call form (FNM_PLAN_BWD_*): the second call finds the gradient spectrum Gh where the first one left it
    const bool do_w = !(p->flags & FNM_PLAN_BWD_INPUT_ONLY), do_x = !(p->flags & FNM_PLAN_BWD_WEIGHTS_ONLY);
    FNM_REQUIRE(do_w || do_x, "layer_bwd: FNM_PLAN_BWD_WEIGHTS_ONLY and FNM_PLAN_BWD_INPUT_ONLY exclude each other");
    const bool want_cw = do_w && wc && (dwc || dbc);
    // table of the backward analysis as TMA can address it: byT itself, or its padded copy when S2 is no multiple of 4
    const bool s2_pad = p->S2 % 4 != 0 && p->byT4;
    const float* gtab = s2_pad ? p->byT4 : p->byT;
    const int gld = s2_pad ? ((p->S2 + 3) & ~3) : p->S2;
    const bool gz_fused = want_cw && tc_gz_pass_supported(g_z, xin, gtab, gld, (int64_t)p->B * p->S1, p->S2, LY, p->Ci, p->Co, act_in);
    if (do_w) {
        if (gz_fused) {
            FNM_TRY(tc_gz_pass(g_z, xin, gtab, gld, Tg, dwc, dbc, (int64_t)p->B * p->S1, p->S2, LY, p->Co, partial, p->mode, st));
            FNM_TRY(fnm_xdft(Tg, p->exb, Gh, p->B, p->S1, p->R, p->NY, p->Co, p->mode, stream));
        } else {
            FNM_TRY(plan_ydft_xdft(g_z, p->byT, p->byT4, Tg, p->exb, Gh, p->B, p->S1, p->S2, p->R, p->NY, p->Co, 0, p->mode, stream));
        }
    }
    float* shadowA = w.shadow ? cv.take(w.shadow) : nullptr;
    float* shadowB = w.shadow ? cv.take(w.shadow) : nullptr;
    const int mm = p->w_layout == FNM_W_MODE_MAJOR;
    const int nmodes = p->R * p->NY, m_split = p->rows_per_w * p->NY;
    if (dw0 && do_w) {
        if (w.shadow && mm) {
            FNM_TRY(tc_mix_wgrad(xh_save, Gh, dw0, dw1, m_split, nmodes, p->B, p->Ci, p->Co, p->mode, st));   // written in place
        } else if (w.shadow) {
            FNM_TRY(tc_mix_wgrad(xh_save, Gh, shadowA, nullptr, 0, nmodes, p->B, p->Ci, p->Co, p->mode, st));
            FNM_TRY(tc_mix_unpack(shadowA, dw0, dw1, p->rows_per_w, p->R, p->NY, p->Ci, p->Co, st));
        } else {
            FNM_TRY(simt_mix_bwd_w(xh_save, Gh, dw0, dw1, p->rows_per_w, p->B, p->R, p->NY, p->Ci, p->Co, st, mm));
        }
    }
    if (want_cw && !gz_fused)
        FNM_TRY(conv_wgrad(g_z, xin, act_in, dwc, dbc, (int64_t)p->B * p->S1 * p->S2, p->Ci, p->Co, partial, p->mode, st));
    if (g_in && do_x) {
        const int w1_layout = p->Co % 2 == 0 ? 1 : 0;                  // W1[m][i][o] rows readable with 16-byte loads
        if (w.shadow && mm && w1_layout) {
            FNM_TRY(tc_mix_apply(w0, w1, m_split, Gh, Gxh, nmodes, p->B, p->Co, p->Ci, 1, 1, p->mode, "mix_bwd_x", st));
        } else if (w.shadow && !mm) {
            const float* Wm = shadowB;
            if (w1_layout && p->w_shadow) Wm = p->w_shadow;       // the forward's copy, still valid: no second permute
            else FNM_TRY(tc_mix_pack(w0, w1, p->rows_per_w, p->R, p->NY, p->Ci, p->Co, w1_layout ? 0 : 1, shadowB, st));
            FNM_TRY(tc_mix_apply(Wm, nullptr, 0, Gh, Gxh, nmodes, p->B, p->Co, p->Ci, 1, w1_layout, p->mode, "mix_bwd_x", st));
        } else {
            FNM_TRY(simt_mix_bwd_x(Gh, w0, w1, p->rows_per_w, Gxh, p->B, p->R, p->NY, p->Ci, p->Co, st, mm));
        }
        FNM_TRY(fnm_xsyn(Gxh, p->fxb, Zg, p->B, p->P1, p->R, p->NY, p->Ci, p->mode, stream));
        const float* mul = zin_pre ? zin_pre : (act_in ? xin : nullptr);
        const int pre_grad = (zin_pre && (p->flags & FNM_PLAN_PRE_IS_GELU_GRAD)) ? FNM_PW_MUL_IS_GRAD : 0;
        FNM_TRY(fnm_pointwise_ysyn_ex(wc ? g_z : nullptr, 0, wc, 1, nullptr, Zg, p->ayT, mul, g_in, nullptr,
                                      (int64_t)p->B * p->P1, p->P2, LY, p->Co, p->Ci, p->mode, pre_grad, p->ayT16, stream));
    }
    return FNM_OK;
}

// ------------------------------------------------------------------------------------------------
// F2V
// ------------------------------------------------------------------------------------------------
// The mode contractions of both ends run on the tcgen05 mixing kernels when the weights are held mode-major
// (FNM_W_MODE_MAJOR, what the drop-in modules store): a vector is the same "spectrum" in every mode, and a functional is
// the cc-weighted real part of a mixed spectrum summed over modes -- so
//   F2V: out = reduce_cc( mix_fwd(x^, W) ),   dW = mix_bwd_w(x^, spread_cc(g)),   g^x = mix_bwd_x(spread_cc(g), W)
//   V2F: o^  = mix_fwd(spread(v), W),         dW = mix_bwd_w(spread(v), g^),      dv  = reduce( mix_bwd_x(g^, W) )
// Reference-ordered weights (or shapes the mixing kernels do not serve) keep the fp32 SIMT contractions.
static bool ends_on_mix(const fnm_plan* p) {
    return p->w_layout == FNM_W_MODE_MAJOR && tc_mix_supported(p->B, p->Ci, p->Co) && p->Co % 2 == 0 &&
           ((int64_t)p->B * p->Ci) % 4 == 0 && ((int64_t)p->B * p->Co) % 4 == 0;
}
static size_t ends_spectrum_floats(const fnm_plan* p) {
    const size_t cmax = p->Ci > p->Co ? p->Ci : p->Co;
    return align_up((size_t)p->R * p->NY * 2 * p->B * cmax * sizeof(float), 256) / sizeof(float);
}

size_t fnm_lfunc_workspace_bytes(const fnm_plan* p) {
    if (!p) return 0;
    const size_t LY = 2 * (size_t)p->NY;
    const size_t T = (size_t)p->B * p->P1 * LY * p->Ci;
    const size_t Gxh = (size_t)p->R * p->NY * 2 * p->B * p->Ci;
    const size_t cmax = p->Ci > p->Co ? p->Ci : p->Co;
    return (T + Gxh + 2 * ends_spectrum_floats(p)) * sizeof(float) + simt_mode_split_workspace(p->R * p->NY, p->B, (int)cmax) +
           6 * 256 + padded_table_bytes(p->P2, (int)LY);
}

int fnm_lfunc_fwd(const fnm_plan* p, const float* xin, int act_in, const float* w, const float* cc, float* out,
                  float* xh_save, void* workspace, size_t workspace_bytes, fnm_stream_t stream) {
    FNM_TRY(check_plan(p, true, false));
    FNM_REQUIRE(xin && w && cc && out && xh_save && workspace, "lfunc_fwd: NULL pointer");
    if (workspace_bytes < fnm_lfunc_workspace_bytes(p)) return fail(FNM_EWORKSPACE, "lfunc_fwd: workspace too small");
    cudaStream_t st = as_stream(stream);
    Carver cv(workspace, workspace_bytes);
    const int LY = 2 * p->NY;
    float* T = cv.take((size_t)p->B * p->P1 * LY * p->Ci);
    float* partial = cv.take(simt_mode_split_workspace(p->R * p->NY, p->B, p->Co) / sizeof(float));
    float* Oh = cv.take(ends_spectrum_floats(p));
    const int mm = p->w_layout == FNM_W_MODE_MAJOR, nmodes = p->R * p->NY;
    FNM_TRY(plan_ydft(xin, p->ay, p->ay4, T, (int64_t)p->B * p->P1, p->P2, LY, p->Ci, act_in, p->mode, stream));
    FNM_TRY(fnm_xdft(T, p->ex, xh_save, p->B, p->P1, p->R, p->NY, p->Ci, p->mode, stream));
    if (ends_on_mix(p)) {
        FNM_TRY(tc_mix_apply(w, nullptr, 0, xh_save, Oh, nmodes, p->B, p->Ci, p->Co, 0, 0, p->mode, "lfunc_contract", st));
        return mode_reduce(Oh, cc, out, nmodes, p->B * p->Co, "lfunc_reduce", st);
    }
    return simt_lfunc_contract(xh_save, w, cc, out, partial, p->B, p->R, p->NY, p->Ci, p->Co, st, mm);
}

int fnm_lfunc_bwd(const fnm_plan* p, const float* g_out, const float* xin, int act_in, const float* zin_pre,
                  const float* w, const float* cc, const float* xh_save, float* g_in, float* dw, void* workspace, size_t workspace_bytes,
                  fnm_stream_t stream) {
    FNM_TRY(check_plan(p, true, false));
    FNM_REQUIRE(g_out && xin && w && cc && xh_save && workspace, "lfunc_bwd: NULL pointer");
    if (workspace_bytes < fnm_lfunc_workspace_bytes(p)) return fail(FNM_EWORKSPACE, "lfunc_bwd: workspace too small");
    cudaStream_t st = as_stream(stream);
    Carver cv(workspace, workspace_bytes);
    const int LY = 2 * p->NY;
    float* Zg = cv.take((size_t)p->B * p->P1 * LY * p->Ci);
    float* Gxh = cv.take((size_t)p->R * p->NY * 2 * p->B * p->Ci);
    float* tabpad = LY % 4 ? cv.take(padded_table_bytes(p->P2, LY) / sizeof(float)) : nullptr;
    float* Gh = cv.take(ends_spectrum_floats(p));
    const int mm = p->w_layout == FNM_W_MODE_MAJOR, nmodes = p->R * p->NY;
    const bool on_mix = ends_on_mix(p);
    if (on_mix) FNM_TRY(mode_spread(g_out, cc, Gh, nmodes, p->B * p->Co, "lfunc_spread", st));
    if (dw) {
        if (on_mix) FNM_TRY(tc_mix_wgrad(xh_save, Gh, dw, nullptr, 0, nmodes, p->B, p->Ci, p->Co, p->mode, st, "lfunc_bwd_w"));
        else FNM_TRY(simt_lfunc_bwd_w(g_out, xh_save, cc, dw, p->B, p->R, p->NY, p->Ci, p->Co, st, mm));
    }
    if (g_in) {
        if (on_mix) FNM_TRY(tc_mix_apply(w, nullptr, 0, Gh, Gxh, nmodes, p->B, p->Co, p->Ci, 1, 1, p->mode, "lfunc_bwd_x", st));
        else FNM_TRY(simt_lfunc_bwd_x(g_out, w, cc, Gxh, p->B, p->R, p->NY, p->Ci, p->Co, st, mm));
        FNM_TRY(fnm_xsyn(Gxh, p->fxb, Zg, p->B, p->P1, p->R, p->NY, p->Ci, p->mode, stream));
        const float* mul = zin_pre ? zin_pre : (act_in ? xin : nullptr);
        const int pre_grad = (zin_pre && (p->flags & FNM_PLAN_PRE_IS_GELU_GRAD)) ? FNM_PW_MUL_IS_GRAD : 0;
        FNM_TRY(ysyn_only(Zg, p->ayT, mul, g_in, (int64_t)p->B * p->P1, p->P2, LY, p->Ci, p->mode, tabpad, stream, pre_grad));
    }
    return FNM_OK;
}

// ------------------------------------------------------------------------------------------------
// V2F
// ------------------------------------------------------------------------------------------------
size_t fnm_ldec_workspace_bytes(const fnm_plan* p) {
    if (!p) return 0;
    const size_t LY = 2 * (size_t)p->NY;
    const size_t Z = (size_t)p->B * p->S1 * LY * p->Co;
    const size_t Oh = (size_t)p->R * p->NY * 2 * p->B * p->Co;
    return (Z + Oh + 2 * ends_spectrum_floats(p)) * sizeof(float) + simt_mode_split_workspace(p->R * p->NY, p->B, p->Ci) +
           6 * 256 + padded_table_bytes(p->S2, (int)LY);
}

int fnm_ldec_fwd(const fnm_plan* p, const float* v, const float* w, float* y_out, void* workspace,
                 size_t workspace_bytes, fnm_stream_t stream) {
    FNM_TRY(check_plan(p, false, true));
    FNM_REQUIRE(v && w && y_out && workspace, "ldec_fwd: NULL pointer");
    if (workspace_bytes < fnm_ldec_workspace_bytes(p)) return fail(FNM_EWORKSPACE, "ldec_fwd: workspace too small");
    cudaStream_t st = as_stream(stream);
    Carver cv(workspace, workspace_bytes);
    const int LY = 2 * p->NY;
    float* Z = cv.take((size_t)p->B * p->S1 * LY * p->Co);
    float* Oh = cv.take((size_t)p->R * p->NY * 2 * p->B * p->Co);
    float* tabpad = LY % 4 ? cv.take(padded_table_bytes(p->S2, LY) / sizeof(float)) : nullptr;
    float* Xb = cv.take(ends_spectrum_floats(p));
    const int mm = p->w_layout == FNM_W_MODE_MAJOR, nmodes = p->R * p->NY;
    if (ends_on_mix(p)) {
        FNM_TRY(mode_spread(v, nullptr, Xb, nmodes, p->B * p->Ci, "ldec_spread", st));
        FNM_TRY(tc_mix_apply(w, nullptr, 0, Xb, Oh, nmodes, p->B, p->Ci, p->Co, 0, 0, p->mode, "ldec_mix", st));
    } else {
        FNM_TRY(simt_ldec_mix(v, w, Oh, p->B, p->R, p->NY, p->Ci, p->Co, st, mm));
    }
    FNM_TRY(fnm_xsyn(Oh, p->fx, Z, p->B, p->S1, p->R, p->NY, p->Co, p->mode, stream));
    return ysyn_only(Z, p->by, nullptr, y_out, (int64_t)p->B * p->S1, p->S2, LY, p->Co, p->mode, tabpad, stream);
}

int fnm_ldec_bwd(const fnm_plan* p, const float* g_y, const float* v, const float* w, float* g_v, float* dw,
                 void* workspace, size_t workspace_bytes, fnm_stream_t stream) {
    FNM_TRY(check_plan(p, false, true));
    FNM_REQUIRE(g_y && v && w && workspace, "ldec_bwd: NULL pointer");
    if (workspace_bytes < fnm_ldec_workspace_bytes(p)) return fail(FNM_EWORKSPACE, "ldec_bwd: workspace too small");
    cudaStream_t st = as_stream(stream);
    Carver cv(workspace, workspace_bytes);
    const int LY = 2 * p->NY;
    float* Tg = cv.take((size_t)p->B * p->S1 * LY * p->Co);
    float* Gh = cv.take((size_t)p->R * p->NY * 2 * p->B * p->Co);
    float* partial = cv.take(simt_mode_split_workspace(p->R * p->NY, p->B, p->Ci) / sizeof(float));
    float* Xb = cv.take(ends_spectrum_floats(p));
    float* Gxh = cv.take(ends_spectrum_floats(p));
    const int mm = p->w_layout == FNM_W_MODE_MAJOR, nmodes = p->R * p->NY;
    const bool on_mix = ends_on_mix(p);
    FNM_TRY(plan_ydft(g_y, p->byT, p->byT4, Tg, (int64_t)p->B * p->S1, p->S2, LY, p->Co, 0, p->mode, stream));
    FNM_TRY(fnm_xdft(Tg, p->exb, Gh, p->B, p->S1, p->R, p->NY, p->Co, p->mode, stream));
    if (dw) {
        if (on_mix) {
            FNM_TRY(mode_spread(v, nullptr, Xb, nmodes, p->B * p->Ci, "ldec_spread", st));
            FNM_TRY(tc_mix_wgrad(Xb, Gh, dw, nullptr, 0, nmodes, p->B, p->Ci, p->Co, p->mode, st, "ldec_bwd_w"));
        } else {
            FNM_TRY(simt_ldec_bwd_w(v, Gh, dw, p->B, p->R, p->NY, p->Ci, p->Co, st, mm));
        }
    }
    if (g_v) {
        if (on_mix) {
            FNM_TRY(tc_mix_apply(w, nullptr, 0, Gh, Gxh, nmodes, p->B, p->Co, p->Ci, 1, 1, p->mode, "ldec_bwd_v", st));
            FNM_TRY(mode_reduce(Gxh, nullptr, g_v, nmodes, p->B * p->Ci, "ldec_reduce", st));
        } else {
            FNM_TRY(simt_ldec_bwd_v(Gh, w, g_v, partial, p->B, p->R, p->NY, p->Ci, p->Co, st, mm));
        }
    }
    return FNM_OK;
}

}  // extern "C"
