// Internal C++ interface between the extern "C" layer (fnm_api.cu) and the kernel files.
#pragma once
#include "fnm_common.cuh"

namespace fnm {

// ---- fp32 SIMT stages (fnm_simt.cu)
int simt_ydft(const float* x, const float* tab, float* T, int64_t rows, int P2, int LY, int C, int act, cudaStream_t st);
int simt_xdft(const float* T, const float* ex, float* Xh, int B, int P1, int R, int NY, int C, cudaStream_t st);
int simt_xsyn(const float* Oh, const float* fx, float* Z, int B, int S1, int R, int NY, int C, cudaStream_t st);
// thin x-transforms of the 1-D models (2 P1, 2 R, 2 S1 <= 8): one elementwise pass instead of a table GEMM launch
bool thin_xform_supported(const float* in, const float* out, int K, int M, int C);
int thin_xdft(const float* T, const float* ex, float* Xh, int B, int P1, int R, int NY, int C, cudaStream_t st, int nsum = 1);   // nsum: partial T's to add up
int thin_xsyn(const float* Oh, const float* fx, float* Z, int B, int S1, int R, int NY, int C, cudaStream_t st);
int simt_mix_fwd(const float* Xh, const float* w0, const float* w1, int rpw, float* Oh, int B, int R, int NY, int Ci,
                 int Co, cudaStream_t st, int mode_major = 0);
int simt_mix_bwd_x(const float* Gh, const float* w0, const float* w1, int rpw, float* Gxh, int B, int R, int NY, int Ci,
                   int Co, cudaStream_t st, int mode_major = 0);
int simt_mix_bwd_w(const float* Xh, const float* Gh, float* dw0, float* dw1, int rpw, int B, int R, int NY, int Ci,
                   int Co, cudaStream_t st, int mode_major = 0);
int simt_pointwise_ysyn(const float* x, int act, const float* wc, int wc_is_kn, const float* bias, const float* Z,
                        const float* by, const float* mul, float* out, float* out_act, int64_t rows, int S2, int LY,
                        int K, int N, cudaStream_t st, int flags = 0);
size_t simt_conv_wgrad_workspace(int64_t npos, int Ci, int Co);
int simt_conv_wgrad(const float* g, const float* x, int act, float* dwc, float* dbc, int64_t npos, int Ci, int Co,
                    float* partial, cudaStream_t st);
size_t simt_mode_split_workspace(int nmodes, int B, int C);
// mm: the (Ci, Co, R, NY) weight / gradient tensor is held mode-major, memory order (R, NY, Ci, Co)
int simt_lfunc_contract(const float* Xh, const float* w, const float* cc, float* out, float* partial, int B, int R,
                        int NY, int Ci, int Co, cudaStream_t st, int mm = 0);
int simt_lfunc_bwd_x(const float* g, const float* w, const float* cc, float* Gxh, int B, int R, int NY, int Ci, int Co,
                     cudaStream_t st, int mm = 0);
int simt_lfunc_bwd_w(const float* g, const float* Xh, const float* cc, float* dw, int B, int R, int NY, int Ci, int Co,
                     cudaStream_t st, int mm = 0);
int simt_ldec_mix(const float* v, const float* w, float* Oh, int B, int R, int NY, int Ci, int Co, cudaStream_t st, int mm = 0);
int simt_ldec_bwd_w(const float* v, const float* Gh, float* dw, int B, int R, int NY, int Ci, int Co, cudaStream_t st,
                    int mm = 0);
int simt_ldec_bwd_v(const float* Gh, const float* w, float* dv, float* partial, int B, int R, int NY, int Ci, int Co,
                    cudaStream_t st, int mm = 0);
// real (B, C) matrix <-> planar per-mode spectrum (F2V / V2F ends on the mixing kernels)
int mode_spread(const float* src, const float* cc, float* dst, int nmodes, int n, const char* what, cudaStream_t st);
int mode_reduce(const float* src, const float* cc, float* out, int nmodes, int n, const char* what, cudaStream_t st);
int batch_reduce(const float* in, float* out, int B, int J, int N, const char* what, cudaStream_t st);   // out[b][n] = sum_j in[b][j][n]

// ---- optimiser / elementwise (fnm_adam.cu)
int adam_launch(int n, float* const* p, const float* const* g, float* const* m, float* const* v, float* const* vmax,
                const int64_t* numel, const uint8_t* is_complex, double lr, double beta1, double beta2, double eps,
                double wd, int amsgrad, int step, const int32_t* step_dev, cudaStream_t st);
int counter_inc(int32_t* c, cudaStream_t st);
int gelu_launch(const float* g, const float* z, float* out, int64_t n, bool grad, cudaStream_t st);

// ---- tcgen05 / TMA kernels for the heavy spatial passes (fnm_tc.cu)
int tc_available();
bool tc_ydft_supported(int64_t rows, int P2, int LY, int C);
int tc_ydft(const float* x, const float* tab, float* T, int64_t rows, int P2, int LY, int C, int act, int mode,
            cudaStream_t st);
bool tc_ydft_tma_supported(const float* x, const float* tab, int ld, int64_t rows, int P2, int LY, int C, int act, int nseg = 1);   // ld: row stride of tab
int tc_ydft_tma_segments(int64_t rows, int P2, int C);        // K segments worth using for few long rows (1 = none)
int tc_ydft_tma(const float* x, const float* tab, int ld, float* T, int64_t rows, int P2, int LY, int C, int mode, cudaStream_t st,
                int nseg = 1);                           // nseg > 1: T receives nseg PARTIAL transforms [seg][row][l'][c]
bool tc_xdft_supported(int B, int P1, int R, int NY, int C);
int tc_xdft(const float* T, const float* ex, float* Xh, int B, int P1, int R, int NY, int C, int mode, cudaStream_t st);
bool tc_resample_supported(int64_t G, int K, int M, int inner, int C);
int tc_resample(const float* in, const float* tab, float* out, int64_t G, int K, int M, int inner, int C, int mode,
                cudaStream_t st);
bool tc_xsyn_supported(int B, int S1, int R, int NY, int C);
int tc_xsyn(const float* Oh, const float* fx, float* Z, int B, int S1, int R, int NY, int C, int mode, cudaStream_t st);
bool tc_mix_supported(int B, int Ci, int Co);
size_t tc_mix_shadow_floats(int nmodes, int Ci, int Co);
int tc_mix_pack(const float* w0, const float* w1, int rpw, int R, int NY, int Ci, int Co, int which, float* dst,
                cudaStream_t st);
int tc_mix_unpack(const float* dW1, float* dw0, float* dw1, int rpw, int R, int NY, int Ci, int Co, cudaStream_t st);
int tc_mix_apply(const float* Wm, const float* Wm1, int m_split, const float* In, float* Out, int nmodes, int B, int Kt,
                 int Lt, int conj, int w_is_w1, int mode, const char* what, cudaStream_t st);
int tc_mix_wgrad(const float* Xh, const float* Gh, float* dW1, float* dW1b, int m_split, int nmodes, int B, int Ci, int Co,
                 int mode, cudaStream_t st, const char* what = "mix_bwd_w");
bool tc_conv_wgrad_supported(int64_t npos, int Ci, int Co);
size_t tc_conv_wgrad_workspace(int64_t npos, int Ci, int Co);
int tc_conv_wgrad(const float* g, const float* x, int act, float* dwc, float* dbc, int64_t npos, int Ci, int Co,
                  float* partial, int mode, cudaStream_t st);
bool tc_gz_pass_supported(const float* g, const float* x, const float* tab, int ld, int64_t rows, int S2, int LY, int Ci, int Co, int act);
int tc_gz_pass(const float* g, const float* x, const float* tab, int ld, float* T, float* dwc, float* dbc, int64_t rows, int S2, int LY,
               int C, float* partial, int mode, cudaStream_t st);
bool tc_pw2_supported(int64_t rows, int S2, int LY, int K, int N, int act, const void* x, const void* Z, const void* by,
                      int by_ld = 0);
int tc_pw2(const float* x, const float* wc, int wc_is_kn, const float* bias, const float* Z, const float* by,
           const float* mul, float* out, float* out_act, int64_t rows, int S2, int LY, int K, int N, int mode,
           cudaStream_t st, int by_ld = 0, int flags = 0, const void* by16 = nullptr);
bool tc_local_functional_supported(int64_t rows, int S2, int K, int N, const void* x);
int tc_local_functional_parts(int64_t rows, int S2, int K, int N);
int tc_local_functional(const float* x, const float* w1, const float* b1, const float* wx, const float* wy, const float* gS,
                        float* out, int64_t rows, int P1, int S2, int K, int N, int mode, cudaStream_t st);
int tc_set_sm_limit(int n);                // fnm_set_sm_limit
bool tc_pointwise_ysyn_supported(int64_t rows, int S2, int LY, int K, int N);
int tc_pointwise_ysyn(const float* x, int act, const float* wc, int wc_is_kn, const float* bias, const float* Z,
                      const float* by, const float* mul, float* out, float* out_act, int64_t rows, int S2, int LY,
                      int K, int N, int mode, cudaStream_t st);

}  // namespace fnm
