/*
 * fnm_b200.h -- C ABI of the B200-native Fourier-layer hot path (libfnm_b200.so).
 *
 * The reference (nickhnelsen/fourier-neural-mappings) has NO FFI / plugin boundary: the hot path
 * sits behind Python nn.Modules (models/shared.py) and an optimizer (util/Adam.py) that delegate
 * all arithmetic to torch.fft / einsum / conv / gelu.  This header is the boundary a native
 * replacement of those call sites binds to; each entry point cites the reference lines it
 * replaces.  Python binds it with ctypes; no torch types appear here.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 data unless stated otherwise; complex64 tensors are
 *     passed as float* to interleaved (re, im) pairs in the reference's own layout;
 *   - activations are channels-last:  X[b][x][y][c]  (1-D problems use P1 = S1 = 1, R = 1);
 *   - the library never allocates, frees or synchronises: outputs, saved tensors and scratch are
 *     caller-owned (torch caching allocator), every launch goes to the explicit `stream`;
 *   - re-entrant: no state is carried between calls (backward runs on autograd's worker thread);
 *   - return value 0 = OK, otherwise an FNM_E* code and fnm_last_error() (thread-local) explains.
 *
 * Mode plan.  All transforms are TRUNCATED dense DFTs driven by small host-built tables, so the
 * reference's mode bookkeeping (resize_rfft / resize_fft / resize_rfft2, shared.py:61-108,136-147;
 * which bins are read, where they land, which are dropped) lives entirely in the tables:
 *   R   retained kx rows of the (stacked) weight        NY  retained ky columns, LY = 2*NY
 *   ay  [LY][P2]   analysis along y   (rows: Re a_l, Im a_l;  a_l(y) = scale_in * e^{-i th_y})
 *   ayT [P2][LY]   its transpose      (backward synthesis along y)
 *   by  [S2][LY]   synthesis along y  (cols: Re b_l, -Im b_l; b_l(y) = scale_out * c_l e^{+i th_y})
 *   byT [LY][S2]   its transpose      (backward analysis along y)
 *   ex  [2R][2P1]  analysis along x, real-embedded complex (rows (ro,r), cols (x,ri))
 *   fx  [2S1][2R]  synthesis along x, real-embedded complex (rows (x,ro), cols (ri,r))
 *   exb [2R][2S1]  conj(fx)^T embedded (backward analysis along x)
 *   fxb [2P1][2R]  conj(ex)^T embedded (backward synthesis along x)
 * Intermediate layouts
 *   T  [B][P1][2][NY][C]   y-transformed rows (planar re/im)
 *   Xh [R][NY][2][B][C]    spectrum, mode-major, planar re/im   (x-hat, o-hat, g-hat alike)
 *   Z  [B][S1][2][NY][C]   x-synthesised rows
 */
#ifndef FNM_B200_H
#define FNM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fnm_stream_t;                 /* cudaStream_t */

enum {
    FNM_OK = 0,
    FNM_EINVAL = 1,                         /* bad argument / unsupported shape */
    FNM_ECUDA = 2,                          /* CUDA runtime error at launch */
    FNM_ENOTSUP = 3,                        /* valid request this build cannot serve */
    FNM_EWORKSPACE = 4                      /* workspace too small */
};

/* compute mode of the GEMM-shaped stages */
enum {
    FNM_MODE_FP32 = 0,                      /* fp32-parity mode (SIMT FMA / 3xTF32 split on tensor cores) */
    FNM_MODE_TF32 = 1                       /* single-pass TF32 tensor cores, looser stated tolerance */
};

/* memory order of the complex spectral weights w0 / w1 (and dw0 / dw1) of the fused Fourier layer */
enum {
    FNM_W_REFERENCE = 0,                    /* the reference's (Cin, Cout, kx, ky): modes fastest; each call makes a
                                               mode-major scratch copy (and permutes dW back) */
    FNM_W_MODE_MAJOR = 1                    /* (kx, ky, Cin, Cout): what the kernels contract against -- read and
                                               written in place, no copies.  The drop-in modules keep their
                                               parameters in this order behind the reference's logical shape. */
};

typedef struct fnm_plan {
    int32_t B;                              /* batch (samples on this rank) */
    int32_t P1, P2;                         /* input grid  (P1 = 1 for 1-D) */
    int32_t S1, S2;                         /* output grid (= input grid unless the resolution changes) */
    int32_t R, NY;                          /* retained weight rows / columns */
    int32_t Ci, Co;                         /* channels in / out */
    int32_t rows_per_w;                     /* weight rows held by ONE weight tensor: m1 for
                                               SpectralConv2d (weights1|weights2), R otherwise */
    int32_t mode;                           /* FNM_MODE_* */
    int32_t w_layout;                       /* FNM_W_* : memory order of the spectral weights AND their gradients
                                               (fused Fourier layer only; F2V / V2F weights are always FNM_W_REFERENCE) */
    const float *ay, *ayT, *by, *byT, *ex, *fx, *exb, *fxb;
    float* w_shadow;                        /* optional (fused Fourier layer only): caller-owned buffer of
                                               fnm_fourier_layer_shadow_bytes() that receives the mode-major copy of
                                               the spectral weights in forward; handing the SAME buffer (untouched,
                                               weights unchanged) to backward saves its weight permute.  NULL: both
                                               passes permute into the workspace. */
    int32_t flags;                          /* FNM_PLAN_* (fused Fourier layer only), 0 = none */
    const void* by16;                       /* optional bf16 pairs of by / ayT: [2][rows][LY rounded up to 8] bfloat16 =     */
    const void* ayT16;                      /* { bf16(t), bf16(t - trunc_tf32(t)) } (see fnm_pointwise_ysyn_ex); NULL: the   */
                                            /* fused pointwise kernel derives them from the fp32 tile, tile by tile          */
    const float *ay4, *byT4;                /* optional: copies of ay ([2 NY][P2]) and byT ([2 NY][S2]) whose rows are zero-  */
                                            /* padded to a multiple of 4 floats (row stride (P2 + 3) & ~3, (S2 + 3) & ~3).    */
                                            /* TMA needs 16-byte row strides: with them the TMA-fed y-DFT kernels also serve  */
                                            /* grids whose row length is no multiple of 4 (e.g. 221 x 51 padded to 248 x 57); */
                                            /* NULL: such grids take the register-path kernels                                */
} fnm_plan;

/* Cached GELU derivative (fused Fourier layer).  GELU'(z) of a layer's output costs ~25 instructions per element; in the
 * backward epilogue of the NEXT layer those instructions -- not HBM or the tensor pipe -- paced the kernel.  With
 * FNM_PLAN_ZOUT_IS_GELU_GRAD the forward pass (when x_act_out is given) writes GELU'(z) into z_out instead of z, computed
 * from the same rcp / ex2 pair as the GELU copy; the consumer's backward gets that buffer as zin_pre together with
 * FNM_PLAN_PRE_IS_GELU_GRAD and multiplies by it directly.  z itself is then never stored: nothing else needs it.   */
enum {
    FNM_PLAN_ZOUT_IS_GELU_GRAD = 1,         /* forward:  z_out := GELU'(z) (requires x_act_out != NULL) */
    FNM_PLAN_PRE_IS_GELU_GRAD = 2,          /* backward (fourier_layer_bwd, lfunc_bwd): zin_pre holds GELU'(zin), not zin */
    /* fnm_fourier_layer_bwd in two calls on the same stream, workspace and arguments, so that the caller can start the
     * exchange of the weight gradients (data parallel training) while the input gradient is still being computed:        */
    FNM_PLAN_BWD_WEIGHTS_ONLY = 4,          /* analysis of g_z, dW0 / dW1, dWc / dbc; leaves the gradient spectrum in the workspace */
    FNM_PLAN_BWD_INPUT_ONLY = 8             /* g_in from the spectrum the WEIGHTS_ONLY call left in the workspace */
};
/* flags of fnm_pointwise_ysyn_ex */
enum {
    FNM_PW_MUL_IS_GRAD = 1,                 /* out *= mul[...] instead of GELU'(mul[...]) */
    FNM_PW_OUT_IS_GRAD = 2                  /* with out_act: out := GELU'(v), out_act := GELU(v), v = the pre-activation */
};

const char* fnm_version(void);
/* sizeof(fnm_plan) as this library was compiled: a binding checks its own struct against it at load time (the plan grows
 * by optional trailing fields; a stale mirror of it would make the library read garbage pointers) */
size_t fnm_plan_sizeof(void);
const char* fnm_last_error(void);
/* 1 when the library was built with the tcgen05/TMA kernels (sm_100a) */
int fnm_has_tcgen05(void);

/* ---- SM budget of the persistent kernels.  The tcgen05 kernels launch one CTA per SM and keep it for the whole
 *      pass, so a collective enqueued on another stream (the NCCL gradient all-reduce of data-parallel training,
 *      started from inside backward) finds no SM to run on until they drain.  fnm_set_sm_limit(n) caps the grid of
 *      every later persistent launch of this process at n CTAs (0 = all SMs); it returns the previous value.
 *      Workspace queries always assume the full device.  Process-wide, takes effect at launch (or capture) time. */
int fnm_set_sm_limit(int n);

/* ---- instrumentation (used by bench.py; never on by default) --------------------------------------
 * fnm_launch_count(): kernels launched by this library in this process so far.
 * fnm_profile_begin(): start recording a CUDA-event pair around EVERY kernel launch (on the launching
 *   stream).  fnm_profile_end(buf, n): synchronise the recorded events, stop recording and write
 *   "name:launches:total_ms;..." (aggregated by kernel name) into buf; returns bytes needed.        */
int64_t fnm_launch_count(void);
int fnm_profile_begin(void);
size_t fnm_profile_end(char* buf, size_t n);

/* ---- fused Fourier layer: replaces `x = w(x) + speconv(x)` (+ the activation of the PREVIOUS
 *      layer applied on load), models/func_to_func2d.py:89-95, func_to_func1d.py:85-91,
 *      func_to_vec2d.py:92-94, vec_to_func2d.py:92-95, vec_to_vec2d.py:99-101 and 1-D twins; i.e.
 *      SpectralConv{1,2}d.forward (shared.py:193-216, 315-341) + nn.Conv{1,2}d(C,C,1) + bias, with
 *      F.gelu (shared.py:13-14) folded into the consumer's load.
 *        z[b,x,y,o] = sum_i wc[o,i] f(xin[b,x,y,i]) + bc[o] + SpectralConv(f(xin))[b,x,y,o]
 *      f = GELU(erf) if act_in else identity.  wc == NULL drops the pointwise branch (bare
 *      SpectralConv*).  xh_save [R][NY][2][B][Ci] receives the input spectrum for backward.
 *      x_act_out (may be NULL) additionally receives GELU(z): the activation is evaluated ONCE, in
 *      the epilogue of the layer that produces z, and the next layer is then called with that
 *      tensor as xin and act_in = 0 (measured on B200: GELU-on-load in every consumer costs more
 *      issue slots than the extra write costs HBM time).                                          */
size_t fnm_fourier_layer_workspace_bytes(const fnm_plan* plan);
size_t fnm_fourier_layer_shadow_bytes(const fnm_plan* plan);      /* 0: this shape does not use a weight copy */
int fnm_fourier_layer_fwd(const fnm_plan* plan, const float* xin, int act_in,
                          const float* w0, const float* w1,          /* complex (Ci,Co,rows_per_w,NY) */
                          const float* wc, const float* bc,          /* (Co,Ci), (Co) or NULL */
                          float* z_out, float* x_act_out, float* xh_save,
                          void* workspace, size_t workspace_bytes, fnm_stream_t stream);
/* backward of the above (autograd of the same reference lines).  g_z = dL/dz.  xin / act_in as
 * given to forward.  zin_pre (may be NULL): when forward was fed a cached GELU(zin_pre) as xin, the
 * pre-activation it came from; g_in is then dL/d zin_pre = dL/d xin * GELU'(zin_pre).  Outputs:
 *   g_in (times GELU'(xin) when act_in), dw0/dw1 complex grads in the reference layout, dwc (Co,Ci),
 *   dbc (Co).  Any of g_in / dw0 / dwc may be NULL to skip it.                                      */
int fnm_fourier_layer_bwd(const fnm_plan* plan, const float* g_z, const float* xin, int act_in,
                          const float* zin_pre, const float* w0, const float* w1, const float* wc, const float* xh_save,
                          float* g_in, float* dw0, float* dw1, float* dwc, float* dbc,
                          void* workspace, size_t workspace_bytes, fnm_stream_t stream);

/* ---- F2V: LinearFunctionals{1,2}d.forward (shared.py:239-254, 365-383).
 *      out[b,o] = sum_{i,r,l} cc[r,l] Re( xh[b,i,r,l] W[i,o,r,l] ),  xh = truncated DFT of f(xin)
 *      (tables carry the 1/(P1 P2) of norm="forward").  cc [R][NY] holds the 2/1/0 weights of
 *      shared.py:254,380-381.                                                                     */
size_t fnm_lfunc_workspace_bytes(const fnm_plan* plan);
int fnm_lfunc_fwd(const fnm_plan* plan, const float* xin, int act_in, const float* w, const float* cc,
                  float* out, float* xh_save, void* workspace, size_t workspace_bytes, fnm_stream_t stream);
int fnm_lfunc_bwd(const fnm_plan* plan, const float* g_out, const float* xin, int act_in,
                  const float* zin_pre, const float* w, const float* cc, const float* xh_save,
                  float* g_in, float* dw, void* workspace, size_t workspace_bytes, fnm_stream_t stream);

/* ---- V2F: LinearDecoder{1,2}d.forward (shared.py:276-289, 406-419).
 *      y[b,x,y,o] = irfft2( zero-pad( sum_i v[b,i] W[i,o,r,l] ), s )   (tables carry 1/(S1 S2))     */
size_t fnm_ldec_workspace_bytes(const fnm_plan* plan);
int fnm_ldec_fwd(const fnm_plan* plan, const float* v, const float* w, float* y_out,
                 void* workspace, size_t workspace_bytes, fnm_stream_t stream);
int fnm_ldec_bwd(const fnm_plan* plan, const float* g_y, const float* v, const float* w,
                 float* g_v, float* dw, void* workspace, size_t workspace_bytes, fnm_stream_t stream);

/* ---- stand-alone activation for a trunk that ENDS with GELU (func_to_vec2d.py:94):
 *      y = GELU(z);  g_z = g_y * GELU'(z)   (shared.py:13-14 -> F.gelu, exact erf)                */
int fnm_gelu_fwd(const float* z, float* y, int64_t n, fnm_stream_t stream);
int fnm_gelu_bwd(const float* g_y, const float* z, float* g_z, int64_t n, fnm_stream_t stream);

/* ---- complex-aware Adam, util/Adam.py:25-51 (`adam()`), one fused multi-tensor launch per <=
 *      FNM_ADAM_MAX_TENSORS tensors.  Host arrays of n device pointers.  numel[] counts FLOATS
 *      (2 per complex element); is_complex[t] != 0 selects the shared |g|^2 second moment
 *      (Adam.py:40).  vmax may be NULL unless amsgrad.  step is the 1-based step count t; if
 *      step_dev != NULL the kernel reads t from that device int32 instead (CUDA-graph replay).     */
#define FNM_ADAM_MAX_TENSORS 48
int fnm_adam_multi_tensor(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                          float* const* vmax, const int64_t* numel, const uint8_t* is_complex,
                          double lr, double beta1, double beta2, double eps, double weight_decay,
                          int amsgrad, int step, const int32_t* step_dev, fnm_stream_t stream);
/* *counter += 1 on the stream (device-side step count for graph replay) */
int fnm_counter_inc(int32_t* counter, fnm_stream_t stream);

/* ---- individual stages (exported for tests, profiling and composition) ------------------------- */
/* T[row][l'][c] = sum_y tab[l'][y] f(x[row][y][c]);  rows = B*P1, tab [LY][P2]  (rfft along y,
 * truncated: shared.py:204,326)                                                                    */
int fnm_ydft(const float* x, const float* tab, float* T, int64_t rows, int P2, int LY, int C, int act_in,
             int mode, fnm_stream_t stream);
/* Xh[r][l][ro][b][c] = sum_{x,ri} ex[(ro,r)][(x,ri)] T[b][x][ri][l][c]   (C2C along x, truncated)   */
int fnm_xdft(const float* T, const float* ex, float* Xh, int B, int P1, int R, int NY, int C, int mode,
             fnm_stream_t stream);
/* Z[b][x][ro][l][c] = sum_{ri,r} fx[(x,ro)][(ri,r)] Oh[r][l][ri][b][c]   (inverse C2C along x)       */
int fnm_xsyn(const float* Oh, const float* fx, float* Z, int B, int S1, int R, int NY, int C, int mode,
             fnm_stream_t stream);
/* one axis of projector1d / projector2d (shared.py:121-128, 162-169: rfft -> resize -> irfft) as a real
 * (M x K) resampling table applied to a channels-last tensor with trailing extent inner*C:
 *   out[g][m][j][c] = sum_k tab[m][k] in[g][k][j][c]      (backward: the same call with tab^T)              */
int fnm_resample(const float* in, const float* tab, float* out, int64_t G, int K, int M, int inner, int C, int mode,
                 fnm_stream_t stream);
/* compl_mul (shared.py:48-53) and its two backward products.  `workspace` (fnm_mix_workspace_bytes; may be
 * NULL) lets the tcgen05 path make a mode-major copy of the weights / of dW; without it the fp32 SIMT
 * kernels run on the reference layout.
 *   Oh[m][.][b][o]  = sum_i Xh[m][.][b][i] W[i][o][m]                (complex)
 *   Gxh[m][.][b][i] = sum_o Gh[m][.][b][o] conj(W[i][o][m])
 *   dW[i][o][m]     = sum_b conj(Xh[m][.][b][i]) Gh[m][.][b][o]      (written in the reference layout)        */
size_t fnm_mix_workspace_bytes(int B, int R, int NY, int Ci, int Co);
int fnm_mix_fwd(const float* Xh, const float* w0, const float* w1, int rows_per_w, float* Oh,
                int B, int R, int NY, int Ci, int Co, void* workspace, size_t workspace_bytes, int mode,
                fnm_stream_t stream);
int fnm_mix_bwd_x(const float* Gh, const float* w0, const float* w1, int rows_per_w, float* Gxh,
                  int B, int R, int NY, int Ci, int Co, void* workspace, size_t workspace_bytes, int mode,
                  fnm_stream_t stream);
int fnm_mix_bwd_w(const float* Xh, const float* Gh, float* dw0, float* dw1, int rows_per_w,
                  int B, int R, int NY, int Ci, int Co, void* workspace, size_t workspace_bytes, int mode,
                  fnm_stream_t stream);
/* fused pointwise + y-synthesis + epilogue (irfft along y fused with the 1x1 conv, bias, and the
 * GELU' of the backward pass):
 *   out[row][y][n] = ( sum_k f(x[row][y][k]) Wk[k][n] + sum_l' by[y][l'] Z[row][l'][n] + bias[n] )
 *                    * ( mul_gelu_grad_of ? GELU'(mul_gelu_grad_of[row][y][n]) : 1 )
 *   out_act[row][y][n] = GELU(out[row][y][n])      (only when out_act != NULL)
 *   Wk[k][n] = wc[n][k] (wc_is_kn == 0: forward, wc is (Co,Ci)) or wc[k][n] (wc_is_kn != 0: backward).
 *   x == NULL (or K == 0) drops the pointwise term; bias may be NULL.                               */
int fnm_pointwise_ysyn(const float* x, int act_in, const float* wc, int wc_is_kn, const float* bias,
                       const float* Z, const float* by, const float* mul_gelu_grad_of, float* out, float* out_act,
                       int64_t rows, int S2, int LY, int K, int N, int mode, fnm_stream_t stream);
/* the same with FNM_PW_* flags (cached GELU derivative, see above) and the optional bf16 pair of the table:
 * by16[2][S2][LY8] bfloat16, LY8 = LY rounded up to 8, = { bf16(by), bf16(by - trunc_tf32(by)) } zero padded.  In
 * FNM_MODE_FP32 the tcgen05 kernel multiplies as TF32 main product + two bf16 correction products; with by16 the
 * table's correction operands are TMA-loaded instead of being converted in the kernel for every tile.             */
int fnm_pointwise_ysyn_ex(const float* x, int act_in, const float* wc, int wc_is_kn, const float* bias,
                          const float* Z, const float* by, const float* mul, float* out, float* out_act,
                          int64_t rows, int S2, int LY, int K, int N, int mode, int flags, const void* by16,
                          fnm_stream_t stream);
/* dwc[o][i] = sum_pos g[pos][o] f(x[pos][i]);  dbc[o] = sum_pos g[pos][o]   (conv weight/bias grads) */
size_t fnm_conv_wgrad_workspace_bytes(int64_t npos, int Ci, int Co);
int fnm_conv_wgrad(const float* g, const float* x, int act_in, float* dwc, float* dbc, int64_t npos,
                   int Ci, int Co, void* workspace, size_t workspace_bytes, int mode, fnm_stream_t stream);
/* The two backward passes that read a layer's incoming gradient g[rows][S2][C] in full, as one read of g:
 *   T[row][l][c] = sum_y tab[l][y] g[row][y][c]  (fnm_ydft)   and   dwc / dbc of the C x C 1x1 conv (fnm_conv_wgrad, x = the
 *   layer input, no activation on load).  Backward of func_to_func2d.py:89-95 (x1 = speconv(x); x2 = w(x)).  Shapes the fused
 *   kernel does not serve run as the two separate calls.  workspace: fnm_conv_wgrad_workspace_bytes(rows * S2, C, C). */
int fnm_ydft_conv_wgrad(const float* g, const float* x, const float* tab, float* T, float* dwc, float* dbc, int64_t rows,
                        int S2, int LY, int C, void* workspace, size_t workspace_bytes, int mode, fnm_stream_t stream);

/* ---- the two ends of the trunk (adjacent ops) -----------------------------------
 * lift: fc0 applied to [data channels | grid coordinates] + zero padding at the end of each axis
 *   (models/func_to_func2d.py:79-86; grid = torch.linspace(0, 1, n), shared.py:111-118, 150-159):
 *     out[b][x][y][c] = sum_j w[c][j] feat_j(b, x, y) + bias[c]   for x < N1, y < N2, else 0
 *   x is the reference's (B, Din, N1, N2) tensor (1-D: N1 = P1 = 1); ngrid = 0 (get_grid=False), 1 or 2.
 * lift_bwd: dw[c][j], db[c] from g = dL/d out (the un-padded region).                                   */
int fnm_lift_fwd(const float* x, const float* w, const float* b, float* out, int B, int Din, int N1, int N2, int P1,
                 int P2, int C, int ngrid, fnm_stream_t stream);
size_t fnm_lift_bwd_workspace_bytes(int C, int J);
int fnm_lift_bwd(const float* g, const float* x, float* dw, float* db, int B, int Din, int N1, int N2, int P1, int P2,
                 int C, int ngrid, void* workspace, size_t workspace_bytes, fnm_stream_t stream);
/* projection head of MLP (shared.py:26-45 with act = gelu): out[pos][d] = sum_h w2[d][h] GELU(h1[pos][h]) + b2[d];
 * backward: g_h1 = (g_out w2) * GELU'(h1), dw2[d][h] = sum_pos g_out[pos][d] GELU(h1[pos][h]), db2[d].
 * (fc1 itself is fnm_pointwise_ysyn with LY = 0 and its gradient fnm_conv_wgrad.)                        */
int fnm_mlp_head_fwd(const float* h1, const float* w2, const float* b2, float* out, int64_t npos, int H, int D,
                     fnm_stream_t stream);
size_t fnm_mlp_head_bwd_workspace_bytes(int H, int D);
int fnm_mlp_head_bwd(const float* h1, const float* g_out, const float* w2, float* g_h1, float* dw2, float* db2,
                     int64_t npos, int H, int D, void* workspace, size_t workspace_bytes, fnm_stream_t stream);

/* ---- training loss of the reference harness (adjacent): LpLoss(size_average=False).rel
 *      (util/utilities_module.py:188-204):  loss = sum_b ||out_b - y_b||_2 / (||y_b||_2 + eps)  over B samples of n floats.
 *      forward also leaves coef[b] = 1 / (||out_b - y_b|| (||y_b|| + eps)) for backward:
 *      gout[b][i] = gloss * coef[b] * (out[b][i] - y[b][i]).  Sample rows that are 16-byte aligned take the float4 path. */
size_t fnm_rel_l2_workspace_bytes(int B);
int fnm_rel_l2_fwd(const float* out, const float* y, float* loss, float* coef, int B, int64_t n, float eps, void* workspace,
                   size_t workspace_bytes, fnm_stream_t stream);
int fnm_rel_l2_bwd(const float* out, const float* y, const float* coef, const float* gloss, float* gout, int B, int64_t n,
                   fnm_stream_t stream);

/* ---- F2V local branch (func_to_vec2d.py:100-104, vec_to_vec2d.py:107-111): mlpfunc0 then two
 *      torch.trapz.  fc2 is linear, so only S[b][h] = sum_{x,y} wx[x] wy[y] GELU(h1[b][x][y][h]) touches the field
 *      (h1 = fc1 output, fnm_pointwise_ysyn with LY = 0); wx / wy hold the trapezoid weights dx*(1/2, 1, ..., 1, 1/2).
 *      backward: g_h1 = gS[b][h] wx[x] wy[y] GELU'(h1).                                                           */
/* The same branch in ONE pass over the layer's activation (no (B, P1, P2, H) tensor in forward at all):
 *   S[b][h] = sum_{x,y} wx[x] wy[y] GELU( sum_i w1[h][i] x[b][x][y][i] + b1[h] )
 * fc1 runs on the tcgen05 pointwise kernel and its epilogue integrates; returns FNM_ENOTSUP for shapes that kernel does
 * not serve (the caller then uses fnm_pointwise_ysyn + fnm_wsum_gelu_fwd).  backward recomputes h1 from x and writes
 *   g_h1[b][x][y][h] = gS[b][h] wx[x] wy[y] GELU'(h1)
 * from which fnm_pointwise_ysyn (input gradient) and fnm_conv_wgrad (dW1, db1) proceed as for any pointwise layer.   */
size_t fnm_local_functional_workspace_bytes(int B, int P1, int P2, int K, int N);
int fnm_local_functional_fwd(const float* x, const float* w1, const float* b1, const float* wx, const float* wy, float* S,
                             int B, int P1, int P2, int K, int N, void* workspace, size_t workspace_bytes, int mode,
                             fnm_stream_t stream);
int fnm_local_functional_bwd(const float* x, const float* w1, const float* b1, const float* wx, const float* wy,
                             const float* gS, float* g_h1, int B, int P1, int P2, int K, int N, int mode,
                             fnm_stream_t stream);
size_t fnm_wsum_gelu_workspace_bytes(int B, int H);
int fnm_wsum_gelu_fwd(const float* h1, const float* wx, const float* wy, float* S, int B, int P1, int P2, int H,
                      void* workspace, size_t workspace_bytes, fnm_stream_t stream);
int fnm_wsum_gelu_bwd(const float* h1, const float* wx, const float* wy, const float* gS, float* g_h1, int B, int P1,
                      int P2, int H, fnm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FNM_B200_H */
