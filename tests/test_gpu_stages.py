"""GPU tests of individual C-ABI stages against float64 torch formulas (no model around them).
These pin each kernel -- in particular the tcgen05 fused pass -- independently of the layer chain."""
import ctypes as C

import pytest
import torch

from tests.helpers import TOL_FP32, TOL_TF32, rel_l2

pytestmark = pytest.mark.gpu

import fnm_b200                                   # noqa: E402
from fnm_b200._lib import check, lib              # noqa: E402

DEV = "cuda"


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _gelu64(z):
    return 0.5 * z * (1 + torch.erf(z / 2 ** 0.5))


def _gelu_grad64(z):
    return 0.5 * (1 + torch.erf(z / 2 ** 0.5)) + z * torch.exp(-0.5 * z * z) / (2 * torch.pi) ** 0.5


LY_ACTIVE = [True]


def _pointwise_ysyn_ref(x, act, wc, wc_is_kn, bias, Z, by, mul):
    """float64 statement of fnm_pointwise_ysyn (include/fnm_b200.h)."""
    out = torch.einsum("yl,rln->ryn", by.double(), Z.double()) if LY_ACTIVE[0] else 0
    if x is not None:
        xa = _gelu64(x.double()) if act else x.double()
        wk = wc.double() if wc_is_kn else wc.double().t()        # Wk[k][n]
        out = out + torch.einsum("ryk,kn->ryn", xa, wk)
    if bias is not None:
        out = out + bias.double()
    if mul is not None:
        out = out * _gelu_grad64(mul.double())
    return out


CASES = [
    # rows, S2, LY, K, N, act, wc_is_kn, bias, mul        (tcgen05-eligible shapes first)
    dict(rows=40, S2=72, LY=24, K=32, N=32, act=1, kn=0, bias=1, mul=0),       # config 1 forward
    dict(rows=40, S2=72, LY=24, K=32, N=32, act=0, kn=1, bias=0, mul=1),       # config 1 backward
    dict(rows=300, S2=288, LY=64, K=128, N=128, act=1, kn=0, bias=1, mul=0),   # config 5 forward (TMEM-resident weights)
    dict(rows=200, S2=288, LY=64, K=128, N=128, act=0, kn=1, bias=0, mul=1),   # config 5 backward
    dict(rows=7, S2=1152, LY=32, K=64, N=64, act=1, kn=0, bias=1, mul=0),      # config 2 (1-D)
    dict(rows=33, S2=144, LY=32, K=64, N=64, act=1, kn=0, bias=1, mul=0),      # config 4 trunk
    dict(rows=50, S2=57, LY=24, K=32, N=32, act=1, kn=0, bias=1, mul=1),       # config 3 trunk: odd row length
    dict(rows=20, S2=144, LY=34, K=0, N=64, act=0, kn=0, bias=0, mul=0),       # V2F decoder output (no pointwise term)
    dict(rows=20, S2=57, LY=26, K=0, N=32, act=0, kn=0, bias=0, mul=1),        # F2V backward
    # no activation on load (the trunk's normal case: the producer already wrote GELU(z)) -> TMA kernel
    dict(rows=40, S2=72, LY=24, K=32, N=32, act=0, kn=0, bias=1, mul=0),       # config 1 forward
    dict(rows=300, S2=288, LY=64, K=128, N=128, act=0, kn=0, bias=1, mul=0),   # config 5 forward (+ GELU copy)
    dict(rows=450, S2=288, LY=64, K=128, N=128, act=0, kn=1, bias=0, mul=1),   # config 5 backward, > 3 rows per CTA
    dict(rows=7, S2=1152, LY=32, K=64, N=64, act=0, kn=0, bias=1, mul=0),      # config 2 (1-D): rows split in segments
    dict(rows=33, S2=144, LY=32, K=64, N=64, act=0, kn=1, bias=0, mul=1),      # config 4 trunk backward
    dict(rows=50, S2=57, LY=24, K=32, N=32, act=0, kn=0, bias=1, mul=1),       # config 3 trunk: odd row length
    dict(rows=21, S2=100, LY=20, K=48, N=48, act=0, kn=0, bias=1, mul=0),      # channel count that is no multiple of 32
    dict(rows=64, S2=72, LY=24, K=32, N=32, act=0, kn=1, bias=0, mul=1),       # config 1 backward: two positions folded per lane row
    dict(rows=12, S2=64, LY=16, K=32, N=32, act=0, kn=0, bias=1, mul=0),       # four positions folded (C = 32, 4 x 16 mode columns)
    dict(rows=12, S2=64, LY=0, K=32, N=32, act=0, kn=0, bias=1, mul=0),        # pointwise term only, folded
    dict(rows=9, S2=40, LY=10, K=20, N=20, act=1, kn=0, bias=1, mul=0),        # SIMT-only shape (C = 20)
    dict(rows=5, S2=96, LY=16, K=96, N=96, act=1, kn=0, bias=1, mul=1),        # 64 < C < 128
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "r{rows}_s{S2}_ly{LY}_k{K}_n{N}_a{act}{kn}{bias}{mul}".format(**c))
@pytest.mark.parametrize("mode", ["fp32", "fp32-bf16corr", "fp32-bf16corr-by16", "tf32"])
def test_pointwise_ysyn(case, mode, monkeypatch):
    # fp32: the all-TF32 3-pass split (default); fp32-bf16corr: the opt-in bf16 correction passes (FNM_PW2_CORR=bf16),
    # -by16: with the table's bf16 pair precomputed
    if mode.startswith("fp32-bf16corr"):
        monkeypatch.setenv("FNM_PW2_CORR", "bf16")
        mode = "fp32-by16" if mode.endswith("by16") else "fp32"
    torch.manual_seed(0)
    rows, S2, LY, K, N = case["rows"], case["S2"], case["LY"], case["K"], case["N"]
    x = torch.randn(rows, S2, K, device=DEV) if K else None
    wc = (torch.randn(K, N, device=DEV) if case["kn"] else torch.randn(N, K, device=DEV)) / max(K, 1) ** 0.5 if K else None
    bias = torch.randn(N, device=DEV) if case["bias"] else None
    Z = torch.randn(rows, max(LY, 1), N, device=DEV)
    by = torch.randn(S2, max(LY, 1), device=DEV) / max(LY, 1) ** 0.5
    mul = torch.randn(rows, S2, N, device=DEV) if case["mul"] else None
    out = torch.full((rows, S2, N), float("nan"), device=DEV)
    out_act = torch.full((rows, S2, N), float("nan"), device=DEV) if not case["mul"] else None
    by16 = None
    if mode == "fp32-by16":                      # the table's bf16 correction operands precomputed (fnm_plan.by16)
        from fnm_b200.plan import bf16_table_pair
        by16 = bf16_table_pair(by) if LY > 0 else None
        mode = "fp32"
    check(lib().fnm_pointwise_ysyn_ex(_p(x), case["act"], _p(wc), case["kn"], _p(bias), _p(Z), _p(by), _p(mul), _p(out),
                                      _p(out_act), rows, S2, LY, K, N, 0 if mode == "fp32" else 1, 0, _p(by16), _stream()),
          "fnm_pointwise_ysyn_ex")
    torch.cuda.synchronize()
    LY_ACTIVE[0] = LY > 0
    ref = _pointwise_ysyn_ref(x, case["act"], wc, case["kn"], bias, Z, by, mul)
    LY_ACTIVE[0] = True
    assert torch.isfinite(out).all()
    err = rel_l2(out, ref)
    tol = TOL_FP32 if mode == "fp32" else TOL_TF32
    assert err < tol, err
    if out_act is not None:                       # the activated copy written by the same epilogue
        assert rel_l2(out_act, _gelu64(ref)) < tol


@pytest.mark.parametrize("case", [CASES[0], CASES[9], CASES[10], CASES[12], CASES[14], CASES[19]],
                         ids=lambda c: "r{rows}_s{S2}_ly{LY}_k{K}_n{N}_a{act}".format(**c))
@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_pointwise_ysyn_cached_gelu_derivative(case, mode):
    """fnm_pointwise_ysyn_ex: FNM_PW_OUT_IS_GRAD writes GELU'(v) | GELU(v) (forward of a fused layer), FNM_PW_MUL_IS_GRAD
    multiplies by the operand itself (backward of the consumer) -- against float64 (include/fnm_b200.h)."""
    from fnm_b200._lib import FNM_PW_MUL_IS_GRAD, FNM_PW_OUT_IS_GRAD
    torch.manual_seed(1)
    rows, S2, LY, K, N = case["rows"], case["S2"], case["LY"], case["K"], case["N"]
    x = torch.randn(rows, S2, K, device=DEV)
    wc = torch.randn(N, K, device=DEV) / K ** 0.5
    bias = torch.randn(N, device=DEV)
    Z = torch.randn(rows, max(LY, 1), N, device=DEV)
    by = torch.randn(S2, max(LY, 1), device=DEV) / max(LY, 1) ** 0.5
    imode = 0 if mode == "fp32" else 1
    tol = TOL_FP32 if mode == "fp32" else TOL_TF32
    out = torch.full((rows, S2, N), float("nan"), device=DEV)
    out_act = torch.full((rows, S2, N), float("nan"), device=DEV)
    check(lib().fnm_pointwise_ysyn_ex(_p(x), case["act"], _p(wc), 0, _p(bias), _p(Z), _p(by), None, _p(out), _p(out_act),
                                      rows, S2, LY, K, N, imode, FNM_PW_OUT_IS_GRAD, None, _stream()), "fnm_pointwise_ysyn_ex")
    torch.cuda.synchronize()
    v = _pointwise_ysyn_ref(x, case["act"], wc, 0, bias, Z, by, None)
    assert rel_l2(out_act, _gelu64(v)) < tol
    # GELU' is O(1) where v is: the derivative of an input carrying the matmul's error
    assert rel_l2(out, _gelu_grad64(v)) < tol
    # consumer side: multiply by the cached derivative
    dgrad = _gelu_grad64(v).float()
    g = torch.randn(rows, S2, N, device=DEV)
    wk = torch.randn(N, K, device=DEV) / N ** 0.5                 # [k = out channel of the layer][n = in channel]
    ZK = torch.randn(rows, max(LY, 1), K, device=DEV)
    mulK = torch.randn(rows, S2, K, device=DEV)
    gin = torch.full((rows, S2, K), float("nan"), device=DEV)
    check(lib().fnm_pointwise_ysyn_ex(_p(g), 0, _p(wk), 1, None, _p(ZK), _p(by), _p(mulK), _p(gin), None,
                                      rows, S2, LY, N, K, imode, FNM_PW_MUL_IS_GRAD, None, _stream()), "fnm_pointwise_ysyn_ex")
    torch.cuda.synchronize()
    ref = _pointwise_ysyn_ref(g, 0, wk, 1, None, ZK, by, None) * mulK.double()
    assert rel_l2(gin, ref) < tol
    del dgrad


def test_engine_selection_of_the_fused_pass():
    """Which kernel serves which call (profile names of include/fnm_b200.h: fnm_profile_begin/_end): activations that are
    already activated (every launch of the trunk since the GELU copies are cached) run on `pw2_kernel`
    ("pointwise_ysyn"); GELU ON LOAD -- callers of the C ABI without a cached copy -- keeps the first tcgen05 kernel of
    fnm_tc.cu ("pointwise_ysyn.v1"), and channel counts outside both envelopes the SIMT engine (".simt")."""
    from fnm_b200._lib import profile_begin, profile_end
    torch.manual_seed(0)
    got = {}
    for tag, case in (("tma", CASES[10]), ("v1", CASES[2]), ("simt", CASES[19])):
        rows, S2, LY, K, N = case["rows"], case["S2"], case["LY"], case["K"], case["N"]
        x, wc = torch.randn(rows, S2, K, device=DEV), torch.randn(N, K, device=DEV)
        Z, by = torch.randn(rows, LY, N, device=DEV), torch.randn(S2, LY, device=DEV)
        out = torch.empty(rows, S2, N, device=DEV)
        profile_begin()
        check(lib().fnm_pointwise_ysyn(_p(x), case["act"], _p(wc), 0, None, _p(Z), _p(by), None, _p(out), None,
                                       rows, S2, LY, K, N, 0, _stream()), "fnm_pointwise_ysyn")
        torch.cuda.synchronize()
        got[tag] = set(profile_end())
    assert got["tma"] == {"pointwise_ysyn"}, got
    assert got["v1"] == {"pointwise_ysyn.v1"}, got
    assert all(k.endswith(".simt") for k in got["simt"]) and got["simt"], got


def test_ydft_xdft_xsyn_roundtrip_stages():
    """ydft -> xdft -> xsyn -> ysyn with identity mixing reproduces the band-limited projection."""
    from fnm_b200.plan import spectral_tables
    torch.manual_seed(1)
    B, C, P1, P2, m1, m2 = 2, 32, 24, 40, 5, 7
    tab = spectral_tables(P1, P2, m1, m2)
    d = tab.device(torch.device(DEV))
    x = torch.randn(B, P1, P2, C, device=DEV)
    LY, R, NY = 2 * m2, 2 * m1, m2
    T = torch.empty(B, P1, LY, C, device=DEV)
    Xh = torch.empty(R, NY, 2, B, C, device=DEV)
    Zt = torch.empty(B, P1, LY, C, device=DEV)
    y = torch.empty(B, P1, P2, C, device=DEV)
    L = lib()
    check(L.fnm_ydft(_p(x), _p(d["ay"]), _p(T), B * P1, P2, LY, C, 0, 0, _stream()))
    check(L.fnm_xdft(_p(T), _p(d["ex"]), _p(Xh), B, P1, R, NY, C, 0, _stream()))
    check(L.fnm_xsyn(_p(Xh), _p(d["fx"]), _p(Zt), B, P1, R, NY, C, 0, _stream()))
    check(L.fnm_pointwise_ysyn(None, 0, None, 0, None, _p(Zt), _p(d["by"]), None, _p(y), None, B * P1, P2, LY, 0, C, 0, _stream()))
    xc = x.permute(0, 3, 1, 2).cpu().double()
    xh = torch.fft.rfft2(xc)
    keep = torch.zeros_like(xh)
    keep[..., :m1, :m2] = xh[..., :m1, :m2]
    keep[..., P1 - m1:, :m2] = xh[..., P1 - m1:, :m2]
    ref = torch.fft.irfft2(keep, s=(P1, P2)).permute(0, 2, 3, 1)
    assert rel_l2(y, ref) < TOL_FP32


TAB_CASES = [
    # B, P1, P2, R, NY, C                      (table-GEMM stages: ydft / xdft / xsyn)
    dict(B=3, P1=72, P2=72, R=24, NY=12, C=32),          # config 1
    dict(B=2, P1=288, P2=288, R=64, NY=32, C=128),       # config 5 (two samples)
    dict(B=5, P1=1, P2=1152, R=1, NY=16, C=64),          # config 2 (1-D)
    dict(B=3, P1=248, P2=57, R=24, NY=12, C=32),         # config 3: odd row length, table rows not 16-byte aligned
    dict(B=2, P1=144, P2=144, R=32, NY=17, C=64),        # config 4 decoder (m2 + 1 columns)
    dict(B=2, P1=24, P2=40, R=10, NY=7, C=16),           # smoke-test width
    dict(B=2, P1=20, P2=36, R=6, NY=5, C=20),            # SIMT-only channel count
    dict(B=1, P1=16, P2=24, R=8, NY=4, C=256),           # two 128-channel blocks
    # 128 channels: the TMA-fed ydft (MN-major operand) -- odd number of grid rows (last pair half empty), row length that is
    # no multiple of 32 (ragged last chunk), fewer than 64 table rows, a single chunk
    dict(B=3, P1=7, P2=100, R=6, NY=12, C=128),
    dict(B=1, P1=5, P2=36, R=4, NY=5, C=128),
    dict(B=2, P1=9, P2=28, R=2, NY=3, C=128),
    # 32 / 64 channels folded into the 128 lanes of the TMA-fed ydft (4 / 2 grid rows per group; it takes over once there are
    # at least as many groups as SMs): single groups, paired groups with a ragged last group and a half-empty last pair
    dict(B=9, P1=72, P2=72, R=24, NY=12, C=32),
    dict(B=5, P1=144, P2=144, R=32, NY=17, C=64),
    dict(B=11, P1=119, P2=40, R=8, NY=6, C=64),
    dict(B=21, P1=119, P2=36, R=8, NY=4, C=32),
    # ... and the TMA-fed xdft (even P1): one ragged chunk, two chunks, rows of the next sample behind the last chunk
    dict(B=3, P1=10, P2=24, R=6, NY=4, C=128),
    dict(B=2, P1=22, P2=32, R=8, NY=5, C=128),
    # thin x-transforms (2 P1, 2 R, 2 S1 <= 8: the 1-D models and the folded cross-dimension ones): elementwise kernel
    dict(B=130, P1=1, P2=40, R=1, NY=3, C=12),
    dict(B=7, P1=2, P2=24, R=2, NY=5, C=20),
]


@pytest.mark.parametrize("case", TAB_CASES, ids=lambda c: "b{B}_p{P1}x{P2}_r{R}_ny{NY}_c{C}".format(**c))
@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_table_stages(case, mode):
    """ydft, xdft and xsyn, each against the float64 contraction its header comment states, with random tables."""
    torch.manual_seed(2)
    B, P1, P2, R, NY, C_ = (case[k] for k in ("B", "P1", "P2", "R", "NY", "C"))
    LY, m = 2 * NY, 0 if mode == "fp32" else 1
    tol = TOL_FP32 if mode == "fp32" else TOL_TF32
    L = lib()
    # ydft, with the GELU on load
    x = torch.randn(B * P1, P2, C_, device=DEV)
    tab = torch.randn(LY, P2, device=DEV) / P2 ** 0.5
    T = torch.full((B * P1, LY, C_), float("nan"), device=DEV)
    for act in (0, 1):
        check(L.fnm_ydft(_p(x), _p(tab), _p(T), B * P1, P2, LY, C_, act, m, _stream()), "fnm_ydft")
        xa = _gelu64(x.double()) if act else x.double()
        assert rel_l2(T, torch.einsum("ly,ryc->rlc", tab.double(), xa)) < tol
    # xdft: Xh[r][l][ro][b][c] = sum_{x,ri} ex[(ro,r)][(x,ri)] T[b][x][ri][l][c]
    T5 = torch.randn(B, P1, 2, NY, C_, device=DEV)
    ex = torch.randn(2 * R, 2 * P1, device=DEV) / (2 * P1) ** 0.5
    Xh = torch.full((R, NY, 2, B, C_), float("nan"), device=DEV)
    check(L.fnm_xdft(_p(T5), _p(ex), _p(Xh), B, P1, R, NY, C_, m, _stream()), "fnm_xdft")
    ref = torch.einsum("orxi,bxilc->rlobc", ex.double().view(2, R, P1, 2), T5.double())
    assert rel_l2(Xh, ref) < tol
    # xsyn: Z[b][x][ro][l][c] = sum_{ri,r} fx[(x,ro)][(ri,r)] Oh[r][l][ri][b][c]
    S1 = P1 + 3 if P1 > 1 else 1
    Oh = torch.randn(R, NY, 2, B, C_, device=DEV)
    fx = torch.randn(2 * S1, 2 * R, device=DEV) / (2 * R) ** 0.5
    Z = torch.full((B, S1, 2, NY, C_), float("nan"), device=DEV)
    check(L.fnm_xsyn(_p(Oh), _p(fx), _p(Z), B, S1, R, NY, C_, m, _stream()), "fnm_xsyn")
    ref = torch.einsum("xoir,rlibc->bxolc", fx.double().view(S1, 2, 2, R), Oh.double())
    assert rel_l2(Z, ref) < tol


@pytest.mark.parametrize("npos,Ci,Co,act", [(72 * 72 * 3, 32, 32, 1), (288 * 150, 128, 128, 0), (288 * 37 + 5, 128, 128, 1),
                                            (1152 * 9, 64, 64, 1), (57 * 248, 32, 48, 0), (3000, 20, 20, 1),
                                            # no activation on load: operands straight from TMA as MN-major tiles (128 channels,
                                            # or 32 / 64 with positions folded into the channel axis); ragged last chunk
                                            (288 * 37 + 5, 128, 128, 0), (72 * 72 * 4, 32, 32, 0), (1152 * 8 + 2, 64, 64, 0)])
@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_conv_wgrad(npos, Ci, Co, act, mode):
    """dWc[o][i] = sum_pos g[pos][o] f(x[pos][i]), dbc[o] = sum_pos g[pos][o] against float64."""
    torch.manual_seed(3)
    g = torch.randn(npos, Co, device=DEV)
    x = torch.randn(npos, Ci, device=DEV) + 0.3
    L = lib()
    ws = torch.empty(L.fnm_conv_wgrad_workspace_bytes(npos, Ci, Co), dtype=torch.uint8, device=DEV)
    dwc = torch.full((Co, Ci), float("nan"), device=DEV)
    dbc = torch.full((Co,), float("nan"), device=DEV)
    check(L.fnm_conv_wgrad(_p(g), _p(x), act, _p(dwc), _p(dbc), npos, Ci, Co, _p(ws), ws.numel(),
                           0 if mode == "fp32" else 1, _stream()), "fnm_conv_wgrad")
    xa = _gelu64(x.double()) if act else x.double()
    tol = TOL_FP32 if mode == "fp32" else TOL_TF32
    assert rel_l2(dwc, g.double().t() @ xa) < tol
    assert rel_l2(dbc, g.double().sum(0)) < 1e-5


# rows, S2, LY: the fused kernel needs >= 2 x 148 rows and 512 x 148 positions; (40, 64, 24) takes the two-call route behind
# the same entry.  Ragged row lengths (72, 100: last chunk partly past the row), odd row counts, LY < 64 and LY % 4 != 0.
@pytest.mark.parametrize("rows,S2,LY,C", [(2048, 256, 64, 128), (1101, 72, 24, 128), (999, 100, 26, 128), (40, 64, 24, 128),
                                          (4097, 32, 8, 128),
                                          # 32 / 64 channels: 4 / 2 grid rows folded into the 128 lanes (config-1 and config-4
                                          # shapes; a row count that leaves the last group half empty)
                                          (1440, 72, 24, 32), (2880, 144, 34, 64), (4098, 72, 24, 32), (2403, 40, 16, 64)])
@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_ydft_conv_wgrad_fused(rows, S2, LY, C, mode):
    """fnm_ydft_conv_wgrad (one read of g for the backward y-DFT and the conv weight / bias gradient) against float64, and
    against the two separate entry points it replaces."""
    torch.manual_seed(11)
    g = torch.randn(rows, S2, C, device=DEV) + 0.2                       # a mean: the DC column of the transform is large
    x = torch.randn(rows, S2, C, device=DEV) + 0.3
    tab = torch.randn(LY, S2, device=DEV) / S2 ** 0.5
    L = lib()
    m = 0 if mode == "fp32" else 1
    ws = torch.empty(L.fnm_conv_wgrad_workspace_bytes(rows * S2, C, C), dtype=torch.uint8, device=DEV)
    T = torch.full((rows, LY, C), float("nan"), device=DEV)
    dwc = torch.full((C, C), float("nan"), device=DEV)
    dbc = torch.full((C,), float("nan"), device=DEV)
    check(L.fnm_ydft_conv_wgrad(_p(g), _p(x), _p(tab), _p(T), _p(dwc), _p(dbc), rows, S2, LY, C, _p(ws), ws.numel(), m, _stream()),
          "fnm_ydft_conv_wgrad")
    tol = TOL_FP32 if mode == "fp32" else TOL_TF32
    g64 = g.double()
    assert rel_l2(T, torch.einsum("ly,ryc->rlc", tab.double(), g64)) < tol
    assert rel_l2(dwc, g64.reshape(-1, C).t() @ x.double().reshape(-1, C)) < tol
    assert rel_l2(dbc, g64.reshape(-1, C).sum(0)) < 1e-5
    # the non-DC content of T at the accuracy of the separate kernel (truncating accumulation must not leak the mean)
    T2 = torch.empty_like(T)
    check(L.fnm_ydft(_p(g), _p(tab), _p(T2), rows, S2, LY, C, 0, m, _stream()), "fnm_ydft")
    assert rel_l2(T, T2.double()) < tol
    # gradient-only calls: either output may be NULL
    dw2 = torch.full((C, C), float("nan"), device=DEV)
    check(L.fnm_ydft_conv_wgrad(_p(g), _p(x), _p(tab), _p(T2), _p(dw2), None, rows, S2, LY, C, _p(ws), ws.numel(), m, _stream()),
          "fnm_ydft_conv_wgrad")
    assert torch.equal(dw2, dwc) and torch.equal(T2, T)


@pytest.mark.parametrize("B,R,NY,rpw,Ci,Co", [(20, 24, 12, 12, 32, 32), (32, 8, 32, 4, 128, 128), (64, 1, 16, 1, 64, 64),
                                              (5, 6, 5, 3, 20, 24), (40, 4, 7, 4, 48, 32), (3, 2, 3, 1, 8, 16)])
@pytest.mark.parametrize("mode", ["fp32", "tf32"])
@pytest.mark.parametrize("use_ws", [True, False])
def test_mix_stages(B, R, NY, rpw, Ci, Co, mode, use_ws):
    """compl_mul and its two backward products against complex128 einsums; with the scratch buffer the
    tcgen05 path (mode-major weight copy) runs, without it the SIMT kernels on the reference layout."""
    torch.manual_seed(4)
    two = rpw != R
    w0 = torch.randn(Ci, Co, rpw, NY, dtype=torch.complex64, device=DEV) / Ci ** 0.5
    w1 = torch.randn(Ci, Co, rpw, NY, dtype=torch.complex64, device=DEV) / Ci ** 0.5 if two else None
    W = torch.cat([w0, w1], 2) if two else w0                              # (Ci, Co, R, NY)
    Xh = torch.randn(R, NY, 2, B, Ci, device=DEV)
    Gh = torch.randn(R, NY, 2, B, Co, device=DEV)
    Xc = torch.complex(Xh[:, :, 0].double(), Xh[:, :, 1].double())          # (R, NY, B, Ci)
    Gc = torch.complex(Gh[:, :, 0].double(), Gh[:, :, 1].double())
    Wc = W.to(torch.complex128)
    L = lib()
    m = 0 if mode == "fp32" else 1
    tol = TOL_FP32 if mode == "fp32" else TOL_TF32
    nb = L.fnm_mix_workspace_bytes(B, R, NY, Ci, Co) if use_ws else 0
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=DEV) if use_ws else None
    wr = lambda t: None if t is None else _p(torch.view_as_real(t))
    Oh = torch.full((R, NY, 2, B, Co), float("nan"), device=DEV)
    check(L.fnm_mix_fwd(_p(Xh), wr(w0), wr(w1), rpw, _p(Oh), B, R, NY, Ci, Co, _p(ws), nb, m, _stream()), "fnm_mix_fwd")
    ref = torch.einsum("rlbi,iorl->rlbo", Xc, Wc)
    assert rel_l2(torch.complex(Oh[:, :, 0], Oh[:, :, 1]), ref) < tol
    Gx = torch.full((R, NY, 2, B, Ci), float("nan"), device=DEV)
    check(L.fnm_mix_bwd_x(_p(Gh), wr(w0), wr(w1), rpw, _p(Gx), B, R, NY, Ci, Co, _p(ws), nb, m, _stream()), "fnm_mix_bwd_x")
    ref = torch.einsum("rlbo,iorl->rlbi", Gc, Wc.conj())
    assert rel_l2(torch.complex(Gx[:, :, 0], Gx[:, :, 1]), ref) < tol
    dw0 = torch.full_like(w0, float("nan"))
    dw1 = torch.full_like(w1, float("nan")) if two else None
    check(L.fnm_mix_bwd_w(_p(Xh), _p(Gh), wr(dw0), wr(dw1), rpw, B, R, NY, Ci, Co, _p(ws), nb, m, _stream()), "fnm_mix_bwd_w")
    ref = torch.einsum("rlbi,rlbo->iorl", Xc.conj(), Gc)
    got = torch.cat([dw0, dw1], 2) if two else dw0
    assert rel_l2(got, ref) < tol


@pytest.mark.parametrize("B,Din,N1,N2,pad,C,ngrid", [(3, 1, 16, 24, 8, 32, 2), (2, 2, 13, 9, 4, 20, 2), (4, 1, 1, 40, 8, 64, 1),
                                                     (2, 3, 8, 8, 8, 16, 0), (2, 1, 64, 64, 8, 128, 2)])
def test_lift_kernels(B, Din, N1, N2, pad, C, ngrid):
    """fc0 on [data | grid coordinates] + zero padding, forward and weight gradients, against the torch
    statement of func_to_func2d.py:79-86 in float64."""
    torch.manual_seed(5)
    P1, P2 = (N1 + N1 // pad if N1 > 1 else 1), N2 + N2 // pad
    x = torch.randn(B, Din, N1, N2, device=DEV)
    fc0 = torch.nn.Linear(Din + ngrid, C).to(DEV)
    out = fnm_b200.ops.lift(x, fc0.weight, fc0.bias, P1, P2, ngrid)
    g = torch.randn_like(out)
    out.backward(g)
    got_dw, got_db = fc0.weight.grad.clone(), fc0.bias.grad.clone()
    fc0.zero_grad()
    feats = [x.double().permute(0, 2, 3, 1)]
    if ngrid == 2:
        feats.append(torch.linspace(0, 1, N1, device=DEV, dtype=torch.float64).view(1, N1, 1, 1).expand(B, N1, N2, 1))
    if ngrid >= 1:
        feats.append(torch.linspace(0, 1, N2, device=DEV, dtype=torch.float64).view(1, 1, N2, 1).expand(B, N1, N2, 1))
    f = torch.cat(feats, -1)
    w64, b64 = fc0.weight.detach().double().requires_grad_(True), fc0.bias.detach().double().requires_grad_(True)
    ref = torch.nn.functional.pad(f @ w64.t() + b64, [0, 0, 0, P2 - N2, 0, P1 - N1])
    ref.backward(g.double())
    assert rel_l2(out, ref) < 1e-6
    assert rel_l2(got_dw, w64.grad) < TOL_FP32 and rel_l2(got_db, b64.grad) < TOL_FP32


@pytest.mark.parametrize("shape,Ci,H,D", [((2, 20, 24), 32, 128, 1), ((3, 1, 50), 64, 128, 2), ((2, 72, 72), 128, 128, 1),
                                          ((2, 9, 11), 20, 64, 3)])
def test_projection_kernels(shape, Ci, H, D):
    """fc1 (fused pointwise kernel without spectral term) -> GELU -> fc2 (head kernel), forward and all
    gradients, against torch in float64."""
    torch.manual_seed(6)
    B, P1, P2 = shape
    x = torch.randn(B, P1, P2, Ci, device=DEV, requires_grad=True)
    fc1, fc2 = torch.nn.Linear(Ci, H).to(DEV), torch.nn.Linear(H, D).to(DEV)
    out = fnm_b200.ops.mlp_head(fnm_b200.ops.pointwise_conv(x, fc1.weight, fc1.bias), fc2.weight, fc2.bias)
    g = torch.randn_like(out)
    out.backward(g)
    got = [x.grad, fc1.weight.grad, fc1.bias.grad, fc2.weight.grad, fc2.bias.grad]
    p64 = [t.detach().double().requires_grad_(True) for t in (x, fc1.weight, fc1.bias, fc2.weight, fc2.bias)]
    h = p64[0] @ p64[1].t() + p64[2]
    ref = _gelu64(h) @ p64[3].t() + p64[4]
    ref.backward(g.double())
    assert rel_l2(out, ref) < TOL_FP32
    for a, b in zip(got, p64):
        assert rel_l2(a, b.grad) < TOL_FP32


@pytest.mark.parametrize("B,P1,P2,H", [(3, 20, 24, 128), (2, 1, 50, 64), (5, 9, 11, 32)])
def test_weighted_gelu_sum(B, P1, P2, H):
    """S[b,h] = sum wx[x] wy[y] GELU(h1) and its backward against float64 torch (the field-sized part of
    mlpfunc0 + double trapz)."""
    torch.manual_seed(8)
    h1 = torch.randn(B, P1, P2, H, device=DEV, requires_grad=True)
    wx, wy = torch.rand(P1, device=DEV), torch.rand(P2, device=DEV)
    S = fnm_b200.ops.weighted_gelu_sum(h1, wx, wy)
    g = torch.randn_like(S)
    S.backward(g)
    h64 = h1.detach().double().requires_grad_(True)
    ref = torch.einsum("x,y,bxyh->bh", wx.double(), wy.double(), _gelu64(h64))
    ref.backward(g.double())
    assert rel_l2(S, ref) < 1e-6
    assert rel_l2(h1.grad, h64.grad) < 1e-6


@pytest.mark.parametrize("B,P1,P2,C,s", [
    (3, 24, 20, 32, (12, 10)),     # even -> even: two-term form (Nyquist bin kept on one side), tcgen05 table GEMM
    (2, 29, 29, 64, (13, 13)),     # odd -> odd: one real table per axis
    (2, 145, 145, 32, (37, 37)),   # the paper's F2F grids (train_fno.py:45-46)
    (2, 16, 12, 128, (24, 18)),    # upsampling
    (2, 9, 11, 8, (14, 7)),        # odd upsampling (resize_fft's N // 2 split) on the SIMT engine
    (2, 1, 48, 32, 20),            # 1-D
    (3, 1, 21, 12, 30)])
def test_project_resolution(B, P1, P2, C, s):
    """projector1d / projector2d (shared.py:121-128, 162-169) on channels-last tensors: forward against the FFT
    statement in float64, backward against its autograd."""
    torch.manual_seed(9)
    one_d = not isinstance(s, tuple)
    x = torch.randn(B, P1, P2, C, device=DEV, requires_grad=True)
    out = fnm_b200.ops.project_resolution(x, s, one_d)
    g = torch.randn_like(out)
    out.backward(g)
    x64 = x.detach().double().requires_grad_(True)
    if one_d:
        ref = fnm_b200.shared.projector1d(x64.squeeze(1).permute(0, 2, 1), s).permute(0, 2, 1).unsqueeze(1)
    else:
        ref = fnm_b200.shared.projector2d(x64.permute(0, 3, 1, 2), s).permute(0, 2, 3, 1)
    ref.backward(g.double())
    assert out.shape == ref.shape
    assert rel_l2(out, ref) < TOL_FP32
    assert rel_l2(x.grad, x64.grad) < TOL_FP32


def test_project_resolution_vs_reference_fixtures():
    """Outputs of the reference's own projector1d / projector2d (tests/golden/projectors.npz, blocks.npz)."""
    from tests.helpers import load_npz
    g = load_npz("projectors.npz")
    b = load_npz("blocks.npz")
    g.update({"b2d.x": b["proj2.x"], "b2d.y": b["proj2.down"], "b2d.s": [8, 6], "b2u.x": b["proj2.x"], "b2u.y": b["proj2.up"],
              "b2u.s": [16, 14], "b1d.x": b["proj1.x"], "b1d.y": b["proj1.down"], "b1d.s": [7], "b1u.x": b["proj1.x"],
              "b1u.y": b["proj1.up"], "b1u.s": [20]})
    for tag in sorted({k.split(".")[0] for k in g}):
        x, y, s = torch.from_numpy(g[f"{tag}.x"]).to(DEV), g[f"{tag}.y"], [int(v) for v in g[f"{tag}.s"]]
        if x.dim() == 4:
            got = fnm_b200.ops.project_resolution(x.permute(0, 2, 3, 1).contiguous(), tuple(s), False).permute(0, 3, 1, 2)
        else:
            got = fnm_b200.ops.project_resolution(x.permute(0, 2, 1).unsqueeze(1).contiguous(), s[0], True)
            got = got.squeeze(1).permute(0, 2, 1)
        assert rel_l2(got, y) < TOL_FP32, tag


def test_gelu_kernels():
    z = torch.randn(100003, device=DEV) * 3
    g = torch.randn_like(z)
    assert rel_l2(fnm_b200.ops.gelu(z), _gelu64(z.double())) < 1e-6
    zz = z.clone().requires_grad_(True)
    fnm_b200.ops.gelu(zz).backward(g)
    assert rel_l2(zz.grad, g.double() * _gelu_grad64(z.double())) < 1e-6


@pytest.mark.parametrize("B,P1,P2,K,H", [(3, 20, 24, 32, 128), (2, 1, 50, 64, 128), (5, 9, 57, 32, 128), (2, 248, 57, 32, 128),
                                         (4, 7, 72, 32, 64)])
@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_local_functional_one_pass(B, P1, P2, K, H, mode):
    """fnm_local_functional_fwd / _bwd: fc1 + GELU + trapezoid-weighted sum in one pass (and its gradient with fc1
    recomputed) against float64 (include/fnm_b200.h; func_to_vec2d.py:100-104)."""
    torch.manual_seed(4)
    L = lib()
    imode = 0 if mode == "fp32" else 1
    tol = TOL_FP32 if mode == "fp32" else TOL_TF32
    x = torch.randn(B, P1, P2, K, device=DEV)
    w1 = torch.randn(H, K, device=DEV) / K ** 0.5
    b1 = torch.randn(H, device=DEV)
    wx = torch.rand(P1, device=DEV) + 0.5
    wy = torch.rand(P2, device=DEV) + 0.5
    nbytes = L.fnm_local_functional_workspace_bytes(B, P1, P2, K, H)
    assert nbytes > 0
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    S = torch.full((B, H), float("nan"), device=DEV)
    check(L.fnm_local_functional_fwd(_p(x), _p(w1), _p(b1), _p(wx), _p(wy), _p(S), B, P1, P2, K, H, _p(ws), ws.numel(), imode,
                                     _stream()), "fnm_local_functional_fwd")
    h1 = torch.einsum("bxyk,hk->bxyh", x.double(), w1.double()) + b1.double()
    ref = torch.einsum("x,y,bxyh->bh", wx.double(), wy.double(), _gelu64(h1))
    assert rel_l2(S, ref) < tol
    gS = torch.randn(B, H, device=DEV)
    g = torch.full((B, P1, P2, H), float("nan"), device=DEV)
    check(L.fnm_local_functional_bwd(_p(x), _p(w1), _p(b1), _p(wx), _p(wy), _p(gS), _p(g), B, P1, P2, K, H, imode, _stream()),
          "fnm_local_functional_bwd")
    gref = torch.einsum("bh,x,y,bxyh->bxyh", gS.double(), wx.double(), wy.double(), _gelu_grad64(h1))
    assert rel_l2(g, gref) < tol


@pytest.mark.parametrize("shape", [(20, 1, 64, 64), (64, 1, 1024), (128, 2), (1, 3, 7, 5), (5, 1, 12)])
def test_rel_l2_loss_kernels(shape):
    """fnm_rel_l2_fwd / _bwd through ops.rel_l2_sum against the torch statement of LpLoss(size_average=False).rel
    (util/utilities_module.py:188-204), value and gradient."""
    from fnm_b200 import ops
    from fnm_b200.parallel import rel_l2_sum
    torch.manual_seed(6)
    out = torch.randn(*shape, device=DEV, requires_grad=True)
    y = torch.randn(*shape, device=DEV)
    assert ops.rel_l2_sum_supported(out, y)
    loss = rel_l2_sum(out, y)
    (3.0 * loss).backward()
    o64 = out.detach().double().requires_grad_(True)
    b = shape[0]
    ref = (torch.norm(o64.reshape(b, -1) - y.double().reshape(b, -1), 2, 1) / (torch.norm(y.double().reshape(b, -1), 2, 1) + 1e-6)).sum()
    (3.0 * ref).backward()
    assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item())
    assert rel_l2(out.grad, o64.grad) < 2e-6
    # a sample that matches its target exactly: zero gradient (torch.norm's subgradient), not NaN
    out2 = y.clone().requires_grad_(True)
    rel_l2_sum(out2, y).backward()
    assert torch.equal(out2.grad, torch.zeros_like(out2))
