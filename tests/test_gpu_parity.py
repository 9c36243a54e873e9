"""GPU parity tests: the CUDA path (through the C ABI, via the drop-in nn.Modules) against
 (a) the golden fixtures produced by the unmodified reference, and
 (b) the CPU oracle on seeded inputs at sizes the oracle finishes in seconds.
Tolerance (north star): rel-L2 <= 1e-5 for fp32 outputs, gradients and parameters."""
import numpy as np
import pytest
import torch

from oracle import fnm_oracle as O
from tests.helpers import TOL_FP32, assert_parity, group, load_npz, rel_l2, t, to64

pytestmark = pytest.mark.gpu

import fnm_b200                                    # noqa: E402
from fnm_b200 import shared as S                   # noqa: E402

DEV = "cuda"
BLOCKS = load_npz("blocks.npz")


def _s(g, n):
    s = g["s"]
    if s[0] < 0:
        return None
    return int(s[0]) if n == 1 else tuple(int(v) for v in s)


def _load(mod, g):
    with torch.no_grad():
        for n, p in mod.named_parameters():
            p.copy_(torch.from_numpy(g[n]))
    return mod.to(DEV)


def _check_block(mod, g, fwd_kwargs):
    x = t(g["x"]).to(DEV).requires_grad_(True)
    y = mod(x, **fwd_kwargs)
    assert tuple(y.shape) == g["y"].shape
    assert rel_l2(y, g["y"]) < TOL_FP32
    y.backward(t(g["g"]).to(DEV))
    assert rel_l2(x.grad, g["dx"]) < TOL_FP32
    for n, p in mod.named_parameters():
        assert rel_l2(p.grad, g["d" + n]) < TOL_FP32, n


# ------------------------------------------------------------ golden fixtures: blocks ------
@pytest.mark.parametrize("tag", ["sc2_odd", "sc2_nyq", "sc2_rect", "sc2_down", "sc2_up", "sc2_cut"])
def test_golden_spectral_conv2d(tag):
    g = group(BLOCKS, tag)
    ci, co, m1, m2 = g["weights1"].shape
    _check_block(_load(S.SpectralConv2d(ci, co, m1, m2), g), g, {"s": _s(g, 2)})


@pytest.mark.parametrize("tag", ["sc1_even", "sc1_odd", "sc1_nyq", "sc1_down", "sc1_up"])
def test_golden_spectral_conv1d(tag):
    g = group(BLOCKS, tag)
    ci, co, m = g["weights1"].shape
    _check_block(_load(S.SpectralConv1d(ci, co, m), g), g, {"s": _s(g, 1)})


@pytest.mark.parametrize("tag", ["lf2", "lf2_odd", "lf2_pad"])
def test_golden_linear_functionals2d(tag):
    g = group(BLOCKS, tag)
    ci, co, r, c = g["weights"].shape
    mod = _load(S.LinearFunctionals2d(ci, co, r // 2, c - 1), g)
    _check_block(mod, g, {})
    assert torch.all(mod.weights.grad[:, :, r // 2:, 0] == 0)        # dead weights: exactly zero grad


@pytest.mark.parametrize("tag", ["lf1", "lf1_odd", "lf1_pad"])
def test_golden_linear_functionals1d(tag):
    g = group(BLOCKS, tag)
    ci, co, c = g["weights"].shape
    _check_block(_load(S.LinearFunctionals1d(ci, co, c - 1), g), g, {})


@pytest.mark.parametrize("tag", ["ld2", "ld2_odd", "ld2_cut", "ld2_cut_odd"])
def test_golden_linear_decoder2d(tag):
    g = group(BLOCKS, tag)
    ci, co, r, c = g["weights"].shape
    _check_block(_load(S.LinearDecoder2d(ci, co, r // 2, c - 1), g), g, {"s": _s(g, 2)})


@pytest.mark.parametrize("tag", ["ld1", "ld1_odd", "ld1_cut"])
def test_golden_linear_decoder1d(tag):
    g = group(BLOCKS, tag)
    ci, co, c = g["weights"].shape
    _check_block(_load(S.LinearDecoder1d(ci, co, c - 1), g), g, {"s": _s(g, 1)})


# ------------------------------------------------------------ golden fixtures: models ------
MODELS = {
    "fno2d": lambda: fnm_b200.FNO2d(3, 3, 8, n_layers=3, width_final=16),
    "fno2d_res": lambda: fnm_b200.FNO2d(3, 3, 8, s_outputspace=(8, 8), n_layers=2, width_final=16, d_in=2, d_out=2),
    "fno1d": lambda: fnm_b200.FNO1d(5, 8, n_layers=3, width_final=16),
    "fnf2d": lambda: fnm_b200.FNF2d(3, 3, 8, width_final=16, d_in=2, d_out=2, n_layers=3),
    "fnf1d": lambda: fnm_b200.FNF1d(5, 8, width_final=16, d_out=3, n_layers=3),
    "fnd2d": lambda: fnm_b200.FND2d(5, (16, 16), modes1=3, modes2=3, width=8, width_initial=16, width_final=16,
                                    n_layers=3),
    "fnd1d": lambda: fnm_b200.FND1d(5, 32, modes1=5, width=8, width_initial=16, width_final=16, n_layers=3),
    "fnn2d": lambda: fnm_b200.FNN2d(5, 3, s_latentspace=(16, 16), modes1=3, modes2=3, width=8, width_initial=16,
                                    width_final=16, n_layers=4),
    "fnn1d": lambda: fnm_b200.FNN1d(5, 3, s_latentspace=32, modes1=5, width=8, width_initial=16, width_final=16,
                                    n_layers=3),
    # resolution change at odd sizes and the cross-dimension models
    "fno1d_res": lambda: fnm_b200.FNO1d(5, 8, s_outputspace=21, n_layers=3, width_final=16),
    "fno2d_res_odd": lambda: fnm_b200.FNO2d(3, 3, 8, s_outputspace=(11, 23), n_layers=2, width_final=16),
    "fno1d2": lambda: fnm_b200.FNO1d2(4, 12, 3, 3, 8, width_final=16, n_layers=3),
    "fno1d2_res": lambda: fnm_b200.FNO1d2(4, 12, 3, 3, 8, s_outputspace=(16, 11), width_final=16, n_layers=3, d_in=2,
                                          d_out=2),
    "fno2d1": lambda: fnm_b200.FNO2d1(4, 12, 3, 3, 8, width_final=16, n_layers=3),
    "fno2d1_res": lambda: fnm_b200.FNO2d1(4, 12, 3, 3, 8, s_outputspace=24, width_final=16, n_layers=3, d_in=2,
                                          d_out=2),
}


ORACLE_FWD = {
    "fno2d": lambda p, x: O.f2f_forward(p, x),
    "fno2d_res": lambda p, x: O.f2f_forward(p, x, s_outputspace=(8, 8)),
    "fno1d": lambda p, x: O.f2f_forward(p, x),
    "fnf2d": lambda p, x: O.f2v_forward(p, x),
    "fnf1d": lambda p, x: O.f2v_forward(p, x),
    "fnd2d": lambda p, x: O.v2f_forward(p, x, (16, 16)),
    "fnd1d": lambda p, x: O.v2f_forward(p, x, 32),
    "fnn2d": lambda p, x: O.v2v_forward(p, x, (16, 16)),
    "fnn1d": lambda p, x: O.v2v_forward(p, x, 32),
    "fno1d_res": lambda p, x: O.f2f_forward(p, x, s_outputspace=21),
    "fno2d_res_odd": lambda p, x: O.f2f_forward(p, x, s_outputspace=(11, 23)),
    "fno1d2": lambda p, x: O.f1d_to_f2d_forward(p, x),
    "fno1d2_res": lambda p, x: O.f1d_to_f2d_forward(p, x, s_outputspace=(16, 11)),
    "fno2d1": lambda p, x: O.f2d_to_f1d_forward(p, x),
    "fno2d1_res": lambda p, x: O.f2d_to_f1d_forward(p, x, s_outputspace=24),
}


@pytest.mark.parametrize("name", sorted(MODELS))
def test_golden_model_train_steps(name):
    """Reference weights in, then 3 x (forward, rel-L2 sum loss, backward, fused Adam): output, loss,
    every gradient and every parameter after 3 steps must match the reference run (fixtures); a
    float64 run of the oracle supplies the truth for the noise-floor rule of assert_parity."""
    d = load_npz(f"model_{name}.npz")
    model = MODELS[name]()
    missing = model.load_state_dict({k: torch.from_numpy(v) for k, v in group(d, "w0").items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model = model.to(DEV)
    opt = fnm_b200.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    x, y = t(d["x"]).to(DEV), t(d["y"]).to(DEV)
    # float64 truth of the same 3 steps
    p64 = {k: to64(v).requires_grad_(True) for k, v in group(d, "w0").items()}
    st64 = O.AdamState()
    g64 = {}
    for step in range(3):
        loss64, out64 = O.train_step(ORACLE_FWD[name], p64, to64(d["x"]), to64(d["y"]), st64, lr=1e-3, weight_decay=1e-4)
        if step == 0:
            out64_0, loss64_0 = out64, loss64
            g64 = {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v)) for k, v in p64.items()}
    for step in range(3):
        opt.zero_grad()
        out = model(x)
        loss = fnm_b200.parallel.rel_l2_sum(out, y)
        loss.backward()
        if step == 0:
            assert tuple(out.shape) == d["out0"].shape
            assert_parity(out, d["out0"], out64_0, what="out")
            assert_parity(loss, d["loss0"], loss64_0, what="loss")
            for k, p in model.named_parameters():
                got = p.grad if p.grad is not None else torch.zeros_like(p)
                assert_parity(got, d[f"g0.{k}"], g64[k], what=f"grad {k}")
        opt.step()
    assert_parity(loss, d["loss_last"], loss64, what="loss after 3 steps")
    for k, v in model.state_dict().items():
        assert_parity(v, d[f"w3.{k}"], p64[k], what=f"param {k}")


@pytest.mark.parametrize("tag,kw", [("plain", dict(lr=1e-3)), ("wd", dict(lr=2e-3, weight_decay=1e-2)),
                                    ("ams", dict(lr=1e-3, weight_decay=1e-4, amsgrad=True))])
def test_golden_adam(tag, kw):
    d = group(load_npz("adam.npz"), tag)
    use_c = tag != "ams"
    pc = torch.nn.Parameter(t(d["pc0"]).to(DEV))
    pr = torch.nn.Parameter(t(d["pr0"]).to(DEV))
    opt = fnm_b200.Adam([pc, pr] if use_c else [pr], **kw)
    for step in range(4):
        pc.grad = t(d[f"gc{step}"]).to(DEV) if use_c else None
        pr.grad = t(d[f"gr{step}"]).to(DEV)
        opt.step()
        if use_c:
            assert rel_l2(pc, d[f"pc{step + 1}"]) < 1e-6
        assert rel_l2(pr, d[f"pr{step + 1}"]) < 1e-6
    if use_c:
        assert rel_l2(opt.state[pc]["exp_avg"], d["mc"]) < 1e-6
        assert rel_l2(opt.state[pc]["exp_avg_sq"], d["vc"]) < 1e-6


def test_adam_amsgrad_rejects_complex_like_reference():
    pc = torch.nn.Parameter(torch.ones(2, 2, dtype=torch.cfloat, device=DEV))
    pc.grad = torch.ones_like(pc)
    with pytest.raises(RuntimeError):
        fnm_b200.Adam([pc], amsgrad=True).step()


def test_adam_differs_from_torch_adam():
    """The reference rule (|g|^2 shared moment) is NOT torch.optim.Adam's (separate re/im moments)."""
    torch.manual_seed(0)
    p0 = torch.randn(64, dtype=torch.cfloat, device=DEV)
    a, b = torch.nn.Parameter(p0.clone()), torch.nn.Parameter(p0.clone())
    oa, ob = fnm_b200.Adam([a], lr=1e-2), torch.optim.Adam([b], lr=1e-2)
    for i in range(3):
        g = torch.randn(64, dtype=torch.cfloat, device=DEV)
        a.grad, b.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    assert rel_l2(a, b) > 1e-4


# ------------------------------------------------------------ oracle at BASELINE-like sizes -
def _oracle_layer(x, w1, w2, wc, bc, act):
    y = O.pointwise_conv(x, wc, bc) + (O.spectral_conv1d(x, w1) if w2 is None else O.spectral_conv2d(x, w1, w2))
    return torch.nn.functional.gelu(y) if act else y


@pytest.mark.parametrize("shape", [
    dict(B=3, C=32, P=(72, 72), m=(12, 12)),          # config 1 layer
    dict(B=2, C=32, P=(62, 57), m=(12, 12)),          # odd / non-multiple-of-4 grid (config 3 flavour)
    dict(B=2, C=64, P=(36, 40), m=(16, 16)),          # config 4 width
    dict(B=2, C=20, P=(24, 20), m=(5, 7)),            # channel count that is no multiple of 16
    dict(B=4, C=64, P=(1152,), m=(16,)),              # config 2 layer (1-D)
    dict(B=2, C=24, P=(45,), m=(9,)),                 # odd 1-D grid
    dict(B=5, C=128, P=(20, 33), m=(4, 9)),           # batch / mode counts that are no multiples of 4 on the tcgen05 kernels
    dict(B=1, C=16, P=(17, 30), m=(3, 5)),            # single sample, 16 channels (8 rows folded into the 128 lanes)
    dict(B=3, C=96, P=(18, 24), m=(7, 6)),            # 96 channels: not a power of two, above 64
    # 1-D, few long rows: the y-DFT splits every row into K segments whose partial transforms the thin x-transform adds up
    dict(B=64, C=64, P=(1152,), m=(16,)),             # config 2 at its own batch (4 segments of 9 chunks)
    dict(B=3, C=128, P=(300,), m=(20,)),              # last segment runs past the end of the row
    dict(B=5, C=32, P=(203,), m=(9,)),                # row length that is no multiple of 4 (padded table), ragged row group
])
def test_fused_layers_vs_oracle(shape):
    """Two chained fused layers (GELU between them is applied on load) + trailing GELU, forward and
    all gradients, against the CPU oracle."""
    check_fused_layers(shape["B"], shape["C"], shape["P"], shape["m"], seed=5)


def check_fused_layers(B, C, P, m, seed=5):
    """Shared by the test above and tools/shape_sweep.py (random shapes)."""
    torch.manual_seed(seed)
    one_d = len(P) == 1
    Sc = (lambda: S.SpectralConv1d(C, C, m[0])) if one_d else (lambda: S.SpectralConv2d(C, C, m[0], m[1]))
    Conv = torch.nn.Conv1d if one_d else torch.nn.Conv2d
    scs = torch.nn.ModuleList([Sc(), Sc()])
    ws = torch.nn.ModuleList([Conv(C, C, 1), Conv(C, C, 1)])
    with torch.no_grad():
        for sc in scs:                   # O(1) spectral branch so both branches matter
            for p in sc.parameters():
                p.mul_(C)
    x_cpu = torch.randn(B, C, *P)
    g_cpu = torch.randn(B, C, *P)
    # oracle (CPU fp32)
    xo = x_cpu.clone().requires_grad_(True)
    po = {f"{k}": v.detach().clone().requires_grad_(True) for k, v in
          list(scs.named_parameters(prefix="sc")) + list(ws.named_parameters(prefix="w"))}
    h = xo
    for i in range(2):
        w2 = None if one_d else po[f"sc.{i}.weights2"]
        h = _oracle_layer(h, po[f"sc.{i}.weights1"], w2, po[f"w.{i}.weight"], po[f"w.{i}.bias"], True)
    h.backward(g_cpu)
    # CUDA path
    from fnm_b200.models._trunk import Pending, run_layers
    scs, ws = scs.to(DEV), ws.to(DEV)
    xg = x_cpu.to(DEV).requires_grad_(True)
    xcl = xg.permute(0, 2, 1).unsqueeze(1) if one_d else xg.permute(0, 2, 3, 1)
    state = run_layers(Pending(xcl), scs, ws, "gelu", [True, True], one_d)
    out = state.value()
    out = out.squeeze(1).permute(0, 2, 1) if one_d else out.permute(0, 3, 1, 2)
    assert rel_l2(out, h) < TOL_FP32
    out.backward(g_cpu.to(DEV))
    assert rel_l2(xg.grad, xo.grad) < TOL_FP32
    got = dict(list(scs.named_parameters(prefix="sc")) + list(ws.named_parameters(prefix="w")))
    for k, v in po.items():
        assert rel_l2(got[k].grad, v.grad) < TOL_FP32, k


@pytest.mark.parametrize("shape", [
    dict(B=12, C=32, P=(248, 57), m=(12, 12)),        # config 3 grid: 57-point rows, four rows folded into the lanes
    dict(B=10, C=64, P=(130, 54), m=(9, 14)),         # 64 channels, rows of 54
    dict(B=5, C=128, P=(130, 61), m=(8, 20)),         # 128 channels, odd row length
])
def test_fused_layers_ragged_rows_on_the_tma_kernels(shape):
    """Row lengths that are no multiple of 4 with enough rows for the TMA-fed y-DFT kernels: the plan hands them the
    analysis tables with rows padded to 16-byte strides (fnm_plan.ay4 / byT4).  Same check as above, plus: the backward
    y-DFT really ran fused with the conv weight gradient (profile name `ydft_conv_wgrad`)."""
    from fnm_b200._lib import profile_begin, profile_end
    profile_begin()
    try:
        check_fused_layers(shape["B"], shape["C"], shape["P"], shape["m"], seed=9)
    finally:
        torch.cuda.synchronize()
        names = profile_end()
    assert "ydft_conv_wgrad" in names, sorted(names)
    assert not any(k.endswith(".simt") for k in names), sorted(names)


@pytest.mark.parametrize("act", ["gelu", "relu", "tanh"])
def test_full_model_config1_shape_vs_oracle(act):
    """FNO2d at the config-1 architecture (64x64, modes 12, width 32, 4 layers), batch 4."""
    torch.manual_seed(0)
    model = fnm_b200.FNO2d(12, 12, 32, act=act)
    p = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    x = torch.randn(4, 1, 64, 64, generator=torch.Generator().manual_seed(1234))
    y = torch.randn(4, 1, 64, 64, generator=torch.Generator().manual_seed(1235))
    ref = O.f2f_forward(p, x, act=act)
    O.rel_l2_sum(ref, y).backward()
    model = model.to(DEV)
    out = model(x.to(DEV))
    fnm_b200.parallel.rel_l2_sum(out, y.to(DEV)).backward()
    assert rel_l2(out, ref) < TOL_FP32
    # ReLU's derivative is discontinuous: two correct fp32 evaluations disagree on the gate of the few
    # pre-activations that lie within rounding error of zero, and every flipped gate changes a gradient
    # term by 100 %.  Smooth activations are held to 1e-5; ReLU gradients to 5e-4.
    tol = 5e-4 if act == "relu" else TOL_FP32
    for k, q in model.named_parameters():
        assert rel_l2(q.grad, p[k].grad) < tol, k


def test_extra_leading_dims_like_reference_einsum():
    """The `...` of compl_mul (shared.py:53): extra dims between channels and space act as batch."""
    torch.manual_seed(1)
    sc = S.SpectralConv1d(3, 4, 5)
    x = torch.randn(2, 3, 6, 16)
    ref = O.spectral_conv1d(x, sc.weights1.detach())
    got = sc.to(DEV)(x.to(DEV))
    assert tuple(got.shape) == tuple(ref.shape)
    assert rel_l2(got, ref) < TOL_FP32
    ld = S.LinearDecoder1d(3, 4, 5)
    v = torch.randn(2, 3, 7)
    ref = O.linear_decoder1d(v, ld.weights.detach(), 12)
    got = ld.to(DEV)(v.to(DEV), 12)
    assert tuple(got.shape) == tuple(ref.shape)
    assert rel_l2(got, ref) < TOL_FP32
    lf = S.LinearFunctionals1d(3, 4, 5)
    xx = torch.randn(2, 3, 6, 16)
    ref = O.linear_functionals1d(xx, lf.weights.detach())
    got = lf.to(DEV)(xx.to(DEV))
    assert tuple(got.shape) == tuple(ref.shape)
    assert rel_l2(got, ref) < TOL_FP32


def test_size_independent_properties_at_full_size():
    """At the config-5 layer size (C=128, 288x288, modes 32; batch 2): linearity of the spectral
    branch and the adjoint identity <L x, g> = <x, L^T g> that ties forward to backward."""
    torch.manual_seed(2)
    C = 128
    sc = S.SpectralConv2d(C, C, 32, 32).to(DEV)
    with torch.no_grad():
        sc.weights1.mul_(C)
        sc.weights2.mul_(C)
    x1 = torch.randn(2, C, 288, 288, device=DEV)
    x2 = torch.randn(2, C, 288, 288, device=DEV)
    with torch.no_grad():
        lhs = sc(2.0 * x1 - 3.0 * x2)
        rhs = 2.0 * sc(x1) - 3.0 * sc(x2)
    assert rel_l2(lhs, rhs) < TOL_FP32
    xa = x1.clone().requires_grad_(True)
    y = sc(xa)
    g = torch.randn_like(y)
    y.backward(g)
    a = (y.detach().double() * g.double()).sum().item()
    b = (xa.grad.double() * x1.double()).sum().item()
    assert abs(a - b) <= 1e-5 * max(abs(a), abs(b), 1.0)


def test_config5_full_size_train_step_vs_oracle():
    """BASELINE config 5 at its FULL layer size (FNO2d 256x256, modes 32, width 128, 4 layers; 268 M real
    parameters), batch 2: forward output, loss and every gradient of one train step against the CPU oracle.
    This is the shape the headline benchmark runs (TMEM-resident 128-channel weights, 5 x 64-position
    tiles per grid row, streamed x-tables, mode-major mixing, folded/flushed weight-gradient chains)."""
    torch.manual_seed(0)
    model = fnm_b200.FNO2d(32, 32, 128)
    p = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    x = torch.randn(2, 1, 256, 256, generator=torch.Generator().manual_seed(1234))
    y = torch.randn(2, 1, 256, 256, generator=torch.Generator().manual_seed(1235))
    ref = O.f2f_forward(p, x)
    loss_ref = O.rel_l2_sum(ref, y)
    loss_ref.backward()
    model = model.to(DEV)
    out = model(x.to(DEV))
    loss = fnm_b200.parallel.rel_l2_sum(out, y.to(DEV))
    loss.backward()
    assert rel_l2(out, ref) < TOL_FP32
    assert rel_l2(loss, loss_ref) < TOL_FP32
    late = [k for k, q in model.named_parameters() if rel_l2(q.grad, p[k].grad) >= TOL_FP32]
    if late:
        # spectral-weight gradients of a mean-dominated field are ill-conditioned in fp32 (tests/helpers.py:
        # assert_parity): judge those against the float64 evaluation of the same algorithm
        p64 = {k: v.detach().to(torch.complex128 if v.is_complex() else torch.float64).requires_grad_(True)
               for k, v in p.items()}
        O.rel_l2_sum(O.f2f_forward(p64, x.double()), y.double()).backward()
        grads = dict(model.named_parameters())
        for k in late:
            assert_parity(grads[k].grad, p[k].grad, p64[k].grad, what=k)


@pytest.mark.parametrize("family", ["fno2d", "fnf2d", "fnd2d"])
def test_direct_gradient_sinks_match_autograd_accumulation(family):
    """FlatGradients(direct_spectral=True): the backward kernels write the spectral-weight gradients straight into
    the flat buffer (no AccumulateGrad pass).  Gradients and parameters after two Adam steps must be bit-identical
    to the default path, and a second use of a weight without zero_() must raise."""
    ctor = {"fno2d": lambda: fnm_b200.FNO2d(4, 4, 32, n_layers=3, width_final=32),
            "fnf2d": lambda: fnm_b200.FNF2d(4, 4, 32, width_final=32, d_out=3, n_layers=3),
            "fnd2d": lambda: fnm_b200.FND2d(6, (24, 24), modes1=4, modes2=4, width=32, width_initial=32, width_final=32,
                                            n_layers=3)}[family]
    torch.manual_seed(3)
    x = torch.randn(3, 6, device=DEV) if family == "fnd2d" else torch.randn(3, 1, 24, 24, device=DEV)
    results = []
    for direct in (False, True):
        torch.manual_seed(11)
        model = ctor().to(DEV)
        opt = fnm_b200.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
        grads = fnm_b200.parallel.FlatGradients(model.parameters(), direct_spectral=direct)
        y = None
        for _ in range(2):
            grads.zero_()
            out = model(x)
            y = torch.ones_like(out) if y is None else y
            fnm_b200.parallel.rel_l2_sum(out, y).backward()
            g = {k: p.grad.clone() for k, p in model.named_parameters()}
            opt.step()
        results.append((g, {k: v.clone() for k, v in model.state_dict().items()}))
        if direct:
            assert all(p._fnm_sink_uses == 1 for p in model.parameters() if p.is_complex())   # the sinks were used
            with pytest.raises(RuntimeError):
                model(x)                                       # second use without zero_()
            with torch.no_grad():
                model(x)                                       # inference does not touch the sinks
    (g0, w0), (g1, w1) = results
    for k in g0:
        assert torch.equal(torch.view_as_real(g0[k]) if g0[k].is_complex() else g0[k],
                           torch.view_as_real(g1[k]) if g1[k].is_complex() else g1[k]), k
    for k in w0:
        assert torch.equal(torch.view_as_real(w0[k]) if w0[k].is_complex() else w0[k],
                           torch.view_as_real(w1[k]) if w1[k].is_complex() else w1[k]), k


def test_two_call_backward_and_partial_adam_steps_are_bit_identical():
    """What the overlapped data-parallel step adds on each rank, checked on one GPU: (a) fnm_fourier_layer_bwd in two
    calls (FNM_PLAN_BWD_WEIGHTS_ONLY, then FNM_PLAN_BWD_INPUT_ONLY -- the exchange of the layer's dW starts between
    them), (b) one optimizer step spread over several Adam.step(params=...) calls.  Gradients and parameters after two
    steps must equal the single-call path bit for bit, eagerly and with Adam(capturable=True)."""
    torch.manual_seed(5)
    x = torch.randn(3, 1, 24, 24, device=DEV)
    results = []
    for split, capturable in ((False, False), (True, False), (True, True)):
        torch.manual_seed(11)
        model = fnm_b200.FNO2d(4, 4, 32, n_layers=3, width_final=32).to(DEV)
        opt = fnm_b200.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, capturable=capturable)
        grads = fnm_b200.parallel.FlatGradients(model.parameters(), direct_spectral=True, overlap=split)
        staged = []
        if split:
            assert grads.overlap
            grads._sink_ready = lambda off, n, q: staged.append(q)      # single process: record instead of exchanging
        y = None
        for _ in range(2):
            grads.zero_()
            out = model(x)
            y = torch.ones_like(out) if y is None else y
            fnm_b200.parallel.rel_l2_sum(out, y).backward()
            g = {k: p.grad.clone() for k, p in model.named_parameters()}
            if split:
                cplx = [p for p in model.parameters() if p.is_complex()]
                rest = [p for p in model.parameters() if not p.is_complex()]
                opt.step(params=cplx[2:])
                opt.step(params=cplx[:2], continued=True)
                opt.step(params=rest, continued=True)
            else:
                opt.step()
        if capturable:
            opt.sync_steps()
        assert {int(opt.state[p]["step"]) for p in model.parameters()} == {2}
        results.append((g, {k: v.clone() for k, v in model.state_dict().items()}))
        if split:
            assert len(staged) == 2 * sum(p.is_complex() for p in model.parameters())
    fr = lambda t: torch.view_as_real(t) if t.is_complex() else t
    for g1, w1 in results[1:]:
        for k in results[0][0]:
            assert torch.equal(fr(results[0][0][k]), fr(g1[k])), k
        for k in results[0][1]:
            assert torch.equal(fr(results[0][1][k]), fr(w1[k])), k


def test_spectral_weights_live_mode_major_and_skip_the_layout_copies():
    """The drop-in modules keep weights1 / weights2 in (kx, ky, Cin, Cout) memory order behind the reference's
    logical shape: a train step launches no weight permutes, gradients and Adam state share the order, and the
    results equal those of the same model with reference-ordered (contiguous) parameters, which take the
    pack / unpack path of the C ABI."""
    from fnm_b200 import _lib
    from fnm_b200.ops import _is_mode_major
    torch.manual_seed(5)
    ref_sd = fnm_b200.FNO2d(4, 4, 32, n_layers=3, width_final=32).state_dict()
    x = torch.randn(3, 1, 24, 24, device=DEV)
    y = torch.randn(3, 1, 24, 24, device=DEV)
    outs = {}
    for order in ("mode_major", "reference"):
        model = fnm_b200.FNO2d(4, 4, 32, n_layers=3, width_final=32)
        model.load_state_dict(ref_sd)
        model = model.to(DEV)
        spectral = [p for p in model.parameters() if p.is_complex()]
        if order == "reference":
            for sc in model.speconvs:
                sc.weights1 = torch.nn.Parameter(sc.weights1.detach().contiguous())
                sc.weights2 = torch.nn.Parameter(sc.weights2.detach().contiguous())
        else:
            assert all(_is_mode_major(p) and not p.is_contiguous() for p in spectral)     # survives load + .to()
        opt = fnm_b200.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
        _lib.profile_begin()
        for _ in range(2):
            opt.zero_grad()
            out = model(x)
            fnm_b200.parallel.rel_l2_sum(out, y).backward()
            opt.step()
        prof = _lib.profile_end()
        packs = sum(v[0] for k, v in prof.items() if k in ("mix_pack", "mix_unpack"))
        assert (packs == 0) == (order == "mode_major"), prof
        for p in model.parameters():
            if p.is_complex():
                assert p.grad.stride() == p.stride() and opt.state[p]["exp_avg"].stride() == p.stride()
        outs[order] = (out.detach().clone(), {k: v.detach().clone() for k, v in model.state_dict().items()})
    assert rel_l2(outs["mode_major"][0], outs["reference"][0]) < 1e-6
    for k, v in outs["mode_major"][1].items():
        assert v.shape == outs["reference"][1][k].shape
        assert rel_l2(v, outs["reference"][1][k]) < 1e-6, k


def test_edge_inputs_like_reference():
    """Grid smaller than ``padding`` (no padding is added and the reference's ``:-0`` crop yields an EMPTY output,
    func_to_func2d.py:101 -- kept), batch of one, and a 1-D model on the shortest grid the modes allow."""
    torch.manual_seed(2)
    model = fnm_b200.FNO2d(2, 2, 8, n_layers=2, width_final=16)
    p = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(DEV)
    x = torch.randn(1, 1, 6, 6)
    ref = O.f2f_forward(p, x)
    out = model(x.to(DEV))
    assert tuple(out.shape) == tuple(ref.shape) and out.numel() == 0          # (1, 1, 0, 0) like the reference
    x = torch.randn(1, 1, 9, 10)                                               # batch 1, padded to 10 x 11
    ref = O.f2f_forward(p, x)
    out = model(x.to(DEV))
    assert tuple(out.shape) == tuple(ref.shape) == (1, 1, 9, 10)
    assert rel_l2(out, ref) < TOL_FP32
    m1 = fnm_b200.FNO1d(4, 8, n_layers=2, width_final=16)
    p1 = {k: v.detach().clone() for k, v in m1.state_dict().items()}
    x1 = torch.randn(2, 1, 8)                                                  # padded to 9 points: modes 4 = floor(9 / 2)
    ref1 = O.f2f_forward(p1, x1)
    out1 = m1.to(DEV)(x1.to(DEV))
    assert rel_l2(out1, ref1) < TOL_FP32


# ------------------------------------------------------------ the path the benchmark times --
def _flat_step_eager(model, opt, grads, x, y):
    grads.zero_()
    out = model(x)
    loss = fnm_b200.parallel.rel_l2_sum(out, y)
    loss.backward()
    grads.all_reduce()
    opt.step()
    return loss, out


def _as_real(v):
    return torch.view_as_real(v) if v.is_complex() else v


@pytest.mark.parametrize("name", ["fno2d", "fnf2d", "fnd2d"])
def test_graphed_train_step_is_bit_equal_to_eager_and_matches_golden(name):
    """bench.py's headline path: GraphedTrainStep + FlatGradients(direct_spectral=True, overlap=True) +
    Adam(capturable=True).  Three graph replays must leave loss, gradients and parameters BIT-IDENTICAL to three eager
    steps of the same configuration, match the reference fixture to 1e-5, leave the model untouched by the warm-up
    steps of the capture, and keep the optimizer's step count right (checkpoint + eager continuation)."""
    from fnm_b200.parallel import FlatGradients, GraphedTrainStep
    d = load_npz(f"model_{name}.npz")
    x, y = t(d["x"]).to(DEV), t(d["y"]).to(DEV)
    runs = {}
    for how in ("eager", "graph"):
        model = MODELS[name]()
        model.load_state_dict({k: torch.from_numpy(v) for k, v in group(d, "w0").items()}, strict=True)
        model = model.to(DEV)
        opt = fnm_b200.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)
        grads = FlatGradients(model.parameters(), direct_spectral=True, overlap=True)
        losses = []
        if how == "graph":
            before = {k: v.detach().clone() for k, v in model.state_dict().items()}
            step = GraphedTrainStep(model, opt, grads, fnm_b200.parallel.rel_l2_sum, x, y, warmup=3)
            for k, v in model.state_dict().items():               # the capture's warm-up steps were undone
                assert torch.equal(_as_real(v), _as_real(before[k])), k
            # the first replay takes its inputs directly; the next two through the input pipeline of bench.py's end-to-end
            # loop: staged from pinned host memory on a side stream while the previous step runs, consumed by step()
            xh, yh = x.cpu().pin_memory(), y.cpu().pin_memory()
            losses.append(step(x, y).clone())
            for _ in range(2):
                step.stage(xh, yh)
                losses.append(step().clone())
            with pytest.raises(RuntimeError):
                step()                                                # nothing staged
            g = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
            assert step.launches_per_step > 0
            sd = opt.state_dict()
            assert {int(s["step"]) for s in sd["state"].values()} == {3}     # replays are counted
            # eager continuation after the replays (device counter and host count agree: bias correction of step 4)
            loss4, _ = _flat_step_eager(model, opt, grads, x, y)
            step.close()
        else:
            for _ in range(3):
                loss, out = _flat_step_eager(model, opt, grads, x, y)
                losses.append(loss.detach().clone())
            g = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
            loss4, _ = _flat_step_eager(model, opt, grads, x, y)
        torch.cuda.synchronize()
        runs[how] = (losses, g, {k: v.detach().clone() for k, v in model.state_dict().items()}, loss4.detach().clone())
    (le, ge, we, l4e), (lg, gg, wg, l4g) = runs["eager"], runs["graph"]
    for a, b in zip(le, lg):
        assert torch.equal(a, b)
    for k in ge:
        assert torch.equal(_as_real(ge[k]), _as_real(gg[k])), k
    for k in we:
        assert torch.equal(_as_real(we[k]), _as_real(wg[k])), k
    assert torch.equal(l4e, l4g)
    # ... and the fixture of the unmodified reference: loss of the third step
    assert rel_l2(lg[2], d["loss_last"]) < TOL_FP32


def test_adam_resumes_from_reference_format_checkpoint():
    """An optimizer state_dict saved by the reference's util.Adam (or any contiguous save) has CONTIGUOUS moments; the
    drop-in spectral weights are stored mode-major.  Loading must re-lay the moments out and continue bit-identically to
    an uninterrupted run; step counts survive the round trip."""
    torch.manual_seed(4)
    x = torch.randn(3, 1, 24, 24, device=DEV)
    y = torch.randn(3, 1, 24, 24, device=DEV)

    def fresh():
        torch.manual_seed(9)
        m = fnm_b200.FNO2d(4, 4, 32, n_layers=3, width_final=32).to(DEV)
        return m, fnm_b200.Adam(m.parameters(), lr=1e-3, weight_decay=1e-4)

    def run(m, o, n):
        for _ in range(n):
            o.zero_grad()
            fnm_b200.parallel.rel_l2_sum(m(x), y).backward()
            o.step()

    m_ref, o_ref = fresh()
    run(m_ref, o_ref, 4)
    m1, o1 = fresh()
    run(m1, o1, 2)
    sd_model = {k: v.detach().clone().contiguous() for k, v in m1.state_dict().items()}
    sd_opt = o1.state_dict()
    for st in sd_opt["state"].values():                          # what torch.save of the reference's optimizer holds
        for k, v in list(st.items()):
            if isinstance(v, torch.Tensor):
                st[k] = v.detach().clone().contiguous()
                assert st[k].is_contiguous()
    m2, o2 = fresh()
    m2.load_state_dict(sd_model)
    o2.load_state_dict(sd_opt)
    for p in m2.parameters():
        if p.is_complex():
            assert o2.state[p]["exp_avg"].stride() == p.stride() and not p.is_contiguous()
        assert o2.state[p]["step"] == 2
    run(m2, o2, 2)
    for (k, a), b in zip(m_ref.state_dict().items(), m2.state_dict().values()):
        assert torch.equal(_as_real(a), _as_real(b)), k


# ------------------------------------------------------------ BASELINE architectures 2, 3, 4 at model level --
def _train_step_vs_oracle(model, fwd, x, y, tol=TOL_FP32):
    p = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    ref = fwd(p, x)
    loss_ref = O.rel_l2_sum(ref, y)
    loss_ref.backward()
    model = model.to(DEV)
    out = model(x.to(DEV))
    loss = fnm_b200.parallel.rel_l2_sum(out, y.to(DEV))
    loss.backward()
    assert tuple(out.shape) == tuple(ref.shape)
    assert rel_l2(out, ref) < tol
    assert rel_l2(loss, loss_ref) < tol
    late = [k for k, q in model.named_parameters()
            if rel_l2(q.grad if q.grad is not None else torch.zeros_like(q), p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])) >= tol]
    if late:                                                     # ill-conditioned in fp32: judge against float64 (helpers.assert_parity)
        p64 = {k: to64(v).requires_grad_(True) for k, v in p.items()}
        O.rel_l2_sum(fwd(p64, x.double()), y.double()).backward()
        grads = dict(model.named_parameters())
        for k in late:
            assert_parity(grads[k].grad, p[k].grad, p64[k].grad, tol=tol, what=k)


def test_config2_architecture_train_step_vs_oracle():
    """BASELINE config 2: FNO1d(modes 16, width 64) on the 1024-point grid (padded 1152), batch cut to 4."""
    torch.manual_seed(0)
    model = fnm_b200.FNO1d(16, 64)
    x = torch.randn(4, 1, 1024, generator=torch.Generator().manual_seed(1234))
    y = torch.randn(4, 1, 1024, generator=torch.Generator().manual_seed(1235))
    _train_step_vs_oracle(model, lambda p, xx: O.f2f_forward(p, xx), x, y)


def test_config3_architecture_train_step_vs_oracle():
    """BASELINE config 3: FNF2d(12, 12, 32, d_in=2, d_out=2) on the 221 x 51 airfoil grid (padded 248 x 57: an odd row
    length, LY = 26 table columns in the F2V end -> the zero-padded table route of the tcgen05 kernel), batch cut to 3."""
    torch.manual_seed(0)
    model = fnm_b200.FNF2d(12, 12, 32, d_in=2, d_out=2)
    x = torch.randn(3, 2, 221, 51, generator=torch.Generator().manual_seed(1234))
    y = torch.randn(3, 2, generator=torch.Generator().manual_seed(1235))
    _train_step_vs_oracle(model, lambda p, xx: O.f2v_forward(p, xx), x, y)


def test_config4_architecture_train_step_vs_oracle():
    """BASELINE config 4: FND2d(20 -> 128 x 128, modes 16, width 64) (padded 144 x 144, V2F start with LY = 34), batch cut to 3."""
    torch.manual_seed(0)
    model = fnm_b200.FND2d(20, (128, 128), modes1=16, modes2=16, width=64)
    x = torch.randn(3, 20, generator=torch.Generator().manual_seed(1234))
    y = torch.randn(3, 1, 128, 128, generator=torch.Generator().manual_seed(1235))
    _train_step_vs_oracle(model, lambda p, xx: O.v2f_forward(p, xx, (128, 128)), x, y)


def test_config5_bench_batch_is_consistent_with_the_verified_batch():
    """The benchmark runs config 5 at batch 32 per GPU; the oracle comparison above runs batch 2 (a different number of
    weight-gradient partials and another persistent-tile schedule).  Samples are independent and the loss sums over the
    batch, so at batch 32: (a) the outputs of samples (0, 1) and (30, 31) equal the CPU oracle's outputs for those
    samples, and (b) every gradient equals the sum of the sixteen batch-2 gradients of the verified path."""
    torch.manual_seed(0)
    model = fnm_b200.FNO2d(32, 32, 128)
    p = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x = torch.randn(32, 1, 256, 256, generator=torch.Generator().manual_seed(1234))
    y = torch.randn(32, 1, 256, 256, generator=torch.Generator().manual_seed(1235))
    with torch.no_grad():
        ref_head, ref_tail = O.f2f_forward(p, x[:2]), O.f2f_forward(p, x[30:])
    model = model.to(DEV)
    xd, yd = x.to(DEV), y.to(DEV)
    out = model(xd)
    assert rel_l2(out[:2], ref_head) < TOL_FP32 and rel_l2(out[30:], ref_tail) < TOL_FP32
    loss = fnm_b200.parallel.rel_l2_sum(out, yd)
    loss.backward()
    big = {k: q.grad.detach().clone() for k, q in model.named_parameters()}
    loss_big = loss.detach().clone()
    del out, loss
    acc = {k: torch.zeros_like(v, dtype=torch.complex128 if v.is_complex() else torch.float64) for k, v in big.items()}
    loss_sum = 0.0
    for i in range(0, 32, 2):
        model.zero_grad(set_to_none=True)
        li = fnm_b200.parallel.rel_l2_sum(model(xd[i:i + 2]), yd[i:i + 2])
        li.backward()
        loss_sum += float(li)
        for k, q in model.named_parameters():
            acc[k] += q.grad.to(acc[k].dtype)
    assert abs(float(loss_big) - loss_sum) <= 1e-5 * abs(loss_sum)
    for k in big:
        assert rel_l2(big[k], acc[k]) < TOL_FP32, k


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_two_gpu_nccl_data_parallel_matches_single_rank_global_batch():
    """SURVEY section 4 'data parallel (b)': two ranks over NCCL (torch.distributed.run), batch sharded, per-slot
    overlapped SUM all-reduce issued from inside backward, eager AND replayed from a CUDA graph, against the single-rank
    global-batch run.  The worker (tests/dp_worker.py) asserts and exits non-zero on a mismatch."""
    import os
    import subprocess
    import sys
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dp_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", worker]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DP_PARITY_OK" in r.stdout


def test_overlapping_corner_blocks_like_reference():
    """modes1 <= P1 < 2 modes1: the reference assigns the weights1 block, then the weights2 block OVER the shared rows
    (shared.py:330-333).  Forward, input gradient and both weight gradients (zero on the dead weights1 rows)."""
    torch.manual_seed(6)
    sc = S.SpectralConv2d(4, 4, 5, 3)
    with torch.no_grad():
        sc.weights1.mul_(16)
        sc.weights2.mul_(16)
    x = torch.randn(2, 4, 8, 10)
    g = torch.randn(2, 4, 8, 10)
    xo = x.clone().requires_grad_(True)
    w1, w2 = sc.weights1.detach().clone().requires_grad_(True), sc.weights2.detach().clone().requires_grad_(True)
    O.spectral_conv2d(xo, w1, w2).backward(g)
    sc = sc.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    y = sc(xg)
    y.backward(g.to(DEV))
    assert rel_l2(y, O.spectral_conv2d(x, w1.detach(), w2.detach())) < TOL_FP32
    assert rel_l2(xg.grad, xo.grad) < TOL_FP32
    assert rel_l2(sc.weights1.grad, w1.grad) < TOL_FP32 and rel_l2(sc.weights2.grad, w2.grad) < TOL_FP32
    assert torch.all(sc.weights1.grad[:, :, 3:, :] == 0)           # rows 8 - 5 = 3 .. 4 of weights1 are overwritten


@pytest.mark.parametrize("family", ["fno2d", "fno1d", "fnd2d", "fnf2d", "fno2d1"])
def test_cached_gelu_derivative_equals_the_recomputed_one(family):
    """The trunk's default (a fused layer writes GELU'(z) in place of z for the next fused layer,
    FNM_PLAN_ZOUT_IS_GELU_GRAD) against the same model with the cache switched off: same formula, evaluated in the
    producer's epilogue instead of the consumer's backward."""
    from fnm_b200.models import _trunk
    torch.manual_seed(3)
    if family == "fno2d":
        model = fnm_b200.FNO2d(6, 6, 32, n_layers=4).to(DEV)
        x = torch.randn(3, 1, 40, 40, device=DEV)
    elif family == "fno1d":
        model = fnm_b200.FNO1d(8, 64, n_layers=3).to(DEV)
        x = torch.randn(3, 1, 128, device=DEV)
    elif family == "fnf2d":                                   # the F2V end consumes GELU(z) | GELU'(z) of the last layer too
        model = fnm_b200.FNF2d(6, 6, 32, d_in=1, d_out=2, n_layers=3).to(DEV)
        x = torch.randn(3, 1, 40, 32, device=DEV)
    elif family == "fno2d1":
        model = fnm_b200.FNO2d1(4, 16, 4, 4, 16, n_layers=3).to(DEV)
        x = torch.randn(2, 1, 24, 24, device=DEV)
    else:
        model = fnm_b200.FND2d(5, (24, 24), 6, 6, 32, n_layers=4).to(DEV)
        x = torch.randn(3, 5, device=DEV)
    res = {}
    for on in (True, False):
        old = _trunk._CACHE_GELU_GRAD
        _trunk._CACHE_GELU_GRAD = on
        try:
            model.zero_grad(set_to_none=True)
            out = model(x)
            out.square().sum().backward()
            res[on] = (out.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()})
        finally:
            _trunk._CACHE_GELU_GRAD = old
    assert torch.equal(res[True][0], res[False][0])
    for k in res[True][1]:
        assert rel_l2(res[True][1][k], res[False][1][k]) < 2e-6, k


@pytest.mark.parametrize("width", [4, 16, 64])
@pytest.mark.parametrize("end", ["f2v", "v2f"])
def test_ends_mode_major_weights_match_reference_ordered_weights(end, width):
    """F2V / V2F weights live mode-major in the drop-in modules (tcgen05 mixing kernels for the mode contractions, SIMT
    for narrow widths); the C ABI also takes them in the reference's memory order (fp32 SIMT contractions).  Same values,
    same gradients -- and both against the float64 statement of SURVEY A.3 through the oracle."""
    from fnm_b200 import ops
    torch.manual_seed(11)
    B, P1, P2, m1, m2 = 6, 20, 24, 4, 5
    if end == "f2v":
        mod = S.LinearFunctionals2d(width, width, m1, m2).to(DEV)
        x = torch.randn(B, width, P1, P2, device=DEV, requires_grad=True)
        run_mod = lambda: mod(x)
        tables = mod.tables(P1, P2)
        xcl = x.detach().permute(0, 2, 3, 1).contiguous().requires_grad_(True)
        wref = mod.weights.detach().contiguous().clone().requires_grad_(True)
        run_ref = lambda: ops.linear_functional(xcl, wref, tables)
        ref = O.linear_functionals2d(x.detach().cpu(), mod.weights.detach().cpu().contiguous())
    else:
        mod = S.LinearDecoder2d(width, width, m1, m2).to(DEV)
        x = torch.randn(B, width, device=DEV, requires_grad=True)
        run_mod = lambda: mod(x, (P1, P2))
        tables = mod.tables((P1, P2))
        xcl = x.detach().clone().requires_grad_(True)
        wref = mod.weights.detach().contiguous().clone().requires_grad_(True)
        run_ref = lambda: ops.linear_decoder(xcl, wref, tables).permute(0, 3, 1, 2)
        ref = O.linear_decoder2d(x.detach().cpu(), mod.weights.detach().cpu().contiguous(), (P1, P2))
    assert not mod.weights.is_contiguous() and wref.is_contiguous()
    out_m = run_mod()
    gsel = torch.randn_like(out_m)
    (out_m * gsel).sum().backward()
    out_r = run_ref()
    (out_r * gsel).sum().backward()
    assert rel_l2(out_m, ref) < TOL_FP32 and rel_l2(out_r, ref) < TOL_FP32
    assert rel_l2(out_m, out_r) < TOL_FP32
    assert rel_l2(mod.weights.grad, wref.grad) < TOL_FP32
    gx_r = xcl.grad.permute(0, 3, 1, 2) if end == "f2v" else xcl.grad
    assert rel_l2(x.grad, gx_r) < TOL_FP32


@pytest.mark.parametrize("family", ["fnf2d", "fnf1d", "fnn2d"])
def test_one_pass_local_branch_matches_the_two_kernel_form(family):
    """The F2V end's local branch (mlpfunc0 + trapz) in one kernel against fc1 kernel + weighted-sum kernel."""
    from fnm_b200.models import _trunk
    torch.manual_seed(5)
    if family == "fnf2d":
        model = fnm_b200.FNF2d(6, 6, 32, d_in=1, d_out=2, n_layers=3).to(DEV)
        x = torch.randn(3, 1, 40, 32, device=DEV)
    elif family == "fnf1d":
        model = fnm_b200.FNF1d(8, 64, d_in=1, d_out=2, n_layers=3).to(DEV)
        x = torch.randn(3, 1, 256, device=DEV)
    else:
        model = fnm_b200.FNN2d(5, 3, s_latentspace=(24, 24), modes1=6, modes2=6, width=32, n_layers=4).to(DEV)
        x = torch.randn(3, 5, device=DEV)
    res = {}
    for on in (True, False):
        old = _trunk._ONE_PASS_LOCAL
        _trunk._ONE_PASS_LOCAL = on
        try:
            model.zero_grad(set_to_none=True)
            out = model(x)
            out.square().sum().backward()
            res[on] = (out.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()})
        finally:
            _trunk._ONE_PASS_LOCAL = old
    assert rel_l2(res[True][0], res[False][0]) < 2e-6
    for k in res[True][1]:
        assert rel_l2(res[True][1][k], res[False][1][k]) < 5e-6, k
