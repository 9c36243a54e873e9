"""Pins the oracle (oracle/fnm_oracle.py fp32 torch, oracle/dense_f64.py float64 numpy) against
fixtures produced by the unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import dense_f64 as D
from oracle import fnm_oracle as O
from tests.helpers import group, load_npz, rel_l2, t

BLOCKS = load_npz("blocks.npz")
# fp32 oracle uses the same library calls as the reference: agreement to rounding of a re-run.
TIGHT = 2e-6
# the float64 restatement vs the fp32 reference: the reference's own fp32 noise floor.
F64_VS_REF = 5e-6


def _s(g, n):
    s = g["s"]
    if s[0] < 0:
        return None
    return int(s[0]) if n == 1 else tuple(int(v) for v in s)


# ------------------------------------------------------------------ SpectralConv ----------
@pytest.mark.parametrize("tag", ["sc2_odd", "sc2_nyq", "sc2_rect", "sc2_down", "sc2_up", "sc2_cut"])
def test_spectral_conv2d(tag):
    g = group(BLOCKS, tag)
    s = _s(g, 2)
    x, w1, w2 = t(g["x"], grad=True), t(g["weights1"], grad=True), t(g["weights2"], grad=True)
    y = O.spectral_conv2d(x, w1, w2, s)
    y.backward(t(g["g"]))
    assert rel_l2(y, g["y"]) < TIGHT
    assert rel_l2(x.grad, g["dx"]) < TIGHT
    assert rel_l2(w1.grad, g["dweights1"]) < TIGHT
    assert rel_l2(w2.grad, g["dweights2"]) < TIGHT
    # float64 dense restatement, forward and hand-derived backward
    xd, w1d, w2d = g["x"].astype(np.float64), g["weights1"].astype(np.complex128), g["weights2"].astype(np.complex128)
    assert rel_l2(D.spectral_conv2d(xd, w1d, w2d, s), g["y"]) < F64_VS_REF
    dx, dw1, dw2 = D.spectral_conv2d_backward(xd, w1d, w2d, g["g"].astype(np.float64), s)
    assert rel_l2(dx, g["dx"]) < F64_VS_REF
    assert rel_l2(dw1, g["dweights1"]) < F64_VS_REF
    assert rel_l2(dw2, g["dweights2"]) < F64_VS_REF


@pytest.mark.parametrize("tag", ["sc1_even", "sc1_odd", "sc1_nyq", "sc1_down", "sc1_up"])
def test_spectral_conv1d(tag):
    g = group(BLOCKS, tag)
    s = _s(g, 1)
    x, w1 = t(g["x"], grad=True), t(g["weights1"], grad=True)
    y = O.spectral_conv1d(x, w1, s)
    y.backward(t(g["g"]))
    assert rel_l2(y, g["y"]) < TIGHT
    assert rel_l2(x.grad, g["dx"]) < TIGHT
    assert rel_l2(w1.grad, g["dweights1"]) < TIGHT
    xd, w1d = g["x"].astype(np.float64), g["weights1"].astype(np.complex128)
    assert rel_l2(D.spectral_conv1d(xd, w1d, s), g["y"]) < F64_VS_REF
    dx, dw = D.spectral_conv1d_backward(xd, w1d, g["g"].astype(np.float64), s)
    assert rel_l2(dx, g["dx"]) < F64_VS_REF
    assert rel_l2(dw, g["dweights1"]) < F64_VS_REF


# ------------------------------------------------------------------ F2V -------------------
@pytest.mark.parametrize("tag", ["lf2", "lf2_odd", "lf2_pad"])
def test_linear_functionals2d(tag):
    g = group(BLOCKS, tag)
    x, w = t(g["x"], grad=True), t(g["weights"], grad=True)
    y = O.linear_functionals2d(x, w)
    y.backward(t(g["g"]))
    assert rel_l2(y, g["y"]) < TIGHT
    assert rel_l2(x.grad, g["dx"]) < TIGHT
    assert rel_l2(w.grad, g["dweights"]) < TIGHT
    xd, wd = g["x"].astype(np.float64), g["weights"].astype(np.complex128)
    assert rel_l2(D.linear_functionals2d(xd, wd), g["y"]) < F64_VS_REF
    dx, dw = D.linear_functionals2d_backward(xd, wd, g["g"].astype(np.float64))
    assert rel_l2(dx, g["dx"]) < F64_VS_REF
    assert rel_l2(dw, g["dweights"]) < F64_VS_REF
    # dead weights (rows >= m1 of column 0) receive exactly zero gradient in the reference
    m1 = g["weights"].shape[-2] // 2
    assert np.all(g["dweights"][:, :, m1:, 0] == 0)
    assert np.all(dw[:, :, m1:, 0] == 0)


@pytest.mark.parametrize("tag", ["lf1", "lf1_odd", "lf1_pad"])
def test_linear_functionals1d(tag):
    g = group(BLOCKS, tag)
    x, w = t(g["x"], grad=True), t(g["weights"], grad=True)
    y = O.linear_functionals1d(x, w)
    y.backward(t(g["g"]))
    assert rel_l2(y, g["y"]) < TIGHT
    assert rel_l2(x.grad, g["dx"]) < TIGHT
    assert rel_l2(w.grad, g["dweights"]) < TIGHT
    xd, wd = g["x"].astype(np.float64), g["weights"].astype(np.complex128)
    assert rel_l2(D.linear_functionals1d(xd, wd), g["y"]) < F64_VS_REF
    dx, dw = D.linear_functionals1d_backward(xd, wd, g["g"].astype(np.float64))
    assert rel_l2(dx, g["dx"]) < F64_VS_REF
    assert rel_l2(dw, g["dweights"]) < F64_VS_REF


# ------------------------------------------------------------------ V2F -------------------
@pytest.mark.parametrize("tag", ["ld2", "ld2_odd", "ld2_cut", "ld2_cut_odd"])
def test_linear_decoder2d(tag):
    g = group(BLOCKS, tag)
    s = _s(g, 2)
    v, w = t(g["x"], grad=True), t(g["weights"], grad=True)
    y = O.linear_decoder2d(v, w, s)
    y.backward(t(g["g"]))
    assert rel_l2(y, g["y"]) < TIGHT
    assert rel_l2(v.grad, g["dx"]) < TIGHT
    assert rel_l2(w.grad, g["dweights"]) < TIGHT
    vd, wd = g["x"].astype(np.float64), g["weights"].astype(np.complex128)
    assert rel_l2(D.linear_decoder2d(vd, wd, s), g["y"]) < F64_VS_REF
    dv, dw = D.linear_decoder2d_backward(vd, wd, g["g"].astype(np.float64))
    assert rel_l2(dv, g["dx"]) < F64_VS_REF
    assert rel_l2(dw, g["dweights"]) < F64_VS_REF


@pytest.mark.parametrize("tag", ["ld1", "ld1_odd", "ld1_cut"])
def test_linear_decoder1d(tag):
    g = group(BLOCKS, tag)
    s = _s(g, 1)
    v, w = t(g["x"], grad=True), t(g["weights"], grad=True)
    y = O.linear_decoder1d(v, w, s)
    y.backward(t(g["g"]))
    assert rel_l2(y, g["y"]) < TIGHT
    assert rel_l2(v.grad, g["dx"]) < TIGHT
    assert rel_l2(w.grad, g["dweights"]) < TIGHT
    vd, wd = g["x"].astype(np.float64), g["weights"].astype(np.complex128)
    assert rel_l2(D.linear_decoder1d(vd, wd, s), g["y"]) < F64_VS_REF
    dv, dw = D.linear_decoder1d_backward(vd, wd, g["g"].astype(np.float64))
    assert rel_l2(dv, g["dx"]) < F64_VS_REF
    assert rel_l2(dw, g["dweights"]) < F64_VS_REF


def test_projectors():
    assert rel_l2(O.project2d(t(BLOCKS["proj2.x"]), (8, 6)), BLOCKS["proj2.down"]) < TIGHT
    assert rel_l2(O.project2d(t(BLOCKS["proj2.x"]), (16, 14)), BLOCKS["proj2.up"]) < TIGHT
    assert rel_l2(O.project1d(t(BLOCKS["proj1.x"]), 7), BLOCKS["proj1.down"]) < TIGHT
    assert rel_l2(O.project1d(t(BLOCKS["proj1.x"]), 20), BLOCKS["proj1.up"]) < TIGHT


def test_projectors_odd_and_mixed_sizes():
    g = load_npz("projectors.npz")
    for tag in sorted({k.split(".")[0] for k in g}):
        x, s = t(g[f"{tag}.x"]), g[f"{tag}.s"]
        got = O.project2d(x, tuple(int(v) for v in s)) if tag.startswith("p2") else O.project1d(x, int(s[0]))
        assert rel_l2(got, g[f"{tag}.y"]) < TIGHT, tag


def test_errors_match_reference():
    with pytest.raises(ValueError):
        O.activation("swish")                      # shared.py:22
    with pytest.raises(ValueError):
        O.keep_half_spectrum(torch.zeros(3, dtype=torch.cfloat), 0)   # shared.py:78
    with pytest.raises(ValueError):
        O.keep_full_spectrum(torch.zeros(3, dtype=torch.cfloat), 0)   # shared.py:104


# ------------------------------------------------------------------ models + Adam ---------
MODEL_FWD = {
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


@pytest.mark.parametrize("name", sorted(MODEL_FWD))
def test_model_train_steps(name):
    """forward, loss, every gradient, and every parameter after 3 reference-Adam steps."""
    torch.set_num_threads(4)
    d = load_npz(f"model_{name}.npz")
    p = {k: t(v, grad=True) for k, v in group(d, "w0").items()}
    x, y = t(d["x"]), t(d["y"])
    state = O.AdamState()
    fwd = MODEL_FWD[name]
    for step in range(3):
        loss, out = O.train_step(fwd, p, x, y, state, lr=1e-3, weight_decay=1e-4)
        if step == 0:
            assert rel_l2(out, d["out0"]) < TIGHT
            assert rel_l2(loss, d["loss0"]) < TIGHT
            for k, g in group(d, "g0").items():
                got = p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])
                assert rel_l2(got, g) < 5e-6, k
    assert rel_l2(loss, d["loss_last"]) < 1e-5
    for k, w in group(d, "w3").items():
        assert rel_l2(p[k], w) < 1e-5, k


@pytest.mark.parametrize("tag,kw", [("plain", dict(lr=1e-3)), ("wd", dict(lr=2e-3, weight_decay=1e-2)),
                                    ("ams", dict(lr=1e-3, weight_decay=1e-4, amsgrad=True))])
def test_adam(tag, kw):
    d = group(load_npz("adam.npz"), tag)
    use_c = tag != "ams"
    pc, pr = t(d["pc0"], grad=True), t(d["pr0"], grad=True)
    params = [pc, pr] if use_c else [pr]
    state = O.AdamState()
    for step in range(4):
        pc.grad = t(d[f"gc{step}"]) if use_c else None
        pr.grad = t(d[f"gr{step}"])
        O.adam_step(params, state, **kw)
        if use_c:
            assert rel_l2(pc, d[f"pc{step + 1}"]) < 1e-6
        assert rel_l2(pr, d[f"pr{step + 1}"]) < 1e-6
    if use_c:
        assert rel_l2(state.m[0], d["mc"]) < 1e-6
        assert rel_l2(state.v[0], d["vc"]) < 1e-6
        # second moment is |g|^2: real up to the rounding of the complex multiply (~1e-8 relative)
        assert state.v[0].imag.abs().max() < 1e-6 * state.v[0].real.abs().max()
        # float64 restatement of the update rule, one step from zero state
        p1, m1, v1, _ = D.adam_update(d["pc0"].astype(np.complex128), d["gc0"].astype(np.complex128),
                                      0.0, 0.0, 1, lr=kw["lr"], wd=kw.get("weight_decay", 0.0))
        assert rel_l2(p1, d["pc1"]) < 1e-6


def test_reference_amsgrad_rejects_complex():
    """util/Adam.py:43 calls torch.maximum on the complex second moment, which raises."""
    pc = t(np.ones((2, 2), dtype=np.complex64), grad=True)
    pc.grad = torch.ones_like(pc)
    with pytest.raises(RuntimeError):
        O.adam_step([pc], O.AdamState(), amsgrad=True)
