"""CPU tests of the host side: mode-plan tables (checked by emulating the kernel stage chain in
numpy against the float64 oracle), the C-ABI library (loads, exports every declared symbol),
state_dict layout parity with the reference, error behaviour, and the loud no-fallback rule."""
import json
import os
import re

import numpy as np
import pytest
import torch

import fnm_b200
from fnm_b200 import plan as P
from oracle import dense_f64 as D
from tests.helpers import GOLDEN, ROOT, rel_l2


# ---------------------------------------------------------------- stage-chain emulation -----
def _analysis(tab, xcl):
    """xcl (B,P1,P2,C) -> complex spectrum (B,R,NY,C) using the ay / ex tables exactly as the kernels do."""
    h = {k: v.astype(np.float64) for k, v in tab.host.items()}
    B, P1, P2, C = xcl.shape
    R, NY = tab.R, tab.NY
    T = np.einsum("ly,bxyc->bxlc", h["ay"], xcl)                       # [b][x][(ri,l)][c]
    Tm = T.reshape(B, P1 * 2, NY * C)                                   # k = x*2 + ri
    Xm = np.einsum("mk,bkn->bmn", h["ex"], Tm).reshape(B, 2, R, NY, C)  # m = ro*R + r
    return Xm[:, 0] + 1j * Xm[:, 1]


def _synthesis(tab, oh):
    """complex (B,R,NY,Co) -> real (B,S1,S2,Co) using fx / by."""
    h = {k: v.astype(np.float64) for k, v in tab.host.items()}
    B, R, NY, Co = oh.shape
    Om = np.stack([oh.real, oh.imag], axis=1).reshape(B, 2 * R, NY * Co)   # k = ri*R + r
    Z = np.einsum("mk,bkn->bmn", h["fx"], Om)                               # m = x*2 + ro
    Zl = Z.reshape(B, tab.S1, 2 * NY, Co)                                   # l' = ro*NY + l
    return np.einsum("yl,bxlo->bxyo", h["by"], Zl)


def _analysis_of_grad(tab, gcl):
    h = {k: v.astype(np.float64) for k, v in tab.host.items()}
    B, S1, S2, C = gcl.shape
    R, NY = tab.R, tab.NY
    T = np.einsum("ly,bxyc->bxlc", h["byT"], gcl)
    Xm = np.einsum("mk,bkn->bmn", h["exb"], T.reshape(B, S1 * 2, NY * C)).reshape(B, 2, R, NY, C)
    return Xm[:, 0] + 1j * Xm[:, 1]


def _synthesis_of_grad(tab, gxh):
    h = {k: v.astype(np.float64) for k, v in tab.host.items()}
    B, R, NY, C = gxh.shape
    Om = np.stack([gxh.real, gxh.imag], axis=1).reshape(B, 2 * R, NY * C)
    Z = np.einsum("mk,bkn->bmn", h["fxb"], Om).reshape(B, tab.P1, 2 * NY, C)
    return np.einsum("yl,bxlo->bxyo", h["ayT"], Z)


def _cl(x):
    return np.moveaxis(x, 1, -1)


def _cf(x):
    return np.moveaxis(x, -1, 1)


RNG = np.random.default_rng(0)


def _cw(*shape):
    return RNG.standard_normal(shape) + 1j * RNG.standard_normal(shape)


@pytest.mark.parametrize("P1,P2,m1,m2,s", [(9, 9, 2, 5, None), (8, 8, 4, 5, None), (12, 10, 3, 4, None),
                                           (16, 12, 3, 3, (8, 10)), (8, 10, 2, 3, (12, 15)), (12, 12, 4, 5, (6, 6)),
                                           (9, 11, 3, 4, (13, 7))])
def test_spectral_tables_2d(P1, P2, m1, m2, s):
    S1, S2 = (P1, P2) if s is None else s
    tab = P.spectral_tables(P1, P2, m1, m2, S1, S2)
    x = RNG.standard_normal((2, 3, P1, P2))
    w1, w2 = _cw(3, 4, m1, m2), _cw(3, 4, m1, m2)
    w = np.concatenate([w1, w2], axis=2)
    xh = _analysis(tab, _cl(x))
    oh = np.einsum("brli,iorl->brlo", xh, w)
    y = _cf(_synthesis(tab, oh))
    assert rel_l2(y, D.spectral_conv2d(x, w1, w2, s)) < 2e-7
    g = RNG.standard_normal(y.shape)
    gh = _analysis_of_grad(tab, _cl(g))
    dw = np.einsum("brli,brlo->iorl", np.conj(xh), gh)
    dx = _cf(_synthesis_of_grad(tab, np.einsum("brlo,iorl->brli", gh, np.conj(w))))
    rdx, rdw1, rdw2 = D.spectral_conv2d_backward(x, w1, w2, g, s)
    assert rel_l2(dx, rdx) < 2e-7
    assert rel_l2(dw[:, :, :m1], rdw1) < 2e-7 and rel_l2(dw[:, :, m1:], rdw2) < 2e-7


@pytest.mark.parametrize("Pn,m,s", [(16, 5, None), (9, 4, None), (8, 5, None), (16, 4, 10), (9, 3, 14)])
def test_spectral_tables_1d(Pn, m, s):
    tab = P.spectral_tables(1, Pn, 0, m, 1, Pn if s is None else s, one_d=True)
    x = RNG.standard_normal((2, 3, Pn))
    w1 = _cw(3, 4, m)
    xh = _analysis(tab, _cl(x[:, :, None, :]))
    y = _cf(_synthesis(tab, np.einsum("brli,iol->brlo", xh, w1)))[:, :, 0]
    assert rel_l2(y, D.spectral_conv1d(x, w1, s)) < 2e-7


@pytest.mark.parametrize("P1,P2,m1,m2", [(12, 10, 3, 3), (9, 7, 2, 2), (6, 6, 4, 4)])
def test_functional_tables(P1, P2, m1, m2):
    tab = P.functional_tables(P1, P2, m1, m2)
    x = RNG.standard_normal((2, 3, P1, P2))
    w = _cw(3, 4, 2 * m1, m2 + 1)
    xh = _analysis(tab, _cl(x))
    out = np.einsum("rl,brli,iorl->bo", tab.host["cc"].astype(np.float64), xh, w).real
    assert rel_l2(out, D.linear_functionals2d(x, w)) < 2e-7
    tab1 = P.functional_tables(1, P2, 0, m2, one_d=True)
    x1, w1 = RNG.standard_normal((2, 3, P2)), _cw(3, 4, m2 + 1)
    xh1 = _analysis(tab1, _cl(x1[:, :, None, :]))
    out1 = np.einsum("rl,brli,iol->bo", tab1.host["cc"].astype(np.float64), xh1, w1).real
    assert rel_l2(out1, D.linear_functionals1d(x1, w1)) < 2e-7


@pytest.mark.parametrize("m1,m2,s", [(3, 3, (12, 10)), (2, 2, (9, 7)), (4, 4, (6, 5)), (4, 3, (5, 8))])
def test_decoder_tables(m1, m2, s):
    tab = P.decoder_tables(s[0], s[1], m1, m2)
    v, w = RNG.standard_normal((2, 3)), _cw(3, 4, 2 * m1, m2 + 1)
    oh = np.einsum("bi,iorl->brlo", v, w)
    assert rel_l2(_cf(_synthesis(tab, oh)), D.linear_decoder2d(v, w, s)) < 2e-7
    tab1 = P.decoder_tables(1, s[1], 0, m2, one_d=True)
    w1 = _cw(3, 4, m2 + 1)
    oh1 = np.einsum("bi,iol->blo", v, w1)[:, None]
    assert rel_l2(_cf(_synthesis(tab1, oh1))[:, :, 0], D.linear_decoder1d(v, w1, s[1])) < 2e-7


def test_index_maps_match_oracle_bookkeeping():
    for n in range(1, 14):
        for s in range(1, 16):
            assert np.array_equal(P.onesided_source(n, s), D.half_map(n, s))
            assert np.array_equal(P.twosided_source(n, s), D.full_map(n, s))


def test_tables_snap_exact_zeros():
    tab = P.spectral_tables(8, 8, 4, 5)
    ay = tab.host["ay"].reshape(2, 5, 8)
    assert np.all(ay[1, 0] == 0) and np.all(ay[1, 4] == 0)          # Im of DC and Nyquist rows is exactly 0
    by = tab.host["byT"].reshape(2, 5, 8)
    assert np.allclose(by[0, 0], 1.0 / 64) and np.allclose(np.abs(by[0, 4]), 1.0 / 64)   # c_l = 1 at DC/Nyquist
    assert np.allclose(np.abs(by[0, 1]).max(), 2.0 / 64)


# ---------------------------------------------------------------- errors (reference behaviour) -
def test_value_errors_like_reference():
    with pytest.raises(ValueError):
        fnm_b200.shared._get_act("swish")                           # shared.py:22
    with pytest.raises(ValueError):
        P.onesided_source(5, 0)                                     # shared.py:78
    with pytest.raises(ValueError):
        P.twosided_source(5, 0)                                     # shared.py:104
    with pytest.raises(ValueError):
        fnm_b200.FNN2d(4, 2, n_layers=1)                            # vec_to_vec2d.py:60-61
    with pytest.raises(ValueError):
        fnm_b200.Adam([torch.nn.Parameter(torch.zeros(1))], lr=-1.0)   # Adam.py:81-82
    with pytest.raises(ValueError):
        P.spectral_tables(3, 8, 4, 3)                               # modes1 > grid: the reference's slice assignment fails too
    # overlapping corner blocks (modes1 <= P1 < 2 modes1) RUN in the reference: the weights2 block is assigned last
    # and wins on the shared rows (shared.py:330-333) -> those weights1 rows are dead in the tables
    tb = P.spectral_tables(6, 8, 4, 3)
    assert list(tb.rows_in) == [0, 1, -1, -1, 2, 3, 4, 5]


def test_resize_helpers_match_oracle():
    from oracle import fnm_oracle as O
    a = torch.randn(2, 3, 7, 6, dtype=torch.cfloat)
    for s in [(4, 4), (9, 14), (5, 3), (7, 10), (1, 1)]:
        assert torch.equal(fnm_b200.shared.resize_rfft2(a, s), O.keep_spectrum2(a, s))


@pytest.mark.parametrize("P1,P2,S1,S2", [(12, 10, 8, 6), (12, 10, 16, 14), (9, 7, 14, 11), (9, 9, 5, 5), (145, 145, 37, 37),
                                         (8, 9, 16, 18), (7, 8, 3, 1), (16, 16, 10, 16), (6, 1, 4, 3)])
def test_resample_tables_2d(P1, P2, S1, S2):
    """projector2d as table contractions (one real table per axis, or the two-term form where the two-sided
    resize keeps a non-Hermitian bin set) against the FFT statement, including the golden vectors."""
    tab = P.resample_tables(P1, P2, S1, S2)
    x = torch.randn(2, 3, P1, P2, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    assert rel_l2(tab.apply_host(x), fnm_b200.shared.projector2d(x, (S1, S2))) < 5e-7
    if (P1, P2) == (12, 10):
        g = np.load(os.path.join(GOLDEN, "blocks.npz"))
        xg = torch.from_numpy(g["proj2.x"]).double()
        assert rel_l2(tab.apply_host(xg), g["proj2.down" if S1 == 8 else "proj2.up"]) < 5e-7
    if P1 % 2 == 1 and S1 % 2 == 1 and S1 <= P1:
        assert tab.separable            # odd -> odd truncation keeps a symmetric bin set
    if S1 % 2 == 0 and 1 < S1 < P1:
        assert not tab.separable        # an even target keeps its Nyquist bin on the negative side only


def test_resample_tables_vs_reference_projector_fixtures():
    """projectors.npz holds outputs of the reference's projector1d / projector2d at odd and mixed sizes (the
    N // 2 split of resize_fft, one-sided Nyquist bins)."""
    g = np.load(os.path.join(GOLDEN, "projectors.npz"))
    for tag in sorted({k.split(".")[0] for k in g.files}):
        x, y, s = torch.from_numpy(g[f"{tag}.x"]).double(), g[f"{tag}.y"], g[f"{tag}.s"]
        if tag.startswith("p2"):
            tab = P.resample_tables(x.shape[-2], x.shape[-1], int(s[0]), int(s[1]))
        else:
            tab = P.resample_tables(None, x.shape[-1], None, int(s[0]))
        assert rel_l2(tab.apply_host(x), y) < 5e-7, tag


@pytest.mark.parametrize("Pn,s", [(12, 7), (12, 20), (9, 14), (16, 16), (5, 1)])
def test_resample_tables_1d(Pn, s):
    tab = P.resample_tables(None, Pn, None, s)
    x = torch.randn(2, 3, Pn, dtype=torch.float64, generator=torch.Generator().manual_seed(4))
    assert rel_l2(tab.apply_host(x), fnm_b200.shared.projector1d(x, s)) < 5e-7
    if Pn == 12:
        g = np.load(os.path.join(GOLDEN, "blocks.npz"))
        assert rel_l2(tab.apply_host(torch.from_numpy(g["proj1.x"]).double()), g["proj1.down" if s == 7 else "proj1.up"]) < 5e-7


# ---------------------------------------------------------------- no CPU fallback --------------
def test_cpu_tensors_fail_loudly():
    m = fnm_b200.FNO2d(3, 3, 8, n_layers=2, width_final=16)
    with pytest.raises(fnm_b200.FnmError):
        m(torch.randn(1, 1, 16, 16))
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(fnm_b200.FnmError):
        fnm_b200.Adam([p]).step()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fourier-neural-mappings_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


# ---------------------------------------------------------------- C ABI ------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "fnm_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(fnm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(fnm_b200._lib.PROTOTYPES), declared ^ set(fnm_b200._lib.PROTOTYPES)
    handle = fnm_b200._lib.lib()                                     # loads the .so; no compute call
    for name in declared:
        assert hasattr(handle, name), name
    assert b"fnm_b200" in handle.fnm_version()
    # the ctypes mirror of struct fnm_plan has the size the library was compiled with (checked at load time too), and
    # its field names are the header's, in order
    import ctypes
    assert handle.fnm_plan_sizeof() == ctypes.sizeof(fnm_b200._lib.FnmPlan)
    body = re.search(r"typedef struct fnm_plan \{(.*?)\} fnm_plan;", header, flags=re.S).group(1)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            names += [n.strip(" *") for n in re.sub(r"^(const\s+)?\w+\s*\*?", "", decl, count=1).split(",")]
    assert names == [n for n, _ in fnm_b200._lib.FnmPlan._fields_], names


def test_layer_workspace_holds_the_partial_transforms_of_the_segmented_1d_ydft():
    """1-D models have few long rows: the TMA-fed y-DFT splits each row into K segments and the x-transform adds the partial
    transforms up, so the layer workspace of a 1-D plan holds that many copies of T (sized from the device's SM count --
    148 when no GPU is present -- and independent of fnm_set_sm_limit); many-row 1-D plans get no copies."""
    import ctypes
    from fnm_b200.plan import spectral_tables
    L = fnm_b200._lib.lib()

    def ws(tables, B, C):
        plan = tables.c_plan(torch.device("cpu"), B, C, C)
        return L.fnm_fourier_layer_workspace_bytes(ctypes.byref(plan))

    t1 = spectral_tables(1, 1152, 1, 16, one_d=True)
    few, many = ws(t1, 64, 64), ws(t1, 640, 64)                     # 32 row groups -> 4 segments; 320 groups -> none
    per_sample_T = 4 * 32 * 64                                      # bytes of one sample's T: LY = 32 rows of 64 channels
    assert few - 64 * per_sample_T * 3 > 0
    assert few / 64 > many / 640 + 2.5 * per_sample_T               # three extra copies of T per sample
    prev = L.fnm_set_sm_limit(100)
    try:
        assert ws(t1, 64, 64) == few                                # workspace sizes do not follow the SM budget
    finally:
        L.fnm_set_sm_limit(prev)


def test_c_abi_rejects_bad_arguments_without_gpu():
    L = fnm_b200._lib.lib()
    assert L.fnm_gelu_fwd(None, None, 4, None) == 1                  # FNM_EINVAL
    assert b"NULL" in L.fnm_last_error()
    assert L.fnm_fourier_layer_workspace_bytes(None) == 0


# ---------------------------------------------------------------- state_dict contract ----------
def test_state_dict_layout_matches_reference():
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_layouts.json")))
    ref.update(json.load(open(os.path.join(GOLDEN, "state_dict_layouts_cross_dim.json"))))
    ctors = {
        "FNO1d2": lambda: fnm_b200.FNO1d2(),
        "FNO2d1": lambda: fnm_b200.FNO2d1(),
        "FNO2d": lambda: fnm_b200.FNO2d(12, 12, 32),
        "FNO1d": lambda: fnm_b200.FNO1d(16, 64),
        "FNF2d": lambda: fnm_b200.FNF2d(12, 12, 32, d_in=2, d_out=2),
        "FNF1d": lambda: fnm_b200.FNF1d(16, 64),
        "FND2d": lambda: fnm_b200.FND2d(20, (128, 128), modes1=16, modes2=16, width=64),
        "FND2d_linear_first": lambda: fnm_b200.FND2d(20, (32, 32), nonlinear_first=False),
        "FND1d": lambda: fnm_b200.FND1d(20, 1024),
        "FNN2d": lambda: fnm_b200.FNN2d(20, 4),
        "FNN1d": lambda: fnm_b200.FNN1d(20, 4),
    }
    for name, ctor in ctors.items():
        got = {k: [list(v.shape), str(v.dtype)] for k, v in ctor().state_dict().items()}
        assert list(got.keys()) == list(ref[name].keys()), name      # same keys, same ORDER
        assert got == ref[name], name


def test_init_distribution_matches_reference():
    """scale * U[0,1) on both real and imaginary parts (shared.py:311-313)."""
    torch.manual_seed(0)
    sc = fnm_b200.SpectralConv2d(4, 8, 3, 3)
    w = torch.view_as_real(sc.weights1.detach())
    assert w.min() >= 0 and w.max() <= 1.0 / 32 and abs(w.mean().item() * 32 - 0.5) < 0.05
