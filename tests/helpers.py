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
