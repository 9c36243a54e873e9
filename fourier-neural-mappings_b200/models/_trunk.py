"""The layer loop every Fourier Neural Mapping shares, on channels-last activations.

Reference: the three-line loop ``x = w(x) + speconv(x); x = self.act(x)`` of
/root/reference/models/func_to_func2d.py:89-95 (and its twins in the other model files).  With
GELU the activation is never materialised between layers: each fused layer kernel stores its
pre-activation z and the NEXT layer applies GELU while loading ("save only
the pre-activation"); backward multiplies by GELU'(z) in the producing kernel's epilogue.
"""
import os

import torch
import torch.nn.functional as F

from .. import ops
from ..shared import projector1d, projector2d

# FNM_CACHE_GELU_GRAD=0 keeps z between fused layers (the derivative is then evaluated in the consumer's backward)
_CACHE_GELU_GRAD = os.environ.get("FNM_CACHE_GELU_GRAD", "1") != "0"

# FNM_ONE_PASS_LOCAL=0: F2V local branch as fc1 kernel + weighted-sum kernel (the hidden tensor is materialised)
_ONE_PASS_LOCAL = os.environ.get("FNM_ONE_PASS_LOCAL", "1") != "0"

_TORCH_ACTS = {"tanh": torch.tanh, "relu": F.relu, "elu": F.elu, "leaky_relu": F.leaky_relu}


class Pending:
    """A channels-last activation of the trunk.  ``z`` is the differentiable tensor; when ``act`` is
    set its GELU is still pending, and ``x`` may hold GELU(z) already (written by the producing
    kernel's epilogue, detached: gradients flow through ``z``)."""

    __slots__ = ("z", "act", "x", "zgrad")

    def __init__(self, z, act=False, x=None, zgrad=False):
        # zgrad: ``z`` holds GELU'(z) instead of z (ops.fourier_layer(grad_out=True)); only the next fused layer may read it
        self.z, self.act, self.x, self.zgrad = z, act, x, zgrad

    def value(self):
        """The activated tensor as a differentiable value (for consumers outside the fused layers)."""
        if self.zgrad:
            raise RuntimeError("this activation carries GELU'(z) for the next fused layer only")
        if not self.act:
            return self.z
        y = ops.gelu(self.z, self.x)                          # no forward kernel when the producer already wrote GELU(z)
        self.x = y.detach()
        return y


def _layer_tables(speconv, P1, P2, one_d, s=None):
    if one_d:
        return speconv.tables(P2, s)
    return speconv.tables(P1, P2, s)


def run_layers(state, speconvs, ws, act_name, act_flags, one_d, last_s=None, end_takes_grad=False):
    """Apply the Fourier layers to ``state`` (a Pending over a (B, P1, P2, C) tensor; P1 = 1 in 1-D).
    ``act_flags[i]`` says whether layer i is followed by the activation.  ``last_s`` (padded output
    resolution) switches the LAST layer to the resolution-changing form
    ``w(projector(x, s)) + speconv(x, s)`` (func_to_func2d.py:95).  ``end_takes_grad``: what follows the trunk is
    ``functional_end`` on its fused path (``functional_end_is_fused``), which like a fused layer reads GELU(z) and
    GELU'(z) only -- so the last layer may write the derivative in place of z too."""
    n = len(speconvs)
    fused_gelu = act_name == "gelu"

    def changes(P1, P2):
        if last_s is None:
            return False
        target = (int(last_s),) if one_d else tuple(int(v) for v in last_s)
        return target != ((P2,) if one_d else (P1, P2))

    for i, (sc, w) in enumerate(zip(speconvs, ws)):
        z_in = state.z
        _, P1, P2, _ = z_in.shape
        w1 = None if one_d else sc.weights2
        last = i == n - 1
        change = last and changes(P1, P2)
        zgrad = False
        if change:
            x = state.value()
            z, _ = ops.fourier_layer(x, sc.weights1, w1, None, None, _layer_tables(sc, P1, P2, one_d, last_s))
            # the pointwise branch acts on the Fourier-resampled input: one table GEMM per axis,
            # then the 1x1 conv on the (small) output grid
            if _fused_ends_ok(x) and w.in_channels % 4 == 0 and w.out_channels % 4 == 0 and \
                    max(w.in_channels, w.out_channels) <= 128:
                xc = ops.project_resolution(x, last_s, one_d)
                z = z + ops.pointwise_conv(xc, w.weight.view(w.out_channels, w.in_channels), w.bias)
            elif one_d:
                xc = projector1d(x.squeeze(1).permute(0, 2, 1), last_s)
                z = z + w(xc).permute(0, 2, 1).unsqueeze(1)
            else:
                xc = projector2d(x.permute(0, 3, 1, 2), last_s)
                z = z + w(xc).permute(0, 2, 3, 1)
        else:
            xo = None
            # the activated copy is produced in this layer's epilogue when another fused layer consumes it
            # (or, after the last layer, the F2V end: both of its branches read GELU(z))
            want = bool(act_flags[i]) and fused_gelu
            # ... and when that consumer is a plain fused one (the grid does not change inside the trunk), z itself is never
            # needed again: the epilogue writes GELU'(z) in its place and the consumer's backward multiplies by it
            zgrad = want and _CACHE_GELU_GRAD and torch.is_grad_enabled() and \
                (end_takes_grad if last else not (i + 1 == n - 1 and changes(P1, P2)))
            z, xo = ops.fourier_layer(z_in, sc.weights1, w1, w.weight, w.bias, _layer_tables(sc, P1, P2, one_d),
                                      act_in=state.act, xact=state.x, want_act_out=want, pre_is_grad=state.zgrad,
                                      grad_out=zgrad)
        if not act_flags[i]:
            state = Pending(z, False)
        elif fused_gelu:
            state = Pending(z, True, xo if not change else None, zgrad)
        else:
            state = Pending(_TORCH_ACTS[act_name](z), False)
    return state


def _fused_ends_ok(*tensors):
    return all(t.is_cuda and t.dtype == torch.float32 for t in tensors)


def lift(x, fc0, padding, get_grid, one_d):
    """Lifting + zero padding at the END of each axis (func_to_func2d.py:79-86) -> channels-last
    (B, P1, P2, C) (P1 = 1 in 1-D).  On CUDA one fused kernel (fc0 on data + grid coordinates, written
    padded); otherwise, or when the input itself needs a gradient, the plain torch statement."""
    from ..shared import get_grid1d, get_grid2d
    ngrid = (1 if one_d else 2) if get_grid else 0
    C = fc0.out_features
    if (_fused_ends_ok(x, fc0.weight) and not x.requires_grad and fc0.bias is not None and C % 4 == 0 and C <= 512
            and 1 <= x.shape[1] + ngrid <= 8 and x.shape[1] + ngrid == fc0.in_features):
        if one_d:
            n = x.shape[-1]
            return ops.lift(x.unsqueeze(2), fc0.weight, fc0.bias, 1, n + n // padding, ngrid), (n,)
        n1, n2 = x.shape[-2:]
        return ops.lift(x, fc0.weight, fc0.bias, n1 + n1 // padding, n2 + n2 // padding, ngrid), (n1, n2)
    if one_d:
        n = x.shape[-1]
        x = x.permute(0, 2, 1)
        if get_grid:
            x = torch.cat((x, get_grid1d(x.shape, x.device).to(x.dtype)), dim=-1)
        x = fc0(x)
        x = F.pad(x, [0, 0, 0, n // padding])
        return x.unsqueeze(1), (n,)
    n1, n2 = x.shape[-2:]
    x = x.permute(0, 2, 3, 1)
    if get_grid:
        x = torch.cat((x, get_grid2d(x.shape, x.device).to(x.dtype)), dim=-1)
    x = fc0(x)
    x = F.pad(x, [0, 0, 0, n2 // padding, 0, n1 // padding])
    return x, (n1, n2)


def crop(x, cut, one_d):
    """Remove the padding (func_to_func2d.py:98-101).  Like the reference's ``:-c`` slice, c == 0
    yields an empty axis."""
    if one_d:
        return x[:, :, :-cut[0], :] if cut[0] != 0 else x[:, :, :0, :]
    x = x[:, :-cut[0], :, :] if cut[0] != 0 else x[:, :0, :, :]
    return x[:, :, :-cut[1], :] if cut[1] != 0 else x[:, :, :0, :]


def project(x, cut, mlp, one_d):
    """Crop + pointwise MLP (func_to_func2d.py:98-107) on a channels-last (B, P1, P2, C) tensor.  With
    GELU on CUDA: fc1 runs on the fused tcgen05 pointwise kernel over the PADDED grid (no crop copy; the
    padded positions get zero gradient from the crop's backward), GELU + fc2 in one pass that never
    stores the hidden activation; the crop is a view of the small result."""
    fc1, fc2 = mlp.fc1, mlp.fc2
    if (getattr(mlp, "act_name", None) == "gelu" and _fused_ends_ok(x, fc1.weight, fc2.weight)
            and fc1.in_features % 4 == 0 and fc1.in_features <= 128 and fc1.out_features % 4 == 0
            and fc1.out_features <= 128 and fc2.out_features <= 4 and min(cut) > 0):
        h1 = ops.pointwise_conv(x, fc1.weight, fc1.bias)
        return crop(ops.mlp_head(h1, fc2.weight, fc2.bias), cut, one_d)
    return mlp(crop(x, cut, one_d))


def functional_end_is_fused(x, mlpfunc0, one_d, rows=None):
    """Whether ``functional_end`` will take its fused local branch for a trunk state shaped like ``x`` (B, P1, P2, C)
    (``rows``: the batch extent it will see, when the caller folds an axis into the batch)."""
    fc1, fc2 = mlpfunc0.fc1, mlpfunc0.fc2
    _, P1, P2, _ = x.shape
    return (getattr(mlpfunc0, "act_name", None) == "gelu" and _fused_ends_ok(x, fc1.weight, fc2.weight)
            and fc1.in_features % 4 == 0 and fc1.in_features <= 128 and fc1.out_features % 4 == 0
            and fc1.out_features <= 128 and P2 >= 2 and (one_d or P1 >= 2)
            and (x.shape[0] if rows is None else rows) <= 65535)            # wsum kernels put the batch on grid.y


def functional_end(state, lfunc0, mlpfunc0, head, one_d):
    """F2V end (func_to_vec2d.py:97-110): nonlocal functionals || pointwise MLP + trapezoid rule."""
    z = state.z
    _, P1, P2, _ = z.shape
    tables = lfunc0.tables(P2) if one_d else lfunc0.tables(P1, P2)
    nonlocal_feats = ops.linear_functional(z, lfunc0.weights, tables, act_in=state.act, xact=state.x,
                                           pre_is_grad=state.zgrad)
    fc1, fc2 = mlpfunc0.fc1, mlpfunc0.fc2
    if functional_end_is_fused(z, mlpfunc0, one_d):
        # local branch without the (B, P1, P2, width_lfunc) tensors: fc1 on the fused pointwise kernel, then the
        # trapezoid-weighted GELU sum; fc2 is linear, so it commutes with the integration and acts on (B, H)
        wy = _trapz_weights(P2, z.device)
        wx = _trapz_weights(P1, z.device) if not one_d else torch.ones(1, device=z.device)
        cached = state.act and state.x is not None
        xin = state.x if cached else state.value()
        if _ONE_PASS_LOCAL and fc1.bias is not None and ops.local_functional_supported(xin, fc1.weight):
            # fc1 + GELU + integration in ONE kernel: the (B, P1, P2, width_final) hidden tensor is never written in forward
            S = ops.local_functional(xin, fc1.weight, fc1.bias, wx, wy, zpre=z if cached else None, zpre_is_grad=state.zgrad)
        else:
            # fc1 reads the cached GELU(z); its backward multiplies the input gradient by GELU'(z) in the epilogue
            h1 = ops.pointwise_conv(xin, fc1.weight, fc1.bias, zpre=z if cached else None, zpre_is_grad=state.zgrad)
            S = ops.weighted_gelu_sum(h1, wx, wy)
        h = F.linear(S, fc2.weight)
        if fc2.bias is not None:
            h = h + fc2.bias * (((P2 - 1.0) / P2) * (1.0 if one_d else (P1 - 1.0) / P1))   # integral of the constant
    else:
        h = mlpfunc0(state.value())                           # (B, P1, P2, width_lfunc)
        h = torch.trapz(h, dx=1.0 / P2, dim=2)
        h = h.squeeze(1) if one_d else torch.trapz(h, dx=1.0 / P1, dim=1)
    return head(torch.cat((nonlocal_feats, h), dim=1))


_TRAPZ = {}


def _trapz_weights(P, device):
    """torch.trapz(., dx=1/P) over P points as a weight vector: dx * (1/2, 1, ..., 1, 1/2)."""
    key = (P, str(device))
    if key not in _TRAPZ:
        w = torch.full((P,), 1.0 / P, dtype=torch.float32)
        w[0] *= 0.5
        w[-1] *= 0.5
        _TRAPZ[key] = w.to(device)
    return _TRAPZ[key]
