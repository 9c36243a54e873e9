"""Drop-in replacements for the Fourier building blocks of the reference
(/root/reference/models/shared.py): same class names, constructor arguments, ``forward``
signatures, parameter names / shapes / dtypes / initialisation, and error behaviour -- but every
``forward`` runs hand-written sm_100a kernels (forward and backward) through the C ABI instead
of ``torch.fft`` + ``einsum``.

Inputs are accepted in the reference's logical layout ``(batch, channels, ..., nx[, ny])`` with
any strides; the kernels work channels-last, so a channels-last tensor (what the 2-D reference
trunks physically carry) passes through without a copy.  Only values are part of parity, not
output strides.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .plan import decoder_tables, functional_tables, spectral_tables


def _get_act(act):
    """Activation by name (reference: shared.py:7-23; same names, same ValueError)."""
    table = {"tanh": torch.tanh, "gelu": F.gelu, "relu": F.relu_, "elu": F.elu_, "leaky_relu": F.leaky_relu_}
    if act not in table:
        raise ValueError(f"{act} is not supported")
    return table[act]


class MLP(nn.Module):
    """Pointwise one-hidden-layer network on the last axis (reference: shared.py:26-45)."""

    def __init__(self, channels_in, channels_hid, channels_out, act="gelu"):
        super().__init__()
        self.fc1 = nn.Linear(channels_in, channels_hid)
        self.act_name = act
        self.act = _get_act(act)
        self.fc2 = nn.Linear(channels_hid, channels_out)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


# ------------------------------------------------------------------------------------------
# layout helpers
# ------------------------------------------------------------------------------------------
def _fold_extra_dims(x, n_spatial):
    """(B, C, E..., spatial) -> (B*prod(E), C, spatial) plus what is needed to undo it.  The
    ellipsis of the reference's einsum ("bi...,io...->bo...") broadcasts the weights over E."""
    extra = tuple(x.shape[2:x.dim() - n_spatial])
    if not extra:
        return x, None
    B, Cc = x.shape[:2]
    spatial = tuple(x.shape[x.dim() - n_spatial:])
    ne = len(extra)
    perm = [0] + list(range(2, 2 + ne)) + [1] + list(range(2 + ne, x.dim()))
    folded = x.permute(perm).reshape((-1, Cc) + spatial)
    return folded, (B, extra)


def _unfold_extra_dims(y, info, n_tail):
    """(B*prod(E), Co, tail...) -> (B, Co, E..., tail...)."""
    if info is None:
        return y
    B, extra = info
    Co = y.shape[1]
    tail = tuple(y.shape[2:])
    ne = len(extra)
    y = y.reshape((B,) + extra + (Co,) + tail)
    perm = [0, 1 + ne] + list(range(1, 1 + ne)) + list(range(2 + ne, y.dim()))
    return y.permute(perm)


def _to_cl(x):
    """(B, C, P1, P2) or (B, C, P) logical -> channels-last contiguous (B, P1|1, P2, C)."""
    if x.dim() == 3:
        return x.permute(0, 2, 1).unsqueeze(1).contiguous()
    return x.permute(0, 2, 3, 1).contiguous()


def _from_cl(y, one_d):
    return y.squeeze(1).permute(0, 2, 1) if one_d else y.permute(0, 3, 1, 2)


def _complex_param(scale, *shape, mode_major=False):
    """scale * U[0,1) on real and imaginary parts (shared.py:311-313).  ``mode_major``: same logical shape
    (in, out, modes...) and values, but stored with the mode axes slowest -- the order the mixing kernels
    contract against (include/fnm_b200.h: FNM_W_MODE_MAJOR), so no per-call layout copies are needed.
    state_dict / load_state_dict / checkpoints see the reference's shapes and values; only strides differ."""
    w = scale * torch.rand(*shape, dtype=torch.cfloat)
    if mode_major:
        n = len(shape)
        fwd = tuple(range(2, n)) + (0, 1)                     # (modes..., in, out)
        inv = (n - 2, n - 1) + tuple(range(0, n - 2))
        w = w.permute(fwd).contiguous().permute(inv)
    return nn.Parameter(w)


# ------------------------------------------------------------------------------------------
# reference helpers kept for API completeness (pure index bookkeeping / torch ops; they are not
# on the kernel path -- the kernels get the same rules through plan.py)
# ------------------------------------------------------------------------------------------
def compl_mul(input_tensor, weights):
    """(batch, in, ...), (in, out, ...) -> (batch, out, ...) complex product-sum (shared.py:48-53)."""
    return torch.einsum("bi...,io...->bo...", input_tensor, weights)


def resize_rfft(ar, s):
    """One-sided spectrum resize (shared.py:61-80)."""
    from .plan import onesided_source
    src = torch.as_tensor(onesided_source(ar.shape[-1], s), device=ar.device)
    out = ar[..., src.clamp(min=0)]
    return torch.where(src >= 0, out, torch.zeros((), dtype=out.dtype, device=ar.device))


def resize_fft(ar, s):
    """Two-sided spectrum resize (shared.py:83-108)."""
    from .plan import twosided_source
    src = torch.as_tensor(twosided_source(ar.shape[-1], s), device=ar.device)
    out = ar[..., src.clamp(min=0)]
    return torch.where(src >= 0, out, torch.zeros((), dtype=out.dtype, device=ar.device))


def resize_rfft2(ar, s):
    """2-D resize (shared.py:136-147)."""
    s1, s2 = s
    return resize_fft(resize_rfft(ar, s2).transpose(-2, -1), s1).transpose(-2, -1)


def get_grid1d(shape, device):
    """(B, nx, 1) coordinates in [0, 1] (shared.py:111-118)."""
    g = torch.linspace(0, 1, shape[1], device=device)
    return g.reshape(1, shape[1], 1).expand(shape[0], -1, -1)


def get_grid2d(shape, device):
    """(B, nx, ny, 2) coordinates in [0, 1]^2 (shared.py:150-159)."""
    gx = torch.linspace(0, 1, shape[1], device=device).reshape(1, shape[1], 1, 1).expand(shape[0], -1, shape[2], -1)
    gy = torch.linspace(0, 1, shape[2], device=device).reshape(1, 1, shape[2], 1).expand(shape[0], shape[1], -1, -1)
    return torch.cat((gx, gy), dim=-1)


def projector1d(x, s=None):
    """Fourier resampling to resolution s (shared.py:121-128); adjacent to the hot path."""
    if s is not None and s != x.shape[-1]:
        x = torch.fft.irfft(resize_rfft(torch.fft.rfft(x, norm="forward"), s), n=s, norm="forward")
    return x


def projector2d(x, s=None):
    """Fourier resampling to resolution s (shared.py:162-169); adjacent to the hot path."""
    if s is not None and tuple(s) != tuple(x.shape[-2:]):
        x = torch.fft.irfft2(resize_rfft2(torch.fft.rfft2(x, norm="forward"), s), s=tuple(s), norm="forward")
    return x


# ------------------------------------------------------------------------------------------
# Fourier layers
# ------------------------------------------------------------------------------------------
class SpectralConv1d(nn.Module):
    """Reference: shared.py:177-216.  weights1: (in, out, modes1) complex64."""

    def __init__(self, in_channels, out_channels, modes1):
        super().__init__()
        self.in_channels, self.out_channels, self.modes1 = in_channels, out_channels, modes1
        self.scale = 1.0 / (in_channels * out_channels)
        self.weights1 = _complex_param(self.scale, in_channels, out_channels, modes1, mode_major=True)

    def tables(self, P, s=None):
        return spectral_tables(1, P, 0, self.modes1, 1, P if s is None else int(s), one_d=True)

    def forward(self, x, s=None):
        x, info = _fold_extra_dims(x, 1)
        y, _ = ops.fourier_layer(_to_cl(x), self.weights1, None, None, None, self.tables(x.shape[-1], s))
        return _unfold_extra_dims(_from_cl(y, True), info, 1)


class SpectralConv2d(nn.Module):
    """Reference: shared.py:297-341.  weights1 / weights2: (in, out, modes1, modes2) complex64."""

    def __init__(self, in_channels, out_channels, modes1, modes2):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.modes1, self.modes2 = modes1, modes2
        self.scale = 1.0 / (in_channels * out_channels)
        self.weights1 = _complex_param(self.scale, in_channels, out_channels, modes1, modes2, mode_major=True)
        self.weights2 = _complex_param(self.scale, in_channels, out_channels, modes1, modes2, mode_major=True)

    def tables(self, P1, P2, s=None):
        S1, S2 = (P1, P2) if s is None else (int(s[0]), int(s[1]))
        return spectral_tables(P1, P2, self.modes1, self.modes2, S1, S2)

    def forward(self, x, s=None):
        x, info = _fold_extra_dims(x, 2)
        y, _ = ops.fourier_layer(_to_cl(x), self.weights1, self.weights2, None, None,
                                 self.tables(x.shape[-2], x.shape[-1], s))
        return _unfold_extra_dims(_from_cl(y, False), info, 2)


class LinearFunctionals1d(nn.Module):
    """F2V, reference: shared.py:219-254.  weights: (in, out, modes1 + 1) complex64."""

    def __init__(self, in_channels, out_channels, modes1):
        super().__init__()
        self.in_channels, self.out_channels, self.modes1 = in_channels, out_channels, modes1
        self.scale = 1.0 / (in_channels * out_channels)
        self.weights = _complex_param(self.scale, in_channels, out_channels, modes1 + 1, mode_major=True)

    def tables(self, P):
        return functional_tables(1, P, 0, self.modes1, one_d=True)

    def forward(self, x):
        x, info = _fold_extra_dims(x, 1)
        out = ops.linear_functional(_to_cl(x), self.weights, self.tables(x.shape[-1]))
        return _unfold_extra_dims(out, info, 0)


class LinearFunctionals2d(nn.Module):
    """F2V, reference: shared.py:344-383.  weights: (in, out, 2*modes1, modes2 + 1) complex64."""

    def __init__(self, in_channels, out_channels, modes1, modes2):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.modes1, self.modes2 = modes1, modes2
        self.scale = 1.0 / (in_channels * out_channels)
        self.weights = _complex_param(self.scale, in_channels, out_channels, 2 * modes1, modes2 + 1, mode_major=True)

    def tables(self, P1, P2):
        return functional_tables(P1, P2, self.modes1, self.modes2)

    def forward(self, x):
        x, info = _fold_extra_dims(x, 2)
        out = ops.linear_functional(_to_cl(x), self.weights, self.tables(x.shape[-2], x.shape[-1]))
        return _unfold_extra_dims(out, info, 0)


class LinearDecoder1d(nn.Module):
    """V2F, reference: shared.py:257-289.  weights: (in, out, modes1 + 1) complex64."""

    def __init__(self, in_channels, out_channels, modes1):
        super().__init__()
        self.in_channels, self.out_channels, self.modes1 = in_channels, out_channels, modes1
        self.scale = 1.0 / (in_channels * out_channels)
        self.weights = _complex_param(self.scale, in_channels, out_channels, modes1 + 1, mode_major=True)

    def tables(self, s):
        return decoder_tables(1, int(s), 0, self.modes1, one_d=True)

    def forward(self, x, s):
        x, info = _fold_extra_dims(x, 0)
        y = ops.linear_decoder(x, self.weights, self.tables(s))
        return _unfold_extra_dims(_from_cl(y, True), info, 1)


class LinearDecoder2d(nn.Module):
    """V2F, reference: shared.py:386-419.  weights: (in, out, 2*modes1, modes2 + 1) complex64."""

    def __init__(self, in_channels, out_channels, modes1, modes2):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.modes1, self.modes2 = modes1, modes2
        self.scale = 1.0 / (in_channels * out_channels)
        self.weights = _complex_param(self.scale, in_channels, out_channels, 2 * modes1, modes2 + 1, mode_major=True)

    def tables(self, s):
        return decoder_tables(int(s[0]), int(s[1]), self.modes1, self.modes2)

    def forward(self, x, s):
        x, info = _fold_extra_dims(x, 0)
        y = ops.linear_decoder(x, self.weights, self.tables(s))
        return _unfold_extra_dims(_from_cl(y, False), info, 2)
