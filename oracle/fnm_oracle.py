"""CPU oracle for the Fourier-layer hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product package never does.

What it is: a functional (no ``nn.Module``) restatement, in plain fp32 PyTorch CPU ops,
of the algorithm the reference implements for this path.  The arithmetic of the reference
lives in a third-party dependency (PyTorch ATen: ``torch.fft`` -> MKL/pocketfft,
``einsum`` -> MKL bmm, ``conv`` -> oneDNN; pinned by the reference at pytorch==1.9.1,
``/root/reference/Project.yml:13``; this image has torch 2.11), so the restatement calls
the same library entry points the reference calls, at the same call sites:

  * mixing                     /root/reference/models/shared.py:48-53
  * mode bookkeeping           /root/reference/models/shared.py:61-108, 136-147
  * SpectralConv1d / 2d        /root/reference/models/shared.py:193-216, 315-341
  * LinearFunctionals1d / 2d   /root/reference/models/shared.py:239-254, 365-383
  * LinearDecoder1d / 2d       /root/reference/models/shared.py:276-289, 406-419
  * activation table           /root/reference/models/shared.py:7-23
  * model glue                 /root/reference/models/func_to_func{1d,2d}.py,
                               func_to_vec{1d,2d}.py, vec_to_func{1d,2d}.py, vec_to_vec{1d,2d}.py
  * relative L2 loss           /root/reference/util/utilities_module.py:188-204
  * complex-aware Adam         /root/reference/util/Adam.py:25-51

Parity pinning: the reference ships no tests or golden vectors, so
this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: ``tests/golden/*.npz`` were
produced by importing the unmodified reference in the build container
(``tests/golden/make_golden.py``), and ``tests/test_oracle_golden.py`` checks every
function here against them.

Parameters are passed as a flat ``dict`` using the reference's ``state_dict`` key names.

Unlike the reference (whose spectral buffers are hard-coded complex64) the functions follow the
dtype of their inputs, so the same code run in float64 / complex128 gives the "truth" used to
measure the fp32 noise floor of ill-conditioned quantities (tests/helpers.py: ``assert_parity``).
"""
import math

import torch
import torch.fft as tfft
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# activation (shared.py:7-23)
# --------------------------------------------------------------------------------------
_ACTS = {
    "tanh": torch.tanh,
    "gelu": F.gelu,
    "relu": F.relu,
    "elu": F.elu,
    "leaky_relu": F.leaky_relu,
}


def activation(name):
    try:
        return _ACTS[name]
    except KeyError:
        raise ValueError(f"{name} is not supported")


# --------------------------------------------------------------------------------------
# mode bookkeeping (shared.py:61-108, 136-147)
# --------------------------------------------------------------------------------------
def _zeros_like_tail(a, n):
    return torch.zeros(a.shape[:-1] + (n,), dtype=a.dtype, device=a.device)


def keep_half_spectrum(a, s):
    """One-sided spectrum ``a`` (..., N) -> (..., s//2+1): zero-pad or cut the top bins."""
    if s < 1:
        raise ValueError("s must be greater than or equal to 1.")
    n_have, n_want = a.shape[-1], s // 2 + 1
    if n_want >= n_have:
        return torch.cat((a, _zeros_like_tail(a, n_want - n_have)), dim=-1)
    return a[..., :n_want]


def keep_full_spectrum(a, s):
    """Two-sided spectrum ``a`` (..., N) -> (..., s): zeros go in / bins come out at the middle."""
    if s < 1:
        raise ValueError("s must be greater than or equal to 1.")
    n_have = a.shape[-1]
    if s >= n_have:
        h = n_have // 2
        return torch.cat((a[..., :h], _zeros_like_tail(a, s - n_have), a[..., h:]), dim=-1)
    if s == 1:
        return a[..., 0:1]
    lo = s // 2 + 1 if s % 2 else s // 2          # bins kept from the front
    hi = s - lo                                   # bins kept from the back
    return torch.cat((a[..., :lo], a[..., n_have - hi:]), dim=-1)


def keep_spectrum2(a, s):
    s1, s2 = s
    a = keep_half_spectrum(a, s2)
    return keep_full_spectrum(a.transpose(-2, -1), s1).transpose(-2, -1)


def mix(xh, w):
    """(b, i, ...) x (i, o, ...) -> (b, o, ...), complex, no conjugation (shared.py:48-53)."""
    return torch.einsum("bi...,io...->bo...", xh, w)


# --------------------------------------------------------------------------------------
# Fourier blocks
# --------------------------------------------------------------------------------------
def spectral_conv1d(x, w1, s=None):
    """shared.py:193-216.  x (B, Ci, ..., P) real; w1 (Ci, Co, m) complex."""
    m = w1.shape[-1]
    p = x.shape[-1]
    xh = tfft.rfft(x)
    lead = list(x.shape[:-1])
    lead[1] = w1.shape[1]
    oh = torch.zeros(*lead, p // 2 + 1, dtype=xh.dtype, device=x.device)
    oh[..., :m] = mix(xh[..., :m], w1)
    if s is None or s == p:
        return tfft.irfft(oh, n=p)
    return tfft.irfft(keep_half_spectrum(oh, s), n=s, norm="forward") / p


def spectral_conv2d(x, w1, w2, s=None):
    """shared.py:315-341.  x (B, Ci, ..., P1, P2) real; w1/w2 (Ci, Co, m1, m2) complex."""
    m1, m2 = w1.shape[-2:]
    p1, p2 = x.shape[-2:]
    xh = tfft.rfft2(x)
    lead = list(x.shape[:-2])
    lead[1] = w1.shape[1]
    oh = torch.zeros(*lead, p1, p2 // 2 + 1, dtype=xh.dtype, device=x.device)
    oh[..., :m1, :m2] = mix(xh[..., :m1, :m2], w1)
    oh[..., p1 - m1:, :m2] = mix(xh[..., p1 - m1:, :m2], w2)
    if s is None or tuple(s) == (p1, p2):
        return tfft.irfft2(oh, s=(p1, p2))
    return tfft.irfft2(keep_spectrum2(oh, s), s=tuple(s), norm="forward") / (p1 * p2)


def linear_functionals1d(x, w):
    """shared.py:239-254.  x (B, Ci, ..., P) -> (B, Co, ...); w (Ci, Co, m+1) complex."""
    m = w.shape[-1] - 1
    xh = keep_half_spectrum(tfft.rfft(x, norm="forward"), 2 * m)
    r = mix(xh, w).real
    return 2 * r.sum(dim=-1) - r[..., 0]


def linear_functionals2d(x, w):
    """shared.py:365-383.  x (B, Ci, ..., P1, P2) -> (B, Co, ...); w (Ci, Co, 2*m1, m2+1)."""
    m1, m2 = w.shape[-2] // 2, w.shape[-1] - 1
    xh = keep_spectrum2(tfft.rfft2(x, norm="forward"), (2 * m1, 2 * m2))
    r = mix(xh, w).real
    top = r[..., :m1, :].sum(dim=(-2, -1))
    bot = r[..., r.shape[-2] - m1:, 1:].sum(dim=(-2, -1))
    return 2 * top + 2 * bot - r[..., 0, 0]


def linear_decoder1d(v, w, s):
    """shared.py:276-289.  v (B, Ci, ...) real -> (B, Co, ..., s); w (Ci, Co, m+1)."""
    oh = mix(v[..., None].to(w.dtype), w)
    return tfft.irfft(keep_half_spectrum(oh, s), n=s)


def linear_decoder2d(v, w, s):
    """shared.py:406-419.  v (B, Ci, ...) real -> (B, Co, ..., s1, s2); w (Ci, Co, 2*m1, m2+1)."""
    oh = mix(v[..., None, None].to(w.dtype), w)
    return tfft.irfft2(keep_spectrum2(oh, tuple(s)), s=tuple(s))


def project1d(x, s=None):
    """shared.py:121-128."""
    if s is None or s == x.shape[-1]:
        return x
    return tfft.irfft(keep_half_spectrum(tfft.rfft(x, norm="forward"), s), n=s, norm="forward")


def project2d(x, s=None):
    """shared.py:162-169."""
    if s is None or tuple(s) == tuple(x.shape[-2:]):
        return x
    return tfft.irfft2(keep_spectrum2(tfft.rfft2(x, norm="forward"), s), s=tuple(s), norm="forward")


def pointwise_conv(x, weight, bias):
    """nn.Conv1d / nn.Conv2d with a 1x1 kernel (func_to_func2d.py:62-65)."""
    if x.dim() == 3:
        return F.conv1d(x, weight, bias)
    return F.conv2d(x, weight, bias)


def fourier_layer(x, p, idx, act=None, s=None):
    """One layer of the trunk: ``w(x) + speconv(x)`` then optional activation
    (func_to_func2d.py:89-95)."""
    wc, bc = p[f"ws.{idx}.weight"], p[f"ws.{idx}.bias"]
    if x.dim() == 3:
        y = pointwise_conv(project1d(x, s), wc, bc) + spectral_conv1d(x, p[f"speconvs.{idx}.weights1"], s)
    else:
        y = pointwise_conv(project2d(x, s), wc, bc) + spectral_conv2d(
            x, p[f"speconvs.{idx}.weights1"], p[f"speconvs.{idx}.weights2"], s)
    return y if act is None else act(y)


def mlp(x, p, name, act):
    """shared.py:26-45 -- applied on the last axis."""
    return F.linear(act(F.linear(x, p[f"{name}.fc1.weight"], p[f"{name}.fc1.bias"])),
                    p[f"{name}.fc2.weight"], p[f"{name}.fc2.bias"])


def unit_grid(shape, device):
    """shared.py:111-118 / 150-159: coordinates in [0,1] (end points included), channels-last."""
    if len(shape) == 3:
        b, n1 = shape[0], shape[1]
        return torch.linspace(0, 1, n1).reshape(1, n1, 1).repeat(b, 1, 1).to(device)   # fp32 like the reference
    b, n1, n2 = shape[0], shape[1], shape[2]
    gx = torch.linspace(0, 1, n1).reshape(1, n1, 1, 1).repeat(b, 1, n2, 1)
    gy = torch.linspace(0, 1, n2).reshape(1, 1, n2, 1).repeat(b, n1, 1, 1)
    return torch.cat((gx, gy), dim=-1).to(device)


# --------------------------------------------------------------------------------------
# model forwards (functional; ``cfg`` holds the constructor arguments)
# --------------------------------------------------------------------------------------
def _count(p, stem):
    n = 0
    while f"{stem}.{n}.weight" in p:
        n += 1
    return n


def _lift(x, p, padding, get_grid=True):
    d = x.dim() - 2
    res = x.shape[2:]
    x = x.movedim(1, -1)
    if get_grid:
        x = torch.cat((x, unit_grid(x.shape, x.device)), dim=-1)
    x = F.linear(x, p["fc0.weight"], p["fc0.bias"]).movedim(-1, 1)
    pads = []
    for n in reversed(res):
        pads += [0, n // padding]
    return F.pad(x, pads), res


def _crop(x, cut):
    for ax, c in enumerate(cut):
        dim = x.dim() - len(cut) + ax
        # the reference slices ``:-c``; with c == 0 (grid smaller than ``padding``) that is an
        # EMPTY axis (func_to_func2d.py:101) -- kept as is
        x = x.narrow(dim, 0, x.shape[dim] - c if c > 0 else 0)
    return x


def f2f_forward(p, x, padding=8, act="gelu", s_outputspace=None, get_grid=True):
    """FNO1d / FNO2d (func_to_func1d.py:63-102, func_to_func2d.py:69-107)."""
    a = activation(act)
    d = x.dim() - 2
    x, res = _lift(x, p, padding, get_grid)
    n_layers = _count(p, "ws")
    s_pad, cut = None, tuple(n // padding for n in res)
    if s_outputspace is not None:
        s_list = [s_outputspace] if d == 1 else list(s_outputspace)
        s_pad = tuple(r + r // padding for r in s_list)
        cut = tuple(r // padding for r in s_list)
        if d == 1:
            s_pad = s_pad[0]
    for i in range(n_layers):
        last = i == n_layers - 1
        x = fourier_layer(x, p, i, None if last else a, s_pad if last else None)
    x = _crop(x, cut)
    return mlp(x.movedim(1, -1), p, "mlp0", a).movedim(-1, 1)


def f2v_forward(p, x, padding=8, act="gelu", get_grid=True):
    """FNF1d / FNF2d (func_to_vec1d.py:74-110, func_to_vec2d.py:74-112)."""
    a = activation(act)
    d = x.dim() - 2
    x, _ = _lift(x, p, padding, get_grid)
    for i in range(_count(p, "ws")):
        x = fourier_layer(x, p, i, a)
    return _functional_end(x, p, a, d, "mlp0")


def _functional_end(x, p, a, d, head):
    nonlocal_feats = linear_functionals1d(x, p["lfunc0.weights"]) if d == 1 else \
        linear_functionals2d(x, p["lfunc0.weights"])
    loc = mlp(x.movedim(1, -1), p, "mlpfunc0", a).movedim(-1, 1)
    for _ in range(d):
        loc = torch.trapz(loc, dx=1.0 / loc.shape[-1])
    return mlp(torch.cat((nonlocal_feats, loc), dim=1), p, head, a)


def _vector_start(p, v, a, nonlinear_first):
    if nonlinear_first:
        return mlp(v, p, "mlp0", a)
    return F.linear(v, p["mlp0.weight"], p["mlp0.bias"])


def v2f_forward(p, v, s_outputspace, padding=8, act="gelu", nonlinear_first=True):
    """FND1d / FND2d (vec_to_func1d.py:72-99, vec_to_func2d.py:78-104)."""
    a = activation(act)
    d = 1 if isinstance(s_outputspace, int) else 2
    s_list = [s_outputspace] if d == 1 else list(s_outputspace)
    s_pad = tuple(r + r // padding for r in s_list)
    cut = tuple(r // padding for r in s_list)
    x = _vector_start(p, v, a, nonlinear_first)
    x = linear_decoder1d(x, p["ldec0.weights"], s_pad[0]) if d == 1 else \
        linear_decoder2d(x, p["ldec0.weights"], s_pad)
    n = _count(p, "ws")
    for i in range(n):
        x = fourier_layer(x, p, i, None if i == n - 1 else a)
    x = _crop(x, cut)
    return mlp(x.movedim(1, -1), p, "mlp1", a).movedim(-1, 1)


def v2v_forward(p, v, s_latentspace, act="gelu", nonlinear_first=True):
    """FNN1d / FNN2d (vec_to_vec1d.py, vec_to_vec2d.py:86-118)."""
    a = activation(act)
    d = 1 if isinstance(s_latentspace, int) else 2
    x = _vector_start(p, v, a, nonlinear_first)
    x = linear_decoder1d(x, p["ldec0.weights"], s_latentspace) if d == 1 else \
        linear_decoder2d(x, p["ldec0.weights"], tuple(s_latentspace))
    for i in range(_count(p, "ws")):
        x = fourier_layer(x, p, i, a)
    return _functional_end(x, p, a, d, "mlp1")


def f1d_to_f2d_forward(p, x, padding=8, act="gelu", s_outputspace=None, get_grid=True):
    """FNO1d2 (func1d_to_func2d.py:78-121): lift a 1-D function, decode every grid point's channel vector into a
    function of a second axis of the UNPADDED input length (:97), 2-D trunk whose last layer changes resolution."""
    a = activation(act)
    x, res = _lift(x, p, padding, get_grid)
    n = res[0]
    x = linear_decoder1d(x, p["ldec0.weights"], n)                      # (B, width, P, n)
    s_pad, cut = None, (n // padding, n // padding)
    if s_outputspace is not None:
        s_pad = tuple(r + r // padding for r in s_outputspace)
        cut = tuple(r // padding for r in s_outputspace)
    nl = _count(p, "ws")
    for i in range(nl):
        last = i == nl - 1
        x = fourier_layer(x, p, i, None if last else a, s_pad if last else None)
    x = _crop(x, cut)
    return mlp(x.movedim(1, -1), p, "mlp0", a).movedim(-1, 1)


def f2d_to_f1d_forward(p, x, padding=8, act="gelu", s_outputspace=None, get_grid=True):
    """FNO2d1 (func2d_to_func1d.py:78-128): 2-D trunk (activation after every layer), functionals of the LAST axis
    for every x (:103-111), 1-D resolution change, crop (by ny // padding when no output resolution is set,
    :120), pointwise MLP."""
    a = activation(act)
    x, res = _lift(x, p, padding, get_grid)
    for i in range(_count(p, "ws")):
        x = fourier_layer(x, p, i, a)
    nonlocal_feats = linear_functionals1d(x, p["lfunc0.weights"])       # (B, width1d, P1)
    loc = mlp(x.movedim(1, -1), p, "mlpfunc0", a).movedim(-1, 1)
    loc = torch.trapz(loc, dx=1.0 / loc.shape[-1])
    x = torch.cat((nonlocal_feats, loc), dim=1)
    cut = res[-1] // padding
    if s_outputspace is not None:
        x = project1d(x, s_outputspace + s_outputspace // padding)
        cut = s_outputspace // padding
    x = _crop(x, (cut,))
    return mlp(x.movedim(1, -1), p, "mlp0", a).movedim(-1, 1)


# --------------------------------------------------------------------------------------
# loss and optimiser
# --------------------------------------------------------------------------------------
def rel_l2_sum(out, y, eps=1e-6):
    """LpLoss(size_average=False)(out, y)  (utilities_module.py:188-204): sum over the batch of
    ||out_b - y_b||_2 / (||y_b||_2 + eps)."""
    b = out.shape[0]
    num = torch.norm(out.reshape(b, -1) - y.reshape(b, -1), 2, 1)
    den = torch.norm(y.reshape(b, -1), 2, 1) + eps
    return torch.sum(num / den)


class AdamState:
    """Optimiser state of util/Adam.py: one (m, v[, vmax], t) record per parameter, complex
    parameters keep complex moments (Adam.py:130-138)."""

    def __init__(self):
        self.m, self.v, self.vmax, self.t = {}, {}, {}, {}


@torch.no_grad()
def adam_step(params, state, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
    """One step of the reference's complex-aware Adam (Adam.py:25-51) on ``params``: a dict or
    list of leaf tensors whose ``.grad`` is set.  Parameters without a gradient are skipped
    (Adam.py:122)."""
    items = params.items() if isinstance(params, dict) else enumerate(params)
    b1, b2 = betas
    for key, p in items:
        if p.grad is None:
            continue
        if key not in state.t:
            state.t[key] = 0
            state.m[key] = torch.zeros_like(p)
            state.v[key] = torch.zeros_like(p)
            if amsgrad:
                state.vmax[key] = torch.zeros_like(p)
        state.t[key] += 1
        t = state.t[key]
        g = p.grad
        if weight_decay != 0:
            g = g + weight_decay * p
        m, v = state.m[key], state.v[key]
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g.conj(), value=1 - b2)      # |g|^2: ONE moment per complex element
        if amsgrad:
            torch.maximum(state.vmax[key], v, out=state.vmax[key])
            second = state.vmax[key]
        else:
            second = v
        denom = (second.sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
        p.addcdiv_(m, denom, value=-(lr / (1 - b1 ** t)))


def train_step(forward, p, x, y, state, **adam_kw):
    """zero_grad -> forward -> rel-L2 sum loss -> backward -> Adam  (train_fno.py:187-197)."""
    for t in p.values():
        t.grad = None
    out = forward(p, x)
    loss = rel_l2_sum(out, y)
    loss.backward()
    adam_step(p, state, **adam_kw)
    return loss.detach(), out.detach()
