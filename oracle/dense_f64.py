"""float64 dense-DFT restatement of the Fourier blocks  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference cannot run in float64 (its spectral buffers are hard-coded ``torch.cfloat``,
/root/reference/models/shared.py:73,95,207,329), so this file restates each block as explicit
sums over the retained modes in numpy float64/complex128.  It is the
high-precision arbiter when the fp32 oracle (``fnm_oracle.py``) and the CUDA path disagree
at the 1e-6 level, and it states the hand-derived backward formulas the CUDA kernels
implement.  Pinned by ``tests/test_oracle_golden.py`` against fixtures produced by the
reference itself (``tests/golden/make_golden.py``).

Mode bookkeeping (shared.py:61-108,136-147) is reproduced by pushing INDEX arrays through the
same pad/cut rules: every weight row/column ends up with (a) the FFT bin it reads on the
input grid and (b) the bin it lands in on the output grid, or "dead" (-1).
"""
import numpy as np


# --------------------------------------------------------------------------------------
# index bookkeeping
# --------------------------------------------------------------------------------------
def half_map(n_have, s):
    """Index map of resize_rfft (shared.py:61-80): out[j] = source column or -1."""
    if s < 1:
        raise ValueError("s must be greater than or equal to 1.")
    want = s // 2 + 1
    src = np.full(want, -1, dtype=np.int64)
    k = min(want, n_have)
    src[:k] = np.arange(k)
    return src


def full_map(n_have, s):
    """Index map of resize_fft (shared.py:83-108): out[j] = source row or -1."""
    if s < 1:
        raise ValueError("s must be greater than or equal to 1.")
    idx = np.arange(n_have)
    if s >= n_have:
        h = n_have // 2
        return np.concatenate((idx[:h], np.full(s - n_have, -1, dtype=np.int64), idx[h:]))
    if s == 1:
        return idx[:1]
    lo = s // 2 + 1 if s % 2 else s // 2
    return np.concatenate((idx[:lo], idx[n_have - (s - lo):]))


def _invert(src_map, n_src):
    """src_map[j] = source index at destination j  ->  dst[i] = destination of source i or -1."""
    dst = np.full(n_src, -1, dtype=np.int64)
    for j, i in enumerate(src_map):
        if i >= 0:
            dst[i] = j
    return dst


def spectral_rows(p1, m1, s1=None):
    """x-axis of SpectralConv2d: weight rows 0..m1-1 (weights1) read bins 0..m1-1, rows m1..2m1-1
    (weights2) read bins p1-m1..p1-1 (shared.py:330-333).  Returns (bin_in, bin_out)."""
    bin_in = np.concatenate((np.arange(m1), np.arange(p1 - m1, p1)))
    if s1 is None or s1 == p1:
        return bin_in, bin_in.copy()
    dst = _invert(full_map(p1, s1), p1)
    return bin_in, dst[bin_in]


def spectral_cols(p2, m2, s2=None):
    """y-axis (one-sided): columns 0..m2-1."""
    bin_in = np.arange(m2)
    if s2 is None or s2 == p2:
        return bin_in, bin_in.copy()
    dst = _invert(half_map(p2 // 2 + 1, s2), p2 // 2 + 1)
    return bin_in, dst[bin_in]


def functional_rows(p1, m1):
    """F2V x-axis: resize_fft(N=p1 -> 2*m1): row r of the weight reads bin full_map[r]."""
    return full_map(p1, 2 * m1)


def functional_cols(p2, m):
    """F2V y-axis (or 1-D): resize_rfft(N=p2//2+1 -> 2*m): m+1 columns."""
    return half_map(p2 // 2 + 1, 2 * m)


def decoder_rows(m1, s1):
    """V2F x-axis: resize_fft(N=2*m1 -> s1): weight row r lands in bin dst[r] (or is cut)."""
    return _invert(full_map(2 * m1, s1), 2 * m1)


def decoder_cols(m, s2):
    """V2F y-axis (or 1-D): resize_rfft(N=m+1 -> s2): column l lands in l if l < s2//2+1."""
    return _invert(half_map(m + 1, s2), m + 1)


def herm_weight(l_bins, n):
    """c_l of the C2R transform on an n-point axis: 1 for DC and Nyquist, else 2; 0 for dead."""
    c = np.where(l_bins < 0, 0.0, 2.0)
    c[l_bins == 0] = 1.0
    if n % 2 == 0:
        c[l_bins == n // 2] = 1.0
    return c


def _phase(bins, n, sign):
    """exp(sign * 2 pi i * bin * j / n), shape (len(bins), n); dead bins give zero rows."""
    j = np.arange(n)
    ang = 2.0 * np.pi * ((np.maximum(bins, 0)[:, None] * j[None, :]) % n) / n
    e = np.exp(sign * 1j * ang)
    e[bins < 0, :] = 0.0
    return e


# --------------------------------------------------------------------------------------
# generic analysis / synthesis on channel-first arrays (B, C, P1, P2)
# --------------------------------------------------------------------------------------
def analyse(x, rows_in, cols_in, scale=1.0):
    """xh[b,c,r,l] = scale * sum_{x,y} x[b,c,x,y] e^{-i(theta_x(rows_in[r],x) + theta_y(cols_in[l],y))}."""
    p1, p2 = x.shape[-2:]
    ex = _phase(rows_in, p1, -1.0)          # (R, P1)
    ey = _phase(cols_in, p2, -1.0)          # (L, P2)
    t = np.einsum("bcxy,ly->bcxl", x.astype(np.complex128), ey)
    return scale * np.einsum("rx,bcxl->bcrl", ex, t)


def synthesise(oh, rows_out, cols_out, s1, s2, scale):
    """y[b,c,x,y] = scale * sum_l c_l Re( (sum_r oh[b,c,r,l] e^{+i theta_x}) e^{+i theta_y} )
    -- C2C inverse over x followed by C2R over y, which DROPS Im of the DC/Nyquist column."""
    fx = _phase(rows_out, s1, +1.0)         # (R, S1)
    fy = _phase(cols_out, s2, +1.0) * herm_weight(cols_out, s2)[:, None]
    z = np.einsum("bcrl,rx->bcxl", oh, fx)
    return scale * np.real(np.einsum("bcxl,ly->bcxy", z, fy))


def synth_adjoint(g, rows_out, cols_out, scale):
    """PyTorch-convention gradient of ``synthesise`` w.r.t. oh (A.2, first line)."""
    s1, s2 = g.shape[-2:]
    fx = _phase(rows_out, s1, +1.0)
    fy = _phase(cols_out, s2, +1.0) * herm_weight(cols_out, s2)[:, None]
    gz = np.einsum("bcxy,ly->bcxl", g.astype(np.complex128), np.conj(fy))
    return scale * np.einsum("bcxl,rx->bcrl", gz, np.conj(fx))


def analyse_adjoint(gxh, rows_in, cols_in, p1, p2, scale=1.0):
    """PyTorch-convention gradient of ``analyse`` w.r.t. the real input x (A.2, last line)."""
    ex = _phase(rows_in, p1, -1.0)
    ey = _phase(cols_in, p2, -1.0)
    gt = np.einsum("bcrl,rx->bcxl", gxh, np.conj(ex))
    return scale * np.real(np.einsum("bcxl,ly->bcxy", gt, np.conj(ey)))


# --------------------------------------------------------------------------------------
# blocks (2-D; the 1-D variants are the P1 = 1, single-row special case)
# --------------------------------------------------------------------------------------
def _as2d(x):
    return x[:, :, None, :]


def spectral_conv2d(x, w1, w2, s=None):
    p1, p2 = x.shape[-2:]
    m1, m2 = w1.shape[-2:]
    s1, s2 = (p1, p2) if s is None else s
    w = np.concatenate((w1, w2), axis=2)
    rin, rout = spectral_rows(p1, m1, s1)
    cin, cout = spectral_cols(p2, m2, s2)
    xh = analyse(x, rin, cin)
    oh = np.einsum("birl,iorl->borl", xh, w)
    return synthesise(oh, rout, cout, s1, s2, 1.0 / (p1 * p2))


def spectral_conv2d_backward(x, w1, w2, g, s=None):
    """Returns (dx, dw1, dw2) for upstream gradient g (A.2)."""
    p1, p2 = x.shape[-2:]
    m1, m2 = w1.shape[-2:]
    s1, s2 = (p1, p2) if s is None else s
    w = np.concatenate((w1, w2), axis=2)
    rin, rout = spectral_rows(p1, m1, s1)
    cin, cout = spectral_cols(p2, m2, s2)
    xh = analyse(x, rin, cin)
    gh = synth_adjoint(g, rout, cout, 1.0 / (p1 * p2))
    dw = np.einsum("birl,borl->iorl", np.conj(xh), gh)
    gxh = np.einsum("borl,iorl->birl", gh, np.conj(w))
    dx = analyse_adjoint(gxh, rin, cin, p1, p2)
    return dx, dw[:, :, :m1], dw[:, :, m1:]


def spectral_conv1d(x, w1, s=None):
    p = x.shape[-1]
    m = w1.shape[-1]
    s_ = p if s is None else s
    cin, cout = spectral_cols(p, m, s_)
    one = np.zeros(1, dtype=np.int64)
    xh = analyse(_as2d(x), one, cin)
    oh = np.einsum("birl,iol->borl", xh, w1)
    return synthesise(oh, one, cout, 1, s_, 1.0 / p)[:, :, 0, :]


def spectral_conv1d_backward(x, w1, g, s=None):
    p = x.shape[-1]
    m = w1.shape[-1]
    s_ = p if s is None else s
    cin, cout = spectral_cols(p, m, s_)
    one = np.zeros(1, dtype=np.int64)
    xh = analyse(_as2d(x), one, cin)
    gh = synth_adjoint(_as2d(g), one, cout, 1.0 / p)
    dw = np.einsum("birl,borl->iol", np.conj(xh), gh)
    gxh = np.einsum("borl,iol->birl", gh, np.conj(w1))
    return analyse_adjoint(gxh, one, cin, 1, p)[:, :, 0, :], dw


def _functional_coeff_2d(m1, ncol):
    """c(r,l) of shared.py:380-381: 2 everywhere, except c(0,0)=1 and c(r>=m1, 0)=0."""
    c = np.full((2 * m1, ncol), 2.0)
    c[m1:, 0] = 0.0
    c[0, 0] = 1.0
    return c


def linear_functionals2d(x, w):
    p1, p2 = x.shape[-2:]
    m1, m2 = w.shape[-2] // 2, w.shape[-1] - 1
    xh = analyse(x, functional_rows(p1, m1), functional_cols(p2, m2), 1.0 / (p1 * p2))
    c = _functional_coeff_2d(m1, m2 + 1)
    return np.einsum("rl,birl,iorl->bo", c, xh, w).real


def linear_functionals2d_backward(x, w, g):
    p1, p2 = x.shape[-2:]
    m1, m2 = w.shape[-2] // 2, w.shape[-1] - 1
    rows, cols = functional_rows(p1, m1), functional_cols(p2, m2)
    xh = analyse(x, rows, cols, 1.0 / (p1 * p2))
    c = _functional_coeff_2d(m1, m2 + 1)
    dw = np.einsum("rl,bo,birl->iorl", c, g, np.conj(xh))
    gxh = np.einsum("rl,bo,iorl->birl", c, g, np.conj(w))
    return analyse_adjoint(gxh, rows, cols, p1, p2, 1.0 / (p1 * p2)), dw


def linear_functionals1d(x, w):
    p = x.shape[-1]
    m = w.shape[-1] - 1
    xh = analyse(_as2d(x), np.zeros(1, dtype=np.int64), functional_cols(p, m), 1.0 / p)[:, :, 0]
    c = np.full(m + 1, 2.0)
    c[0] = 1.0
    return np.einsum("l,bil,iol->bo", c, xh, w).real


def linear_functionals1d_backward(x, w, g):
    p = x.shape[-1]
    m = w.shape[-1] - 1
    one, cols = np.zeros(1, dtype=np.int64), functional_cols(p, m)
    xh = analyse(_as2d(x), one, cols, 1.0 / p)[:, :, 0]
    c = np.full(m + 1, 2.0)
    c[0] = 1.0
    dw = np.einsum("l,bo,bil->iol", c, g, np.conj(xh))
    gxh = np.einsum("l,bo,iol->bil", c, g, np.conj(w))[:, :, None, :]
    return analyse_adjoint(gxh, one, cols, 1, p, 1.0 / p)[:, :, 0, :], dw


def linear_decoder2d(v, w, s):
    s1, s2 = s
    m1, m2 = w.shape[-2] // 2, w.shape[-1] - 1
    oh = np.einsum("bi,iorl->borl", v.astype(np.complex128), w)
    return synthesise(oh, decoder_rows(m1, s1), decoder_cols(m2, s2), s1, s2, 1.0 / (s1 * s2))


def linear_decoder2d_backward(v, w, g):
    s1, s2 = g.shape[-2:]
    m1, m2 = w.shape[-2] // 2, w.shape[-1] - 1
    gh = synth_adjoint(g, decoder_rows(m1, s1), decoder_cols(m2, s2), 1.0 / (s1 * s2))
    dw = np.einsum("bi,borl->iorl", v.astype(np.complex128), gh)
    dv = np.einsum("borl,iorl->bi", gh, np.conj(w)).real
    return dv, dw


def linear_decoder1d(v, w, s):
    m = w.shape[-1] - 1
    oh = np.einsum("bi,iol->bol", v.astype(np.complex128), w)[:, :, None, :]
    return synthesise(oh, np.zeros(1, dtype=np.int64), decoder_cols(m, s), 1, s, 1.0 / s)[:, :, 0, :]


def linear_decoder1d_backward(v, w, g):
    s = g.shape[-1]
    m = w.shape[-1] - 1
    gh = synth_adjoint(_as2d(g), np.zeros(1, dtype=np.int64), decoder_cols(m, s), 1.0 / s)[:, :, 0]
    dw = np.einsum("bi,bol->iol", v.astype(np.complex128), gh)
    dv = np.einsum("bol,iol->bi", gh, np.conj(w)).real
    return dv, dw


# --------------------------------------------------------------------------------------
# pointwise pieces of the fused layer
# --------------------------------------------------------------------------------------
def _erf(z):
    from math import erf
    return np.vectorize(erf)(z)


def gelu(z):
    return 0.5 * z * (1.0 + _erf(z / np.sqrt(2.0)))


def gelu_grad(z):
    return 0.5 * (1.0 + _erf(z / np.sqrt(2.0))) + z * np.exp(-0.5 * z * z) / np.sqrt(2.0 * np.pi)


def fourier_layer2d(x, w1, w2, wc, bc, act):
    """z = Conv1x1(x) + bias + SpectralConv2d(x); out = GELU(z) if act else z."""
    z = np.einsum("oi,bixy->boxy", wc, x) + bc[None, :, None, None] + spectral_conv2d(x, w1, w2)
    return (gelu(z) if act else z), z


def fourier_layer2d_backward(x, w1, w2, wc, z, g_out, act):
    gz = g_out * gelu_grad(z) if act else g_out
    dx_s, dw1, dw2 = spectral_conv2d_backward(x, w1, w2, gz)
    dwc = np.einsum("boxy,bixy->oi", gz, x)
    dbc = gz.sum(axis=(0, 2, 3))
    dx = dx_s + np.einsum("oi,boxy->bixy", wc, gz)
    return dx, dw1, dw2, dwc, dbc


# --------------------------------------------------------------------------------------
# Adam (util/Adam.py:25-51) in float64
# --------------------------------------------------------------------------------------
def adam_update(p, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=0.0, vmax=None):
    if wd != 0:
        g = g + wd * p
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * (g * np.conj(g))
    second = v
    if vmax is not None:
        vmax = np.maximum(vmax.real, v.real).astype(v.dtype)
        second = vmax
    denom = np.sqrt(second) / np.sqrt(1 - b2 ** t) + eps
    p = p - (lr / (1 - b1 ** t)) * m / denom
    return p, m, v, vmax
