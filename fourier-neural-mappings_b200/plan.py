"""Mode plans: the host-side index bookkeeping of the reference turned into small twiddle tables.

The reference decides which Fourier bins a layer reads, where they land on the output grid and
which are dropped with tensor slicing / ``torch.cat`` / ``zeros`` inside ``resize_rfft``,
``resize_fft`` and ``resize_rfft2`` (/root/reference/models/shared.py:61-108, 136-147) and in the
corner-block assignments of ``SpectralConv2d.forward`` (shared.py:330-333).  Here the same rules
are applied once, on INDEX arrays, and baked into dense truncated-DFT tables (float64 trig,
rounded to fp32) that the CUDA kernels contract against; see include/fnm_b200.h for the layouts.
"""
import math
from functools import lru_cache

import os

import numpy as np
import torch

from ._lib import FNM_MODE_FP32, FnmPlan


# ------------------------------------------------------------------------------------------
# index maps
# ------------------------------------------------------------------------------------------
def onesided_source(n_have, s):
    """Destination j of the one-sided resize (shared.py:61-80) takes source column ``out[j]`` (-1: zero)."""
    if s < 1:
        raise ValueError("s must be greater than or equal to 1.")
    want = s // 2 + 1
    out = np.full(want, -1, dtype=np.int64)
    keep = min(want, n_have)
    out[:keep] = np.arange(keep)
    return out


def twosided_source(n_have, s):
    """Destination j of the two-sided resize (shared.py:83-108) takes source row ``out[j]`` (-1: zero)."""
    if s < 1:
        raise ValueError("s must be greater than or equal to 1.")
    src = np.arange(n_have, dtype=np.int64)
    if s >= n_have:
        head = n_have // 2
        return np.concatenate([src[:head], np.full(s - n_have, -1, dtype=np.int64), src[head:]])
    if s == 1:
        return src[:1]
    front = s // 2 + 1 if s % 2 else s // 2
    return np.concatenate([src[:front], src[n_have - (s - front):]])


def _destination(source_map, n_src):
    dst = np.full(n_src, -1, dtype=np.int64)
    alive = source_map >= 0
    dst[source_map[alive]] = np.nonzero(alive)[0]
    return dst


def _c2r_weight(bins, n):
    c = np.where(bins < 0, 0.0, 2.0)
    c[bins == 0] = 1.0
    if n % 2 == 0:
        c[bins == n // 2] = 1.0
    return c


def _cos_sin(bins, n):
    """cos / sin of 2*pi*bin*j/n for j < n, shape (len(bins), n); dead bins (-1) give zero rows.
    The product bin*j is reduced mod n in integers so large grids lose no precision, and the
    exact zeros of the quarter-period points are snapped."""
    j = np.arange(n, dtype=np.int64)
    k = (np.maximum(bins, 0)[:, None] * j[None, :]) % n
    ang = 2.0 * math.pi * k.astype(np.float64) / n
    c, s = np.cos(ang), np.sin(ang)
    s[(2 * k) % n == 0] = 0.0
    c[((4 * k) % n == 0) & ((2 * k) % n != 0)] = 0.0
    dead = bins < 0
    c[dead, :] = 0.0
    s[dead, :] = 0.0
    return c, s


def _embed_analysis(bins, n):
    """[(ro,r)][(j,ri)] real embedding of e_r(j) = exp(-2 pi i bin_r j / n)."""
    c, s = _cos_sin(bins, n)
    r = len(bins)
    t = np.zeros((2, r, n, 2))
    t[0, :, :, 0], t[0, :, :, 1] = c, s
    t[1, :, :, 0], t[1, :, :, 1] = -s, c
    return t.reshape(2 * r, 2 * n)


def _embed_synthesis(bins, n):
    """[(j,ro)][(ri,r)] real embedding of f_r(j) = exp(+2 pi i bin_r j / n)."""
    c, s = _cos_sin(bins, n)
    r = len(bins)
    t = np.zeros((n, 2, 2, r))
    t[:, 0, 0, :], t[:, 0, 1, :] = c.T, -s.T
    t[:, 1, 0, :], t[:, 1, 1, :] = s.T, c.T
    return t.reshape(2 * n, 2 * r)


class ModeTables:
    """Host tables of one (kind, grid, modes) combination; device copies are made on demand."""

    def __init__(self, kind, P1, P2, S1, S2, rows_in, rows_out, cols_in, cols_out, scale_in, scale_out,
                 rows_per_w, cc=None):
        self.kind = kind
        self.P1, self.P2, self.S1, self.S2 = P1, P2, S1, S2
        self.R, self.NY = len(rows_in if rows_in is not None else rows_out), \
            len(cols_in if cols_in is not None else cols_out)
        self.rows_per_w = rows_per_w
        self.rows_in, self.rows_out, self.cols_in, self.cols_out = rows_in, rows_out, cols_in, cols_out
        host = {}
        if cols_in is not None:
            c, s = _cos_sin(cols_in, P2)
            ay = np.concatenate([scale_in * c, -scale_in * s], axis=0)            # [(ri,l)][y]
            host["ay"], host["ayT"] = ay, ay.T
            host["ex"] = _embed_analysis(rows_in, P1)
            host["fxb"] = _embed_synthesis(rows_in, P1)
        if cols_out is not None:
            c, s = _cos_sin(cols_out, S2)
            w = (scale_out * _c2r_weight(cols_out, S2))[:, None]
            byT = np.concatenate([w * c, -w * s], axis=0)                         # [(ri,l)][y']
            host["byT"], host["by"] = byT, byT.T
            host["fx"] = _embed_synthesis(rows_out, S1)
            host["exb"] = _embed_analysis(rows_out, S1)
        if cc is not None:
            host["cc"] = cc
        self.host = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in host.items()}
        self._dev = {}

    def device(self, device):
        key = str(device)
        if key not in self._dev:
            d = {k: torch.from_numpy(v).to(device) for k, v in self.host.items()}
            for name in ("ay", "byT"):                        # TMA needs 16-byte row strides (fnm_plan.ay4 / byT4)
                if name in d and d[name].shape[1] % 4:
                    t = d[name]
                    d[name + "4"] = torch.nn.functional.pad(t, (0, -t.shape[1] % 4)).contiguous()
            if os.environ.get("FNM_PW2_CORR", "")[:1] == "b":   # opt-in bf16 correction passes of the fused pointwise kernel
                for name in ("by", "ayT"):
                    if name in d:
                        d[name + "16"] = bf16_table_pair(d[name])
            self._dev[key] = d
        return self._dev[key]

    def c_plan(self, device, B, Ci, Co, mode=FNM_MODE_FP32):
        d = self.device(device)
        p = FnmPlan()
        p.B, p.P1, p.P2, p.S1, p.S2 = B, self.P1, self.P2, self.S1, self.S2
        p.R, p.NY, p.Ci, p.Co, p.rows_per_w, p.mode = self.R, self.NY, Ci, Co, self.rows_per_w, mode
        for name in ("ay", "ayT", "by", "byT", "ex", "fx", "exb", "fxb", "by16", "ayT16", "ay4", "byT4"):
            setattr(p, name, d[name].data_ptr() if name in d else None)
        return p


def bf16_table_pair(tab):
    """[rows][cols] fp32 table -> [2][rows][cols8] bfloat16 (cols8 = cols rounded up to 8, zero padded): bf16(tab) and
    bf16(tab - trunc_tf32(tab)), the two shared-memory operands of the bf16 correction passes of the fused pointwise
    kernel (include/fnm_b200.h, fnm_plan.by16).  They depend on the table only, so they are made once here instead of
    being re-derived from the fp32 tile by the kernel's converter warps for every tile."""
    rows, cols = tab.shape
    cols8 = (cols + 7) // 8 * 8
    hi = (tab.view(torch.int32) & -8192).view(torch.float32)               # what the tensor core reads: 13 mantissa bits dropped
    out = torch.zeros((2, rows, cols8), dtype=torch.bfloat16, device=tab.device)
    out[0, :, :cols] = tab.to(torch.bfloat16)
    out[1, :, :cols] = (tab - hi).to(torch.bfloat16)
    return out


@lru_cache(maxsize=256)
def spectral_tables(P1, P2, m1, m2, S1=None, S2=None, one_d=False):
    """SpectralConv1d (one_d: P1 = 1, weights1 only) / SpectralConv2d (weights1 | weights2)."""
    S1 = P1 if S1 is None else S1
    S2 = P2 if S2 is None else S2
    if S1 < 1 or S2 < 1:
        raise ValueError("s must be greater than or equal to 1.")
    if m2 > P2 // 2 + 1:
        raise ValueError(f"modes ({m2}) exceed the one-sided spectrum of a {P2}-point axis ({P2 // 2 + 1})")
    if one_d:
        rows_in = np.zeros(1, dtype=np.int64)
        rows_out = rows_in.copy()
        rows_per_w = 1
    else:
        if m1 > P1:
            raise ValueError(f"modes1={m1} exceeds the {P1} points of the x axis (the reference's corner-block "
                             "assignment fails with a shape mismatch there too)")
        low = np.arange(m1, dtype=np.int64)
        if 2 * m1 > P1:
            # overlapping corner blocks: the reference assigns the weights1 block first and the weights2 block second
            # (shared.py:330-333), so on the shared rows P1 - m1 .. m1 - 1 the LATER assignment wins and the weights1
            # rows that produced them are dead (no contribution, zero gradient): zero table rows, like any dropped bin
            low[P1 - m1:] = -1
        rows_in = np.concatenate([low, np.arange(P1 - m1, P1)]).astype(np.int64)
        if S1 == P1:
            rows_out = rows_in
        else:
            dest = _destination(twosided_source(P1, S1), P1)
            rows_out = np.where(rows_in >= 0, dest[np.maximum(rows_in, 0)], -1)
        rows_per_w = m1
    cols_in = np.arange(m2, dtype=np.int64)
    cols_out = cols_in if S2 == P2 else _destination(onesided_source(P2 // 2 + 1, S2), P2 // 2 + 1)[cols_in]
    return ModeTables("spectral", P1, P2, S1, S2, rows_in, rows_out, cols_in, cols_out,
                      1.0, 1.0 / (P1 * P2), rows_per_w)


@lru_cache(maxsize=256)
def functional_tables(P1, P2, m1, m2, one_d=False):
    """LinearFunctionals1d (weights (Ci,Co,m+1)) / LinearFunctionals2d (weights (Ci,Co,2*m1,m2+1))."""
    cols_in = onesided_source(P2 // 2 + 1, 2 * m2)
    if one_d:
        rows_in = np.zeros(1, dtype=np.int64)
        cc = np.full((1, m2 + 1), 2.0)
        cc[0, 0] = 1.0                                                   # shared.py:254
    else:
        rows_in = twosided_source(P1, 2 * m1)
        cc = np.full((2 * m1, m2 + 1), 2.0)                              # shared.py:380-381
        cc[m1:, 0] = 0.0
        cc[0, 0] = 1.0
    return ModeTables("functional", P1, P2, P1, P2, rows_in, None, cols_in, None,
                      1.0 / (P1 * P2), 0.0, len(rows_in), cc=cc)


@lru_cache(maxsize=256)
def decoder_tables(S1, S2, m1, m2, one_d=False):
    """LinearDecoder1d / LinearDecoder2d: weight row r / column l -> bin of the (S1, S2) grid."""
    if S1 < 1 or S2 < 1:
        raise ValueError("s must be greater than or equal to 1.")
    cols_out = _destination(onesided_source(m2 + 1, S2), m2 + 1)
    if one_d:
        rows_out = np.zeros(1, dtype=np.int64)
    else:
        rows_out = _destination(twosided_source(2 * m1, S1), 2 * m1)
    return ModeTables("decoder", S1, S2, S1, S2, None, rows_out, None, cols_out,
                      0.0, 1.0 / (S1 * S2), len(rows_out))


class ResampleTables:
    """``projector1d`` / ``projector2d`` (shared.py:121-128, 162-169: rfft -> resize_rfft[2] -> irfft) as real
    resampling matrices, obtained by pushing the identity through the reference's own index rules in
    float64.  ``ty`` (S2 x P2) acts along the last axis (one-sided resize), ``tx`` (S1 x P1) along the first
    axis of a 2-D field (two-sided resize).

    The two-sided rule does not always keep Hermitian-symmetric bins (an even target keeps the Nyquist bin
    on the negative side only, shared.py:104-108; an odd-length spectrum is split at N // 2 when zero
    padding, shared.py:94-96).  The x-axis operator is then complex, A = Ar + i Ai, and with Y = rfft_y(x)
        out = Ar (x R^T) + Ai (x H^T),   R = irfft_y(resized Y),  H = irfft_y(i * resized Y).
    ``separable`` False selects that two-term form: ``ty`` is then [R; H] (2 S2 x P2) and ``tx`` holds
    (Ar, Ai) interleaved along its columns (S1 x 2 P1), matching the (x, re/im) order of the intermediate."""

    def __init__(self, P1, P2, S1, S2):
        self.P1, self.P2, self.S1, self.S2 = P1, P2, S1, S2
        spec = np.fft.rfft(np.eye(P2), axis=-1) / P2
        src = onesided_source(spec.shape[-1], S2)
        res = np.where(src >= 0, spec[:, np.clip(src, 0, None)], 0.0)
        ty = (np.fft.irfft(res, n=S2, axis=-1) * S2).T                            # [y'][y]
        self.separable = True
        tx = None
        if P1 is not None:
            spec1 = np.fft.fft(np.eye(P1), axis=-1) / P1
            src1 = twosided_source(P1, S1)
            res1 = np.where(src1 >= 0, spec1[:, np.clip(src1, 0, None)], 0.0)
            txc = (np.fft.ifft(res1, n=S1, axis=-1) * S1).T                       # [x'][x], complex in general
            self.separable = bool(np.abs(txc.imag).max() <= 1e-12 * max(1.0, np.abs(txc.real).max()))
            if self.separable:
                tx = txc.real
            else:
                hy = (np.fft.irfft(1j * res, n=S2, axis=-1) * S2).T
                ty = np.concatenate([ty, hy], axis=0)                             # [(ri, y')][y]
                tx = np.stack([txc.real, txc.imag], axis=-1).reshape(S1, 2 * P1)  # [x'][(x, ri)]
        host = {"ty": ty, "tyT": ty.T}
        if tx is not None:
            host["tx"], host["txT"] = tx, tx.T
        self.host = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in host.items()}
        self._dev = {}

    def device(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = {k: torch.from_numpy(v).to(device) for k, v in self.host.items()}
        return self._dev[key]

    def apply_host(self, x):
        """float64 evaluation on a (..., P1, P2) (or (..., P2)) torch tensor: the definition the CUDA path is
        tested against besides the golden vectors."""
        ty = torch.from_numpy(self.host["ty"]).to(x.dtype)
        if self.P1 is None:
            return torch.einsum("dy,...y->...d", ty, x)
        tx = torch.from_numpy(self.host["tx"]).to(x.dtype)
        if self.separable:
            return torch.einsum("ax,...xy,dy->...ad", tx, x, ty)
        t = torch.einsum("rdy,...xy->...xrd", ty.view(2, self.S2, self.P2), x)
        return torch.einsum("axr,...xrd->...ad", tx.view(self.S1, self.P1, 2), t)


@lru_cache(maxsize=64)
def resample_tables(P1, P2, S1, S2):
    """P1 = S1 = None for the 1-D projector."""
    return ResampleTables(P1, P2, S1, S2)
