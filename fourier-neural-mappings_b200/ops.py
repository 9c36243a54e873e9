"""torch.autograd.Function wrappers over the C ABI (hand-written forward AND backward kernels).

Tensors here are channels-last fp32: ``x[b, x, y, c]`` (1-D problems: ``x[b, 1, y, c]``).
Everything is allocated with ``torch.empty`` (caching allocator) and handed to the library as raw
device pointers together with the current CUDA stream; the library never allocates or syncs.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import FNM_MODE_FP32, FNM_MODE_TF32, check, lib

_MODE = FNM_MODE_FP32


def set_compute_mode(mode):
    """'fp32' (default; 1e-5 parity with the reference) or 'tf32' (single-pass tensor cores,
    looser stated tolerance)."""
    global _MODE
    if mode not in ("fp32", "tf32"):
        raise ValueError("mode must be 'fp32' or 'tf32'")
    _MODE = FNM_MODE_FP32 if mode == "fp32" else FNM_MODE_TF32


def get_compute_mode():
    return "fp32" if _MODE == FNM_MODE_FP32 else "tf32"


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_tensor_device(fn):
    """Run an autograd Function's forward / backward with the CUDA device of its first CUDA tensor argument current:
    the library launches on the current device and `_stream()` is that device's current stream, so a model that
    lives on cuda:1 while cuda:0 is current must switch first (one process per GPU never notices)."""
    import functools

    @functools.wraps(fn)
    def wrapper(ctx, *args):
        dev = next((a.device for a in args if isinstance(a, torch.Tensor) and a.is_cuda), None)
        if dev is None:
            dev = next((t.device for t in getattr(ctx, "saved_tensors", ()) if isinstance(t, torch.Tensor) and t.is_cuda), None)
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(ctx, *args)
        with torch.cuda.device(dev):
            return fn(ctx, *args)
    return wrapper


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.FnmError("the Fourier-layer hot path runs on CUDA only (no CPU fallback): got a "
                                f"{t.device} tensor")


def _f32c(t):
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32, got {t.dtype}")
    return t.contiguous()


def _cplx(t):
    """complex64 parameter -> contiguous float view of interleaved (re, im)."""
    if t.dtype != torch.complex64:
        raise TypeError(f"expected complex64 weights, got {t.dtype}")
    return torch.view_as_real(t.contiguous())


def _grad_sink(w):
    """The slot a FlatGradients(direct_spectral=True) registered for this weight, or None.  Called by the
    wrappers at the bottom of this file BEFORE Function.apply (inside forward grad mode is always off), with
    the module's Parameter object itself."""
    sink = getattr(w, "_fnm_grad_sink", None) if w is not None else None
    if sink is None or not (torch.is_grad_enabled() and w.requires_grad):
        return None
    if w._fnm_sink_uses:
        raise RuntimeError("direct gradient sinks need every weight to enter one autograd Function per "
                           "step (and FlatGradients.zero_() before each backward)")
    w._fnm_sink_uses = 1
    return sink


def _real_sink(p):
    """Sink of a real (conv / linear) parameter as a tensor the kernels can write densely, or None."""
    sink = _grad_sink(p)
    if sink is None:
        return None
    if not sink.is_contiguous():                               # cannot happen for the flat buffer's views of contiguous parameters
        p._fnm_sink_uses = 0
        return None
    return sink


def _is_mode_major(t):
    """(in, out, modes...) tensor whose memory order is (modes..., in, out), dense."""
    n = t.dim()
    return n >= 3 and t.permute(tuple(range(2, n)) + (0, 1)).is_contiguous()


def _spectral_weights(w0, w1):
    """-> (float view of w0, of w1 or None, FNM_W_* layout).  Both tensors must share the layout; anything that is
    neither the reference order nor mode-major is copied into the reference order."""
    if w0.dtype != torch.complex64 or (w1 is not None and w1.dtype != torch.complex64):
        raise TypeError("expected complex64 weights")
    if _is_mode_major(w0) and not w0.is_contiguous() and (w1 is None or (_is_mode_major(w1) and not w1.is_contiguous())):
        return torch.view_as_real(w0), (torch.view_as_real(w1) if w1 is not None else None), _lib.FNM_W_MODE_MAJOR
    return _cplx(w0), (_cplx(w1) if w1 is not None else None), _lib.FNM_W_REFERENCE


def _end_grad_buffer(sink, w, layout):
    """-> (buffer the F2V / V2F kernels write dW into, whether that buffer IS the sink).  The buffer has the memory order
    the kernels were told the weights have (``layout``); a sink in another order gets a copy afterwards."""
    mm = layout == _lib.FNM_W_MODE_MAJOR
    if sink is not None and (sink.stride() == w.stride() if mm else sink.is_contiguous()):
        return sink, True
    return torch.empty_like(w, memory_format=torch.preserve_format if mm else torch.contiguous_format), False


_WS = {}


def _workspace(device, nbytes):
    """Per-(device, stream) scratch, grown on demand.  Stages of one call run in stream order, and
    every call on the same stream is ordered after the previous one, so one buffer suffices."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


class FourierLayerFn(torch.autograd.Function):
    """z = Conv1x1(f(zin)) + bias + SpectralConv(f(zin)), f = GELU if act_in else identity; optionally
    also returns GELU(z) (non-differentiable cache for the next layer).

    ``xact`` is an optional cached GELU(zin) produced by the previous layer's epilogue: the kernels then
    read it with no activation on load, and backward multiplies by GELU'(zin).  wc / bc may be None
    (bare SpectralConv)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, zin, xact, w0, w1, wc, bc, tables, act_in, want_act_out, sinks=(None, None), pre_is_grad=False,
                grad_out=False):
        _need_cuda(zin, xact, w0, w1, wc, bc)
        cached = bool(act_in) and xact is not None
        if pre_is_grad and not cached:
            raise ValueError("pre_is_grad: a cached GELU'(z) input needs the cached GELU(z) beside it")
        if grad_out and not want_act_out:
            raise ValueError("grad_out: GELU'(z) is written beside the GELU copy only")
        xk = _f32c(xact) if cached else _f32c(zin)
        kact = 0 if cached else int(bool(act_in))
        B, P1, P2, Ci = xk.shape
        Co = w0.shape[1]
        if (P1, P2) != (tables.P1, tables.P2) or w0.shape[0] != Ci:
            raise ValueError(f"input {tuple(xk.shape)} does not match the plan / weights")
        plan = tables.c_plan(xk.device, B, Ci, Co, _MODE)
        w0r, w1r, plan.w_layout = _spectral_weights(w0, w1)
        wcc = _f32c(wc.reshape(Co, Ci)) if wc is not None else None
        bcc = _f32c(bc) if bc is not None else None
        z = torch.empty((B, tables.S1, tables.S2, Co), dtype=torch.float32, device=xk.device)
        xo = torch.empty_like(z) if want_act_out else None
        xh = torch.empty((tables.R, tables.NY, 2, B, Ci), dtype=torch.float32, device=xk.device)
        L = lib()
        nbytes = L.fnm_fourier_layer_workspace_bytes(C.byref(plan))
        ws = _workspace(xk.device, nbytes)
        # training: the mode-major weight copy made here is kept for backward (one permute per layer and step)
        wsh = None
        if any(ctx.needs_input_grad[:6]):
            nsh = L.fnm_fourier_layer_shadow_bytes(C.byref(plan))
            if nsh:
                wsh = torch.empty(nsh, dtype=torch.uint8, device=xk.device)
                plan.w_shadow = wsh.data_ptr()
        ctx.wsh = wsh
        if grad_out:
            plan.flags = _lib.FNM_PLAN_ZOUT_IS_GELU_GRAD
        check(L.fnm_fourier_layer_fwd(C.byref(plan), _ptr(xk), kact, _ptr(w0r), _ptr(w1r), _ptr(wcc), _ptr(bcc),
                                      _ptr(z), _ptr(xo), _ptr(xh), _ptr(ws), ws.numel(), _stream()),
              "fnm_fourier_layer_fwd")
        ctx.save_for_backward(xk, _f32c(zin) if cached else None, w0, w1, wc, xh)
        ctx.tables, ctx.kact, ctx.has_bias = tables, kact, bc is not None
        ctx.pre_is_grad = bool(pre_is_grad)
        ctx.sinks = tuple(sinks) + (None,) * (4 - len(sinks))
        ctx.w_version = (w0._version, w1._version if w1 is not None else 0)
        # the GELU copy carries no gradient: without this autograd would MATERIALISE a zero tensor of its size
        # (one activation-sized fill per layer) just to hand it to backward
        ctx.set_materialize_grads(False)
        if xo is not None:
            ctx.mark_non_differentiable(xo)
        return z, xo

    @staticmethod
    @_on_tensor_device
    def backward(ctx, gz, _gxo=None):
        xk, zpre, w0, w1, wc, xh = ctx.saved_tensors
        tables = ctx.tables
        if gz is None:                                          # z did not reach the loss
            for sink in ctx.sinks:
                if sink is not None:
                    sink.zero_()                                # a sink must not keep the previous step's gradient
            return (None,) * 12
        gz = _f32c(gz)
        B, P1, P2, Ci = xk.shape
        Co = w0.shape[1]
        plan = tables.c_plan(xk.device, B, Ci, Co, _MODE)
        if ctx.wsh is not None and (w0._version, w1._version if w1 is not None else 0) == ctx.w_version:      # weights not modified in place since forward
            plan.w_shadow = ctx.wsh.data_ptr()
        if ctx.pre_is_grad:
            plan.flags = _lib.FNM_PLAN_PRE_IS_GELU_GRAD
        need_x, _, need_w0, need_w1, need_wc, need_bc = ctx.needs_input_grad[:6]
        g_in = torch.empty_like(xk) if need_x else None
        want_w = need_w0 or (w1 is not None and need_w1)
        s0, s1, swc, sbc = ctx.sinks
        w0r, w1r, plan.w_layout = _spectral_weights(w0, w1)
        # gradients in the memory order the kernels were told the weights have; a sink in another order gets a copy
        mm = plan.w_layout == _lib.FNM_W_MODE_MAJOR

        def grad_buffer(sink, w):
            if w is None or not want_w:
                return None
            if sink is not None and (sink.stride() == w.stride() if mm else sink.is_contiguous()):
                return sink
            return torch.empty_like(w, memory_format=torch.preserve_format if mm else torch.contiguous_format)

        dw0, dw1 = grad_buffer(s0, w0), grad_buffer(s1, w1)
        # conv weight / bias: straight into their sinks when the flat gradient buffer registered them
        dwc = (swc if swc is not None else torch.empty_like(wc)) if (wc is not None and need_wc) else None
        dbc = (sbc if sbc is not None else torch.empty((Co,), dtype=torch.float32, device=xk.device)) \
            if (ctx.has_bias and need_bc) else None
        wcc = _f32c(wc.reshape(Co, Ci)) if wc is not None else None
        L = lib()
        nbytes = L.fnm_fourier_layer_workspace_bytes(C.byref(plan))
        ws = _workspace(xk.device, nbytes)
        def run(phase):
            flags0 = plan.flags
            plan.flags = flags0 | phase
            check(L.fnm_fourier_layer_bwd(
                C.byref(plan), _ptr(gz), _ptr(xk), ctx.kact, _ptr(zpre), _ptr(w0r), _ptr(w1r), _ptr(wcc), _ptr(xh),
                _ptr(g_in), _ptr(torch.view_as_real(dw0)) if dw0 is not None else None,
                _ptr(torch.view_as_real(dw1)) if dw1 is not None else None, _ptr(dwc), _ptr(dbc),
                _ptr(ws), ws.numel(), _stream()), "fnm_fourier_layer_bwd")
            plan.flags = flags0

        def sinks_done():
            # a sink already holds the gradient: autograd gets None and runs no AccumulateGrad for that weight
            for sink, dw in ((s0, dw0), (s1, dw1)):
                if sink is not None and dw is not None and dw is not sink:
                    sink.copy_(dw)
                ready = getattr(sink, "_fnm_on_ready", None) if sink is not None else None
                if ready is not None:
                    ready()                                         # FlatGradients(overlap=True): this slot is ready for its exchange
            for flush in {id(f): f for f in (getattr(s, "_fnm_flush", None) for s in (s0, s1) if s is not None) if f is not None}.values():
                flush()                                             # ... and the layer's slots go out as one all-reduce

        overlapped = any(getattr(s, "_fnm_on_ready", None) is not None for s in (s0, s1) if s is not None)
        if overlapped and need_x and want_w:
            # weight gradients first, their exchange started, THEN the input gradient (mix_bwd_x, x-synthesis, fused pass):
            # the all-reduce of this layer's slots runs under those kernels instead of waiting for the end of the layer
            run(_lib.FNM_PLAN_BWD_WEIGHTS_ONLY)
            sinks_done()
            run(_lib.FNM_PLAN_BWD_INPUT_ONLY)
        else:
            run(0)
            sinks_done()
        return (g_in, None, None if s0 is not None else dw0, None if s1 is not None else dw1,
                None if swc is not None else dwc, None if sbc is not None else dbc, None, None, None, None, None, None)


class LinearFunctionalFn(torch.autograd.Function):
    """out[b, o] = F2V functional of f(zin) (LinearFunctionals1d/2d); ``xact`` = optional cached GELU(zin)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, zin, xact, w, tables, act_in, sink=None, pre_is_grad=False):
        _need_cuda(zin, xact, w)
        cached = bool(act_in) and xact is not None
        if pre_is_grad and not cached:
            raise ValueError("pre_is_grad: a cached GELU'(z) input needs the cached GELU(z) beside it")
        ctx.pre_is_grad = bool(pre_is_grad)
        xk = _f32c(xact) if cached else _f32c(zin)
        kact = 0 if cached else int(bool(act_in))
        B, P1, P2, Ci = xk.shape
        Co = w.shape[1]
        if (P1, P2) != (tables.P1, tables.P2) or w.shape[0] != Ci:
            raise ValueError(f"input {tuple(xk.shape)} does not match the plan / weights")
        plan = tables.c_plan(xk.device, B, Ci, Co, _MODE)
        cc = tables.device(xk.device)["cc"]
        out = torch.empty((B, Co), dtype=torch.float32, device=xk.device)
        xh = torch.empty((tables.R, tables.NY, 2, B, Ci), dtype=torch.float32, device=xk.device)
        L = lib()
        ws = _workspace(xk.device, L.fnm_lfunc_workspace_bytes(C.byref(plan)))
        wr, _, plan.w_layout = _spectral_weights(w, None)
        check(L.fnm_lfunc_fwd(C.byref(plan), _ptr(xk), kact, _ptr(wr), _ptr(cc), _ptr(out), _ptr(xh),
                              _ptr(ws), ws.numel(), _stream()), "fnm_lfunc_fwd")
        ctx.save_for_backward(xk, _f32c(zin) if cached else None, w, xh)
        ctx.tables, ctx.kact = tables, kact
        ctx.sink = sink
        return out

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g):
        xk, zpre, w, xh = ctx.saved_tensors
        tables = ctx.tables
        g = _f32c(g)
        B, _, _, Ci = xk.shape
        Co = w.shape[1]
        plan = tables.c_plan(xk.device, B, Ci, Co, _MODE)
        cc = tables.device(xk.device)["cc"]
        g_in = torch.empty_like(xk) if ctx.needs_input_grad[0] else None
        wr, _, plan.w_layout = _spectral_weights(w, None)
        if ctx.pre_is_grad:
            plan.flags = _lib.FNM_PLAN_PRE_IS_GELU_GRAD
        dw, sunk = _end_grad_buffer(ctx.sink, w, plan.w_layout) if ctx.needs_input_grad[2] else (None, False)
        L = lib()
        ws = _workspace(xk.device, L.fnm_lfunc_workspace_bytes(C.byref(plan)))
        check(L.fnm_lfunc_bwd(C.byref(plan), _ptr(g), _ptr(xk), ctx.kact, _ptr(zpre), _ptr(wr), _ptr(cc), _ptr(xh),
                              _ptr(g_in), _ptr(torch.view_as_real(dw)) if dw is not None else None,
                              _ptr(ws), ws.numel(), _stream()), "fnm_lfunc_bwd")
        if ctx.sink is not None and dw is not None and not sunk:
            ctx.sink.copy_(dw)
        return g_in, None, None if ctx.sink is not None else dw, None, None, None, None


class LinearDecoderFn(torch.autograd.Function):
    """y[b, x, y, o] = V2F decoder of the vector v[b, i] (LinearDecoder1d/2d)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, v, w, tables, sink=None):
        _need_cuda(v, w)
        v = _f32c(v)
        B, Ci = v.shape
        Co = w.shape[1]
        if w.shape[0] != Ci:
            raise ValueError("vector width does not match the weights")
        plan = tables.c_plan(v.device, B, Ci, Co, _MODE)
        y = torch.empty((B, tables.S1, tables.S2, Co), dtype=torch.float32, device=v.device)
        L = lib()
        ws = _workspace(v.device, L.fnm_ldec_workspace_bytes(C.byref(plan)))
        wr, _, plan.w_layout = _spectral_weights(w, None)
        check(L.fnm_ldec_fwd(C.byref(plan), _ptr(v), _ptr(wr), _ptr(y), _ptr(ws), ws.numel(), _stream()),
              "fnm_ldec_fwd")
        ctx.save_for_backward(v, w)
        ctx.tables = tables
        ctx.sink = sink
        return y

    @staticmethod
    @_on_tensor_device
    def backward(ctx, gy):
        v, w = ctx.saved_tensors
        tables = ctx.tables
        gy = _f32c(gy)
        B, Ci = v.shape
        Co = w.shape[1]
        plan = tables.c_plan(v.device, B, Ci, Co, _MODE)
        gv = torch.empty_like(v) if ctx.needs_input_grad[0] else None
        wr, _, plan.w_layout = _spectral_weights(w, None)
        dw, sunk = _end_grad_buffer(ctx.sink, w, plan.w_layout) if ctx.needs_input_grad[1] else (None, False)
        L = lib()
        ws = _workspace(v.device, L.fnm_ldec_workspace_bytes(C.byref(plan)))
        check(L.fnm_ldec_bwd(C.byref(plan), _ptr(gy), _ptr(v), _ptr(wr), _ptr(gv),
                             _ptr(torch.view_as_real(dw)) if dw is not None else None,
                             _ptr(ws), ws.numel(), _stream()), "fnm_ldec_bwd")
        if ctx.sink is not None and dw is not None and not sunk:
            ctx.sink.copy_(dw)
        return gv, None if ctx.sink is not None else dw, None, None


class GeluFn(torch.autograd.Function):
    """Stand-alone exact GELU for a trunk that ends with an activation."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, z):
        _need_cuda(z)
        z = _f32c(z)
        y = torch.empty_like(z)
        check(lib().fnm_gelu_fwd(_ptr(z), _ptr(y), z.numel(), _stream()), "fnm_gelu_fwd")
        ctx.save_for_backward(z)
        return y

    @staticmethod
    @_on_tensor_device
    def backward(ctx, gy):
        (z,) = ctx.saved_tensors
        gy = _f32c(gy)
        gz = torch.empty_like(z)
        check(lib().fnm_gelu_bwd(_ptr(gy), _ptr(z), _ptr(gz), z.numel(), _stream()), "fnm_gelu_bwd")
        return gz


class LiftFn(torch.autograd.Function):
    """Lifting: fc0 on [data channels | grid coordinates] + zero padding -> channels-last (B, P1, P2, C)
    (reference: func_to_func2d.py:79-86).  ``x`` is (B, Din, N1, N2) (1-D: N1 = 1); it gets no gradient."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, x, w, b, P1, P2, ngrid, sinks=(None, None)):
        _need_cuda(x, w, b)
        ctx.sinks = sinks
        x, w, b = _f32c(x), _f32c(w), _f32c(b)
        B, Din, N1, N2 = x.shape
        Cw = w.shape[0]
        out = torch.empty((B, P1, P2, Cw), dtype=torch.float32, device=x.device)
        check(lib().fnm_lift_fwd(_ptr(x), _ptr(w), _ptr(b), _ptr(out), B, Din, N1, N2, P1, P2, Cw, ngrid, _stream()),
              "fnm_lift_fwd")
        ctx.save_for_backward(x)
        ctx.dims = (B, Din, N1, N2, P1, P2, Cw, ngrid)
        return out

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        B, Din, N1, N2, P1, P2, Cw, ngrid = ctx.dims
        g = _f32c(g)
        J = Din + ngrid
        sw, sb = ctx.sinks
        dw = sw if sw is not None else torch.empty((Cw, J), dtype=torch.float32, device=g.device)
        db = sb if sb is not None else torch.empty((Cw,), dtype=torch.float32, device=g.device)
        L = lib()
        ws = _workspace(g.device, L.fnm_lift_bwd_workspace_bytes(Cw, J))
        check(L.fnm_lift_bwd(_ptr(g), _ptr(x), _ptr(dw), _ptr(db), B, Din, N1, N2, P1, P2, Cw, ngrid,
                             _ptr(ws), ws.numel(), _stream()), "fnm_lift_bwd")
        return None, None if sw is not None else dw, None if sb is not None else db, None, None, None, None


class PointwiseConvFn(torch.autograd.Function):
    """h[..., o] = sum_i x[..., i] w[o, i] + b[o] on a channels-last (B, P1, P2, Ci) tensor: the fused
    pointwise kernel without a spectral term (forward and input gradient) + conv_wgrad (dW, db).

    With ``zpre``: ``x`` is the cached GELU(zpre) a fused layer's epilogue wrote (it carries no gradient) and the
    input gradient is returned for ``zpre``, multiplied by GELU'(zpre) in the same kernel's epilogue."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, x, w, b, zpre=None, zpre_is_grad=False, sinks=(None, None)):
        _need_cuda(x, w, b, zpre)
        ctx.zpre_is_grad = bool(zpre_is_grad) and zpre is not None
        ctx.sinks = sinks
        x, w = _f32c(x), _f32c(w)
        B, P1, P2, Ci = x.shape
        Co = w.shape[0]
        out = torch.empty((B, P1, P2, Co), dtype=torch.float32, device=x.device)
        check(lib().fnm_pointwise_ysyn(_ptr(x), 0, _ptr(w), 0, _ptr(_f32c(b)) if b is not None else None, None, None, None,
                                       _ptr(out), None, B * P1, P2, 0, Ci, Co, _MODE, _stream()), "fnm_pointwise_ysyn")
        ctx.save_for_backward(x, w, _f32c(zpre) if zpre is not None else None)
        ctx.has_bias = b is not None
        return out

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g):
        x, w, zpre = ctx.saved_tensors
        g = _f32c(g)
        B, P1, P2, Ci = x.shape
        Co = w.shape[0]
        L = lib()
        dx = dw = db = None
        if ctx.needs_input_grad[3] if zpre is not None else ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            check(L.fnm_pointwise_ysyn_ex(_ptr(g), 0, _ptr(w), 1, None, None, None, _ptr(zpre), _ptr(dx), None, B * P1, P2, 0,
                                          Co, Ci, _MODE, _lib.FNM_PW_MUL_IS_GRAD if ctx.zpre_is_grad else 0, None, _stream()),
                  "fnm_pointwise_ysyn_ex")
        sw, sb = ctx.sinks
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            npos = B * P1 * P2
            dw = sw if sw is not None else torch.empty_like(w)
            db = sb if sb is not None else torch.empty((Co,), dtype=torch.float32, device=g.device)
            ws = _workspace(g.device, L.fnm_conv_wgrad_workspace_bytes(npos, Ci, Co))
            check(L.fnm_conv_wgrad(_ptr(g), _ptr(x), 0, _ptr(dw), _ptr(db), npos, Ci, Co, _ptr(ws), ws.numel(), _MODE,
                                   _stream()), "fnm_conv_wgrad")
            if not ctx.has_bias:
                db = None
        dw, db = (None if sw is not None else dw), (None if sb is not None else db)
        if zpre is not None:
            return None, dw, db, dx, None, None
        return dx, dw, db, None, None, None


class CachedGeluFn(torch.autograd.Function):
    """GELU(z) whose value already exists (``x``, written by the producing layer's epilogue): no forward kernel; the
    backward is the stand-alone g * GELU'(z)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, z, x):
        _need_cuda(z, x)
        ctx.save_for_backward(_f32c(z))
        return x.view_as(x)

    @staticmethod
    @_on_tensor_device
    def backward(ctx, gy):
        (z,) = ctx.saved_tensors
        gy = _f32c(gy)
        gz = torch.empty_like(z)
        check(lib().fnm_gelu_bwd(_ptr(gy), _ptr(z), _ptr(gz), z.numel(), _stream()), "fnm_gelu_bwd")
        return gz, None


class ResampleFn(torch.autograd.Function):
    """out[g, m, j, c] = sum_k tab[m, k] x[g, k, j, c]: one axis of the reference's Fourier projector
    (shared.py:121-128, 162-169) as a table contraction; backward is the same kernel with the transposed
    table.  ``x`` is viewed as (G, K, inner, C)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, x, tab, tabT, G, K, M, inner):
        _need_cuda(x, tab)
        x = _f32c(x)
        Cc = x.shape[-1]
        out = torch.empty((G, M, inner, Cc), dtype=torch.float32, device=x.device)
        check(lib().fnm_resample(_ptr(x), _ptr(tab), _ptr(out), G, K, M, inner, Cc, _MODE, _stream()), "fnm_resample")
        ctx.tabT, ctx.dims, ctx.in_shape = tabT, (G, K, M, inner, Cc), x.shape
        return out

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g):
        G, K, M, inner, Cc = ctx.dims
        g = _f32c(g)
        dx = torch.empty((G, K, inner, Cc), dtype=torch.float32, device=g.device)
        check(lib().fnm_resample(_ptr(g), _ptr(ctx.tabT), _ptr(dx), G, M, K, inner, Cc, _MODE, _stream()), "fnm_resample")
        return dx.view(ctx.in_shape), None, None, None, None, None, None


def project_resolution(x, s, one_d):
    """``projector1d`` / ``projector2d`` on a channels-last (B, P1, P2, C) tensor (P1 = 1 in 1-D): one table
    contraction per axis (two stacked ones where the reference's resize keeps a non-Hermitian bin set, see
    plan.ResampleTables)."""
    from .plan import resample_tables
    B, P1, P2, Cc = x.shape
    if one_d:
        S2 = int(s)
        d = resample_tables(None, P2, None, S2).device(x.device)
        return ResampleFn.apply(x, d["ty"], d["tyT"], B, P2, S2, 1).view(B, 1, S2, Cc)
    S1, S2 = int(s[0]), int(s[1])
    t = resample_tables(P1, P2, S1, S2)
    d = t.device(x.device)
    if t.separable:
        if S2 != P2:
            x = ResampleFn.apply(x, d["ty"], d["tyT"], B * P1, P2, S2, 1)
        if S1 != P1:
            x = ResampleFn.apply(x, d["tx"], d["txT"], B, P1, S1, S2)
        return x.view(B, S1, S2, Cc)
    x = ResampleFn.apply(x, d["ty"], d["tyT"], B * P1, P2, 2 * S2, 1)             # (B, P1, (re/im, S2), C)
    return ResampleFn.apply(x, d["tx"], d["txT"], B, 2 * P1, S1, S2).view(B, S1, S2, Cc)


class MlpHeadFn(torch.autograd.Function):
    """out[..., d] = sum_h w2[d, h] GELU(h1[..., h]) + b2[d] (the activation and second layer of the
    reference MLP, shared.py:26-45) without materialising GELU(h1)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, h1, w2, b2, sinks=(None, None)):
        _need_cuda(h1, w2, b2)
        ctx.sinks = sinks
        h1, w2 = _f32c(h1), _f32c(w2)
        H, D = h1.shape[-1], w2.shape[0]
        npos = h1.numel() // H
        out = torch.empty(h1.shape[:-1] + (D,), dtype=torch.float32, device=h1.device)
        check(lib().fnm_mlp_head_fwd(_ptr(h1), _ptr(w2), _ptr(_f32c(b2)) if b2 is not None else None, _ptr(out), npos, H, D,
                                     _stream()), "fnm_mlp_head_fwd")
        ctx.save_for_backward(h1, w2)
        ctx.has_bias = b2 is not None
        return out

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g):
        h1, w2 = ctx.saved_tensors
        g = _f32c(g)
        H, D = h1.shape[-1], w2.shape[0]
        npos = h1.numel() // H
        g_h1 = torch.empty_like(h1)
        sw, sb = ctx.sinks
        if not ctx.has_bias:
            sb = None
        dw2 = sw if sw is not None else torch.empty_like(w2)
        db2 = sb if sb is not None else torch.empty((D,), dtype=torch.float32, device=g.device)
        L = lib()
        ws = _workspace(g.device, L.fnm_mlp_head_bwd_workspace_bytes(H, D))
        check(L.fnm_mlp_head_bwd(_ptr(h1), _ptr(g), _ptr(w2), _ptr(g_h1), _ptr(dw2), _ptr(db2), npos, H, D,
                                 _ptr(ws), ws.numel(), _stream()), "fnm_mlp_head_bwd")
        return g_h1, None if sw is not None else dw2, (db2 if (ctx.has_bias and sb is None) else None), None


class WeightedGeluSumFn(torch.autograd.Function):
    """S[b, h] = sum_{x,y} wx[x] wy[y] GELU(h1[b, x, y, h]): the field-sized part of mlpfunc0 + double trapz
    (func_to_vec2d.py:100-104); fc2 is linear and is applied to S afterwards."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, h1, wx, wy):
        _need_cuda(h1, wx, wy)
        h1, wx, wy = _f32c(h1), _f32c(wx), _f32c(wy)
        B, P1, P2, H = h1.shape
        S = torch.empty((B, H), dtype=torch.float32, device=h1.device)
        L = lib()
        ws = _workspace(h1.device, L.fnm_wsum_gelu_workspace_bytes(B, H))
        check(L.fnm_wsum_gelu_fwd(_ptr(h1), _ptr(wx), _ptr(wy), _ptr(S), B, P1, P2, H, _ptr(ws), ws.numel(), _stream()),
              "fnm_wsum_gelu_fwd")
        ctx.save_for_backward(h1, wx, wy)
        return S

    @staticmethod
    @_on_tensor_device
    def backward(ctx, gS):
        h1, wx, wy = ctx.saved_tensors
        B, P1, P2, H = h1.shape
        g = torch.empty_like(h1)
        check(lib().fnm_wsum_gelu_bwd(_ptr(h1), _ptr(wx), _ptr(wy), _ptr(_f32c(gS)), _ptr(g), B, P1, P2, H, _stream()),
              "fnm_wsum_gelu_bwd")
        return g, None, None


class LocalFunctionalFn(torch.autograd.Function):
    """S[b, h] = sum_{x,y} wx[x] wy[y] GELU(fc1(x)[b, x, y, h]) in one pass: fc1 on the tcgen05 pointwise kernel whose
    epilogue integrates (func_to_vec2d.py:100-104 without the hidden tensor).  ``zpre`` as in PointwiseConvFn: ``x`` is then
    the cached GELU(zpre) and the input gradient goes to ``zpre``.  backward recomputes fc1 from ``x``."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, x, w1, b1, wx, wy, zpre=None, zpre_is_grad=False, sinks=(None, None)):
        _need_cuda(x, w1, b1, wx, wy, zpre)
        ctx.sinks = sinks
        x, w1, b1 = _f32c(x), _f32c(w1), _f32c(b1)
        B, P1, P2, Ci = x.shape
        H = w1.shape[0]
        S = torch.empty((B, H), dtype=torch.float32, device=x.device)
        L = lib()
        ws = _workspace(x.device, L.fnm_local_functional_workspace_bytes(B, P1, P2, Ci, H))
        check(L.fnm_local_functional_fwd(_ptr(x), _ptr(w1), _ptr(b1), _ptr(wx), _ptr(wy), _ptr(S), B, P1, P2, Ci, H,
                                         _ptr(ws), ws.numel(), _MODE, _stream()), "fnm_local_functional_fwd")
        ctx.save_for_backward(x, w1, b1, wx, wy, _f32c(zpre) if zpre is not None else None)
        ctx.zpre_is_grad = bool(zpre_is_grad) and zpre is not None
        return S

    @staticmethod
    @_on_tensor_device
    def backward(ctx, gS):
        x, w1, b1, wx, wy, zpre = ctx.saved_tensors
        B, P1, P2, Ci = x.shape
        H = w1.shape[0]
        L = lib()
        g = torch.empty((B, P1, P2, H), dtype=torch.float32, device=x.device)
        check(L.fnm_local_functional_bwd(_ptr(x), _ptr(w1), _ptr(b1), _ptr(wx), _ptr(wy), _ptr(_f32c(gS)), _ptr(g), B, P1, P2,
                                         Ci, H, _MODE, _stream()), "fnm_local_functional_bwd")
        dx = dw = db = None
        if ctx.needs_input_grad[5] if zpre is not None else ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            check(L.fnm_pointwise_ysyn_ex(_ptr(g), 0, _ptr(w1), 1, None, None, None, _ptr(zpre), _ptr(dx), None, B * P1, P2, 0,
                                          H, Ci, _MODE, _lib.FNM_PW_MUL_IS_GRAD if ctx.zpre_is_grad else 0, None, _stream()),
                  "fnm_pointwise_ysyn_ex")
        sw, sb = ctx.sinks
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            npos = B * P1 * P2
            dw = sw if sw is not None else torch.empty_like(w1)
            db = sb if sb is not None else torch.empty((H,), dtype=torch.float32, device=x.device)
            ws = _workspace(x.device, L.fnm_conv_wgrad_workspace_bytes(npos, Ci, H))
            check(L.fnm_conv_wgrad(_ptr(g), _ptr(x), 0, _ptr(dw), _ptr(db), npos, Ci, H, _ptr(ws), ws.numel(), _MODE,
                                   _stream()), "fnm_conv_wgrad")
        dw, db = (None if sw is not None else dw), (None if sb is not None else db)
        if zpre is not None:
            return None, dw, db, None, None, dx, None, None
        return dx, dw, db, None, None, None, None, None


def local_functional_supported(x, w1):
    """Whether the one-pass F2V local branch serves this shape (otherwise: pointwise_conv + weighted_gelu_sum)."""
    B, P1, P2, Ci = x.shape
    return x.is_cuda and lib().fnm_local_functional_workspace_bytes(B, P1, P2, Ci, w1.shape[0]) > 0 and B <= 65535


def local_functional(x, w1, b1, wx, wy, zpre=None, zpre_is_grad=False):
    return LocalFunctionalFn.apply(x, w1, b1, wx, wy, zpre, zpre_is_grad, (_real_sink(w1), _real_sink(b1)))


class RelL2SumFn(torch.autograd.Function):
    """sum_b ||out_b - y_b||_2 / (||y_b||_2 + eps): the reference's training loss (LpLoss(size_average=False),
    util/utilities_module.py:188-204) in two launches forward and one backward."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, out, y, eps):
        _need_cuda(out, y)
        out, y = _f32c(out), _f32c(y)
        B = out.shape[0]
        n = out.numel() // B
        loss = torch.empty((), dtype=torch.float32, device=out.device)
        coef = torch.empty((B,), dtype=torch.float32, device=out.device)
        L = lib()
        ws = _workspace(out.device, L.fnm_rel_l2_workspace_bytes(B))
        check(L.fnm_rel_l2_fwd(_ptr(out), _ptr(y), _ptr(loss), _ptr(coef), B, n, float(eps), _ptr(ws), ws.numel(), _stream()),
              "fnm_rel_l2_fwd")
        ctx.save_for_backward(out, y, coef)
        return loss

    @staticmethod
    @_on_tensor_device
    def backward(ctx, gloss):
        out, y, coef = ctx.saved_tensors
        B = out.shape[0]
        n = out.numel() // B
        g = torch.empty_like(out)
        check(lib().fnm_rel_l2_bwd(_ptr(out), _ptr(y), _ptr(coef), _ptr(_f32c(gloss)), _ptr(g), B, n, _stream()),
              "fnm_rel_l2_bwd")
        return g, None, None


def rel_l2_sum_supported(out, y):
    if not (out.is_cuda and y.is_cuda and out.dtype == torch.float32 and y.dtype == torch.float32 and out.shape == y.shape):
        return False
    B = out.shape[0]
    n = out.numel() // max(B, 1)
    return 0 < B <= 65535 and n > 0 and not y.requires_grad


def rel_l2_sum(out, y, eps=1e-6):
    return RelL2SumFn.apply(out, y, eps)


def weighted_gelu_sum(h1, wx, wy):
    return WeightedGeluSumFn.apply(h1, wx, wy)


def lift(x, w, b, P1, P2, ngrid):
    return LiftFn.apply(x, w, b, P1, P2, ngrid, (_real_sink(w), _real_sink(b)))


def pointwise_conv(x, w, b, zpre=None, zpre_is_grad=False):
    return PointwiseConvFn.apply(x, w, b, zpre, zpre_is_grad, (_real_sink(w), _real_sink(b)))


def mlp_head(h1, w2, b2):
    return MlpHeadFn.apply(h1, w2, b2, (_real_sink(w2), _real_sink(b2)))


def fourier_layer(zin, w0, w1, wc, bc, tables, act_in=False, xact=None, want_act_out=False, pre_is_grad=False,
                  grad_out=False):
    """Returns (z, GELU(z) or None).  ``zin`` is the layer input (pre-activation if act_in).

    Cached GELU derivative (include/fnm_b200.h, FNM_PLAN_*): with ``grad_out`` (and want_act_out) the first result
    holds GELU'(z) INSTEAD of z -- only a fused layer called with ``pre_is_grad`` may consume it (it multiplies its
    outgoing gradient by that buffer instead of evaluating GELU' of it); the gradient it receives is still dL/dz."""
    return FourierLayerFn.apply(zin, xact, w0, w1, wc, bc, tables, act_in, want_act_out,
                                (_grad_sink(w0), _grad_sink(w1), _real_sink(wc), _real_sink(bc)), pre_is_grad, grad_out)


def linear_functional(zin, w, tables, act_in=False, xact=None, pre_is_grad=False):
    """``pre_is_grad``: ``zin`` holds GELU'(z) (ops.fourier_layer(grad_out=True)) and ``xact`` GELU(z)."""
    return LinearFunctionalFn.apply(zin, xact, w, tables, act_in, _grad_sink(w), pre_is_grad)


def linear_decoder(v, w, tables):
    return LinearDecoderFn.apply(v, w, tables, _grad_sink(w))


def gelu(z, cached=None):
    """GELU(z); ``cached`` = the value when a fused layer's epilogue already wrote it."""
    if cached is not None:
        return CachedGeluFn.apply(z, cached)
    return GeluFn.apply(z)
