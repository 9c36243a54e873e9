"""Complex-aware Adam with the reference's update rule (reference: /root/reference/util/Adam.py) as
ONE fused multi-tensor CUDA launch per <= 48 parameter tensors (the reference issues ~9 elementwise
launches per tensor from a Python loop).

Same constructor, ``param_groups`` / ``state`` layout (``step``, ``exp_avg``, ``exp_avg_sq``
[, ``max_exp_avg_sq``]; complex parameters keep complex64 moments, Adam.py:130-138) and the same
semantics: L2 weight decay added to the gradient, ONE shared |g|^2 second moment per complex
element (Adam.py:40), parameters without a gradient skipped (Adam.py:122).
"""
import ctypes as C

import torch
from torch.optim.optimizer import Optimizer

from .parallel import _dense
from ._lib import FnmError, check, lib


def _as_float_view(t):
    return torch.view_as_real(t) if t.is_complex() else t


class Adam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False,
                 capturable=False):
        """``capturable=True`` (extension; the reference has no such flag) keeps the step count in a device
        counter that the kernel reads, so ``step()`` can be recorded in a CUDA graph and replayed."""
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        if not 0.0 <= weight_decay:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))
        self.capturable = bool(capturable)
        self._step_dev = {}

    def __setstate__(self, state):
        super().__setstate__(state)
        for group in self.param_groups:
            group.setdefault("amsgrad", False)
        self.__dict__.setdefault("capturable", False)
        self._step_dev = {}                                    # rebuilt from state["step"] at the next step()
        self._relayout_state()

    # ---- checkpoint compatibility ---------------------------------------------------------------------------------
    def _relayout_state(self):
        """Moments loaded from a checkpoint (the reference's util.Adam, or any contiguous save) arrive in the
        reference's memory order; the drop-in spectral weights are stored mode-major and the fused update runs
        element by element over STORAGE, so every state tensor is brought into its parameter's memory order."""
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state.get(p)
                if not st:
                    continue
                for key in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
                    t = st.get(key)
                    if isinstance(t, torch.Tensor) and (t.stride() != p.stride() or t.dtype != p.dtype or t.device != p.device):
                        st[key] = torch.empty_like(p, memory_format=torch.preserve_format).copy_(t)
                if isinstance(st.get("step"), torch.Tensor):    # torch.optim checkpoints keep a tensor step
                    st["step"] = int(st["step"].item())

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._step_dev = {}
        self._relayout_state()

    def sync_steps(self):
        """Capturable mode: CUDA-graph replays advance only the device counter.  ``GraphedTrainStep`` reports every
        replay (``graph_replayed``); this folds the reported count into ``state[p]["step"]`` so that checkpoints and a
        later eager step see the true count -- host bookkeeping only, no device synchronisation."""
        n = self._step_dev.get("pending", 0)
        if not n:
            return
        self._step_dev["pending"] = 0
        for p in self._step_dev.get("params", ()):
            if self.state.get(p):
                self.state[p]["step"] += n
        if "host" in self._step_dev:
            self._step_dev["host"] += n

    def graph_captured(self):
        """Called once after ``step()`` was RECORDED in a CUDA graph: recording bumped the host-side counts although no
        kernel ran."""
        for p in self._step_dev.get("params", ()):
            if self.state.get(p):
                self.state[p]["step"] -= 1
        if "host" in self._step_dev:
            self._step_dev["host"] -= 1

    def graph_replayed(self, n=1):
        self._step_dev["pending"] = self._step_dev.get("pending", 0) + n

    def reset_step_counter(self):
        """Re-derive the device counter from ``state[p]["step"]`` (after the states were edited or restored)."""
        ctr = self._step_dev.get("ctr")
        if ctr is None:
            return
        self._step_dev["pending"] = 0
        steps = {int(self.state[p]["step"]) for p in self._step_dev.get("params", ()) if self.state.get(p)}
        if len(steps) > 1:
            raise FnmError(f"Adam(capturable=True): parameters disagree on the step count ({sorted(steps)})")
        n = steps.pop() if steps else 0
        ctr.fill_(n)                                            # in place: a captured graph keeps reading this tensor
        self._step_dev["host"] = n

    def state_dict(self):
        self.sync_steps()
        return super().state_dict()

    supports_partial_step = True

    @torch.no_grad()
    def step(self, closure=None, params=None, continued=False):
        """``params`` (extension): update only these parameters -- one optimizer step may then be spread over several
        calls with disjoint subsets (``continued=True`` on every call but the first: they share the step's device counter
        increment).  The data-parallel train step uses it to update each group as soon as its gradients have been exchanged."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        only = None if params is None else {id(p) for p in params}
        if not continued:
            self._ctr_bumped = False
        if self.capturable and self._step_dev.get("ctr") is not None and not torch.cuda.is_current_stream_capturing():
            self.sync_steps()                                  # graph replays since the last eager step moved the counter
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            by_step = {}
            for p in group["params"]:
                if p.grad is None or (only is not None and id(p) not in only):
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                if not p.is_cuda:
                    raise FnmError("fnm_b200.Adam updates CUDA parameters only (no CPU fallback)")
                if group["amsgrad"] and p.is_complex():
                    raise RuntimeError("maximum not implemented for complex tensors.")   # as util/Adam.py:43
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if group["amsgrad"]:
                        state["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                by_step.setdefault(state["step"], []).append(p)
            if self.capturable and len(by_step) > 1:
                raise FnmError("Adam(capturable=True) keeps ONE device step counter: all parameters must have been "
                               f"stepped the same number of times (found steps {sorted(by_step)})")
            for step, plist in by_step.items():
                with torch.cuda.device(plist[0].device):          # the library launches on the current device
                    self._launch(plist, step, group, beta1, beta2)
        return loss

    def _launch(self, plist, step, group, beta1, beta2):
        n = len(plist)
        vp = C.c_void_p * n
        keep = []                      # float views must outlive the launch call

        def table(tensors):
            # the update is elementwise over STORAGE: parameter, gradient and state must be dense and share one memory
            # order (the spectral weights are stored mode-major, their state is created with preserve_format)
            views = []
            for t, p in zip(tensors, plist):
                if t.stride() != p.stride() or not _dense(p):
                    raise FnmError("fnm_b200.Adam needs dense parameters whose gradients and state share their memory order")
                views.append(_as_float_view(t))
            keep.extend(views)
            return vp(*[v.data_ptr() for v in views])

        def grad_of(p):
            g = p.grad if p.grad.dtype == p.dtype else p.grad.to(p.dtype)
            if g.stride() != p.stride():                       # e.g. a contiguous gradient for a mode-major weight
                g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)
            return g

        p_tab = table([p.data for p in plist])
        g_tab = table([grad_of(p) for p in plist])
        m_tab = table([self.state[p]["exp_avg"] for p in plist])
        v_tab = table([self.state[p]["exp_avg_sq"] for p in plist])
        x_tab = table([self.state[p]["max_exp_avg_sq"] for p in plist]) if group["amsgrad"] else None
        numel = (C.c_int64 * n)(*[p.numel() * (2 if p.is_complex() else 1) for p in plist])
        cplx = (C.c_uint8 * n)(*[1 if p.is_complex() else 0 for p in plist])
        stream = C.c_void_p(torch.cuda.current_stream(plist[0].device).cuda_stream)
        step_dev = None
        if self.capturable:
            # one device counter shared by all parameters (they are stepped together); bumped once per step()
            ctr = self._step_dev.get("ctr")
            if ctr is None:
                ctr = torch.full((1,), step - 1, dtype=torch.int32, device=plist[0].device)
                self._step_dev["ctr"] = ctr
                self._step_dev["params"] = []
                self._step_dev["host"] = step - 1
            if not self._ctr_bumped and self._step_dev["host"] != step - 1:
                raise FnmError("Adam(capturable=True): parameter groups disagree on the step count "
                               f"({self._step_dev['host'] + 1} vs {step})")
            known = self._step_dev["params"]
            known.extend(p for p in plist if all(p is not q for q in known))
            if not self._ctr_bumped:
                check(lib().fnm_counter_inc(C.c_void_p(ctr.data_ptr()), stream), "fnm_counter_inc")
                self._ctr_bumped = True
                self._step_dev["host"] = step
            step_dev = C.c_void_p(ctr.data_ptr())
        check(lib().fnm_adam_multi_tensor(n, p_tab, g_tab, m_tab, v_tab, x_tab, numel, cplx,
                                          float(group["lr"]), float(beta1), float(beta2), float(group["eps"]),
                                          float(group["weight_decay"]), int(bool(group["amsgrad"])), int(step), step_dev,
                                          stream), "fnm_adam_multi_tensor")
