"""Data parallelism for the train step: one process per GPU, batch sharded, parameters and Adam
state replicated, ONE exchange per step -- a SUM all-reduce of the weight gradients.

SUM (not DDP's mean) because the reference loss sums over the batch
(``LpLoss(size_average=False)``, /root/reference/advection_diffusion/func_to_func/train_fno.py:98,
util/utilities_module.py:195-199): summed shard gradients equal the global-batch gradient, so the
replicated Adam step is identical on every rank.  The reference itself has no distributed code
(SURVEY.md section 2.2); this layer is new.

Gradients live in ONE flat fp32 buffer (complex64 viewed as 2 x fp32); every ``param.grad`` is a
view into it, so the exchange is a single NCCL all-reduce over NVLink (gloo on CPU in the tests)
with no packing copies, and ``zero_grad`` is one memset.
"""
import torch
import torch.distributed as dist


class FlatGradients:
    """Owns the flat gradient buffer of ``params`` and keeps ``p.grad`` pointing into it.

    ``direct_spectral=True`` additionally makes the slots *gradient sinks*: the backward kernels of ops.py write
    dW straight into the slot and hand autograd no gradient, which removes the AccumulateGrad pass (for the spectral
    weights: read dW, read grad, write grad, 3 x 1.07 GB at config 5; for the conv / lifting / projection weights one
    small kernel each, which is what the launch-bound small configurations notice) and lets ``zero_`` skip the complex
    slots.  It assumes what holds for every model here: each spectral weight enters exactly one
    autograd Function per step and ``zero_()`` is called before every backward (a second use raises)."""

    def __init__(self, params, direct_spectral=False, overlap=False, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group                                     # process group of the exchange (None: the default group)
        # overlap (needs direct_spectral): each spectral weight's slot is all-reduced as soon as the backward kernel
        # that fills it has been enqueued (asynchronously, on the process group's stream), so the exchange of the deep
        # layers runs under the backward pass of the shallow ones; all_reduce() then handles the rest and joins.
        self.overlap = bool(overlap and direct_spectral)
        self._pending, self._staged = [], []
        self._reserve_sms, self._limit_prev = 0, None
        if not self.params:
            raise ValueError("no trainable parameters")
        self.direct = bool(direct_spectral)
        if self.direct:                                        # sinks last: zero_ is one memset of the head
            self.params.sort(key=lambda p: p.is_complex())
        dev = self.params[0].device
        self.offsets, total = [], 0
        for p in self.params:
            if p.device != dev:
                raise ValueError("all parameters must live on one device")
            n = p.numel() * (2 if p.is_complex() else 1)
            self.offsets.append((total, n))
            total += (n + 63) // 64 * 64                     # 256-byte aligned slots
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        firsts = [off for p, (off, _) in zip(self.params, self.offsets) if p.is_complex()]
        self.head = firsts[0] if (self.direct and firsts) else total      # floats that zero_ clears
        self.attach()

    def attach(self):
        for p, (off, n) in zip(self.params, self.offsets):
            chunk = self.flat[off:off + n]
            flat = torch.view_as_complex(chunk.view(-1, 2)) if p.is_complex() else chunk
            # same memory order as the parameter (the spectral weights are stored mode-major): Adam and the dW kernels
            # address parameter and gradient storage element by element
            view = torch.as_strided(flat, p.shape, p.stride()) if _dense(p) else flat.view(p.shape)
            p.grad = view
            if self.direct:
                # every parameter's slot is a sink: the backward kernels that compute a gradient write it there and autograd
                # gets None (no AccumulateGrad kernel).  Real parameters that reach their gradient through torch ops instead
                # simply accumulate into the (zeroed) slot as usual.
                p._fnm_grad_sink, p._fnm_sink_uses = view, 0
                if p.is_complex():
                    view._fnm_on_ready = (lambda o=off, m=n, q=p: self._sink_ready(o, m, q)) if self.overlap else None
                    view._fnm_flush = self.flush_ready if self.overlap else None

    def zero_(self):
        """Replacement for ``optimizer.zero_grad()``: keeps the views alive (set_to_none would drop them)."""
        self.flat[:self.head].zero_()
        for p in self.params:
            if p.grad is None:
                self.attach()
                break
        if self.direct:
            for p in self.params:
                p._fnm_sink_uses = 0

    def _sink_ready(self, off, n, param):
        """A backward kernel that fills the slot [off, off + n) has been enqueued: stage it for the next ``flush_ready``."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            self._staged.append((off, n, param))

    def flush_ready(self):
        """Start the exchange of the staged slots: adjacent ones (the two spectral weights of a layer) travel as ONE
        all-reduce -- per message NCCL's NVLS kernels need ~0.1 ms to get up to speed, and the exchange of the last
        layer is what remains exposed after backward."""
        staged, self._staged = sorted(self._staged, key=lambda t: t[0]), []
        runs = []
        for off, n, q in staged:
            if runs and runs[-1][1] == off:
                runs[-1][1] = off + (n + 63) // 64 * 64
                runs[-1][2].append(q)
            else:
                runs.append([off, off + (n + 63) // 64 * 64, [q]])
        for lo, hi, ps in runs:
            chunk = self.flat[lo:min(hi, self.flat.numel())]
            self._pending.append((dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True), ps))
            if self._reserve_sms > 0 and self._limit_prev is None:
                # from the first exchange in flight on, the persistent kernels leave `_reserve_sms` SMs to the NCCL kernels
                # (train_step lifts the cap after backward); the kernels before it keep every SM
                from ._lib import set_sm_limit
                sms = torch.cuda.get_device_properties(chunk.device).multi_processor_count
                self._limit_prev = set_sm_limit(max(1, sms - self._reserve_sms))

    def reserve_begin(self, sms):
        """Arm the SM reservation for the backward pass that follows (see ``train_step``)."""
        self._reserve_sms, self._limit_prev = int(sms), None

    def reserve_end(self):
        if self._limit_prev is not None:
            from ._lib import set_sm_limit
            set_sm_limit(self._limit_prev)
        self._reserve_sms, self._limit_prev = 0, None

    def all_reduce(self, group=None, async_op=False, join=True):
        """SUM over ranks; no-op in a single-process run.  With ``overlap`` the sinks are already in flight: this
        exchanges the head (everything that is not a sink) and, unless ``join=False``, waits for all of it
        (``join=False`` leaves the waiting to whoever iterates ``update_groups()``)."""
        group = self.group if group is None else group
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        if self._staged:
            self.flush_ready()
        if self.overlap and self._pending:
            work = dist.all_reduce(self.flat[:self.head], op=dist.ReduceOp.SUM, group=group, async_op=True)
            sunk = {id(q) for _, ps in self._pending for q in ps}
            self._pending.append((work, [q for q in self.params if id(q) not in sunk]))
            if join:
                for _ in self.update_groups():
                    pass
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def update_groups(self):
        """Yields lists of parameters in the order their gradients finish the exchange (each list after waiting for its
        all-reduce); empty when nothing is pending."""
        pending, self._pending = self._pending, []
        for work, ps in pending:
            work.wait()
            yield ps


def _dense(t):
    """Non-overlapping and dense (any permutation of a contiguous tensor): numel == storage span."""
    span = 1 + sum((n - 1) * st for n, st in zip(t.shape, t.stride())) if t.numel() else 0
    return span == t.numel()


def storage_order_view(t):
    """1-D contiguous view over the storage of a dense tensor (collectives need contiguous buffers)."""
    return torch.as_strided(t, (t.numel(),), (1,)) if _dense(t) else t.contiguous().view(-1)


def shard_batch(x, rank, world):
    """Rank r trains on samples [r*B/G, (r+1)*B/G)."""
    b = x.shape[0]
    if b % world:
        raise ValueError(f"global batch {b} is not divisible by {world} ranks")
    per = b // world
    return x[rank * per:(rank + 1) * per]


def broadcast_parameters(module, src=0):
    """Make every rank start from rank ``src``'s weights."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            data = storage_order_view(t.data)
            dist.broadcast(torch.view_as_real(data) if data.is_complex() else data, src)


def train_step(model, optimizer, grads, loss_fn, x, y, group=None, sm_reserve=0):
    """zero_grad -> forward -> loss -> backward -> SUM all-reduce -> Adam (train_fno.py:187-197 +
    the exchange).  Returns the LOCAL loss tensor (no host sync).

    ``sm_reserve`` (overlapped exchange only): SMs the persistent backward kernels leave free from the moment the first
    slot's all-reduce is in flight, so that the NCCL kernels have somewhere to run (include/fnm_b200.h, fnm_set_sm_limit);
    it should equal the number of CTAs NCCL launches (NCCL_MAX_CTAS)."""
    group = getattr(grads, "group", None) if group is None else group
    grads.zero_()
    out = model(x)
    loss = loss_fn(out, y)
    reserve = int(sm_reserve) if (grads.overlap and dist.is_available() and dist.is_initialized()
                                  and dist.get_world_size(group) > 1) else 0
    if reserve > 0:
        grads.reserve_begin(reserve)
        try:
            loss.backward()
        finally:
            grads.reserve_end()
    else:
        loss.backward()
    if grads.overlap and getattr(optimizer, "supports_partial_step", False) and dist.is_available() \
            and dist.is_initialized() and dist.get_world_size(group) > 1:
        # every parameter group is updated as soon as ITS exchange has landed (in issue order: deepest layer first, the
        # small real parameters last): the all-reduce of the last layer's slots finishes under the Adam launches of the others
        grads.all_reduce(group, join=False)
        first = True
        for ps in grads.update_groups():
            optimizer.step(params=ps, continued=not first)
            first = False
        if first:
            optimizer.step()
    else:
        grads.all_reduce(group)
        optimizer.step()
    return loss


class GraphedTrainStep:
    """The whole train step (zero_grad -> forward -> loss -> backward -> all-reduce -> Adam) captured in ONE
    CUDA graph and replayed: ~500 kernel launches per step become one graph launch, which is what the
    launch-bound small configurations need.  ``optimizer`` must be
    ``fnm_b200.Adam(..., capturable=True)``.  Inputs are copied into static buffers before each replay; the
    returned loss tensor is overwritten by every replay."""

    def __init__(self, model, optimizer, grads, loss_fn, x, y, group=None, warmup=3, sm_reserve=0):
        """``warmup`` eager steps run first on a side stream (allocator, workspaces, Adam state).  They are real
        optimizer updates, so parameters, moments and step counts are snapshotted before and restored after: building
        the graph leaves the model and the optimizer exactly as they were (moments that did not exist yet are zeroed,
        which is what lazy creation at the first replay would have produced)."""
        if not getattr(optimizer, "capturable", False):
            raise ValueError("GraphedTrainStep needs Adam(capturable=True): the step count must live on the device")
        self.x, self.y = x.clone(), y.clone()
        params = [p for g in optimizer.param_groups for p in g["params"]]
        saved_p = [p.detach().clone(memory_format=torch.preserve_format) for p in params]
        saved_s = []
        for p in params:
            st = optimizer.state.get(p) or {}
            saved_s.append({k: (v.detach().clone(memory_format=torch.preserve_format) if isinstance(v, torch.Tensor) else v)
                            for k, v in st.items()})
        # the capture stream differs from the stream the AccumulateGrad nodes were created on: expected here
        if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                          # allocator / workspace / Adam-state warm-up
            for _ in range(warmup):
                train_step(model, optimizer, grads, loss_fn, self.x, self.y, group, sm_reserve)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():                                  # undo the warm-up updates (in place: the graph records addresses)
            for p, p0, s0 in zip(params, saved_p, saved_s):
                p.copy_(p0)
                st = optimizer.state.get(p)
                if not st:
                    continue
                for k, v in st.items():
                    if isinstance(v, torch.Tensor):
                        v.copy_(s0[k]) if k in s0 else v.zero_()
                    else:
                        st[k] = s0.get(k, 0)
        if hasattr(optimizer, "reset_step_counter"):
            optimizer.reset_step_counter()
        from ._lib import launch_count
        self.graph = torch.cuda.CUDAGraph()
        n0 = launch_count()
        with torch.cuda.graph(self.graph):
            self.loss = train_step(model, optimizer, grads, loss_fn, self.x, self.y, group, sm_reserve)
        self.launches_per_step = launch_count() - n0
        self.optimizer = optimizer
        if hasattr(optimizer, "graph_captured"):
            optimizer.graph_captured()

    def close(self):
        """Drop the CUDA graph (call before ``dist.destroy_process_group()``: NCCL cannot tear a communicator down
        while a graph that captured its kernels is alive)."""
        self.graph = None
        self.loss = None

    def stage(self, x, y):
        """Start copying the NEXT step's inputs (pinned host or device tensors) into a staging pair on a side stream and
        return at once; the following ``__call__()`` without arguments consumes them.  An input pipeline calls this right
        after launching step i, so the host-to-device copy of step i + 1 runs under step i's kernels (the graph itself
        reads fixed buffers, which cannot be overwritten while a replay is in flight)."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.x.device)
            self._stage_x, self._stage_y = torch.empty_like(self.x), torch.empty_like(self.y)
            self._ev_ready = torch.cuda.Event()
            self._ev_taken = None
        if self._ev_taken is not None:                         # the previous step's copy out of the staging pair
            self._copy_stream.wait_event(self._ev_taken)
        with torch.cuda.stream(self._copy_stream):
            self._stage_x.copy_(x, non_blocking=True)
            self._stage_y.copy_(y, non_blocking=True)
            self._ev_ready.record()
        self._staged = True

    def __call__(self, x=None, y=None):
        if x is None:
            if not getattr(self, "_staged", False):
                raise RuntimeError("GraphedTrainStep(): no inputs given and none staged (call stage(x, y) first)")
            cur = torch.cuda.current_stream(self.x.device)
            cur.wait_event(self._ev_ready)
            self.x.copy_(self._stage_x, non_blocking=True)     # device to device: microseconds
            self.y.copy_(self._stage_y, non_blocking=True)
            self._ev_taken = torch.cuda.Event()
            self._ev_taken.record(cur)
            self._staged = False
        else:
            self.x.copy_(x, non_blocking=True)
            self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        if hasattr(self.optimizer, "graph_replayed"):
            self.optimizer.graph_replayed()
        return self.loss


def rel_l2_sum(out, y, eps=1e-6):
    """The reference's training loss: LpLoss(size_average=False) (utilities_module.py:188-204)."""
    from . import ops
    if ops.rel_l2_sum_supported(out, y):                       # CUDA: three launches instead of ~12 (ops.RelL2SumFn)
        return ops.rel_l2_sum(out, y, eps)
    b = out.shape[0]
    num = torch.norm(out.reshape(b, -1) - y.reshape(b, -1), 2, 1)
    den = torch.norm(y.reshape(b, -1), 2, 1) + eps
    return torch.sum(num / den)
