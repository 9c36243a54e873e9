"""The data-parallel host logic (fnm_b200/parallel.py) on CPU with world_size 2 over gloo: batch sharding, the
flat gradient buffer and its ONE SUM all-reduce, parameter broadcast.  The kernels need a GPU, so the model here
is a small torch stand-in with the same ingredients (real + complex64 parameters, a loss that SUMS over the
batch); what is checked is that the two-rank step reproduces the single-process global-batch gradients exactly
as the reference's summed loss requires."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fnm_b200
from fnm_b200.parallel import FlatGradients, broadcast_parameters, rel_l2_sum, shard_batch


class Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(6, 5)
        self.w = torch.nn.Parameter(torch.rand(5, 5, 3, dtype=torch.cfloat) / 5)      # a complex "spectral" weight

    def forward(self, x):                                   # x: (B, 6, 8)
        h = self.lin(x.transpose(1, 2)).transpose(1, 2)     # (B, 5, 8)
        hf = torch.fft.rfft(h)[..., :3]
        y = torch.fft.irfft(torch.einsum("bix,iox->box", hf, self.w), n=8)
        return torch.tanh(h) + y


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, x, y, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                       # ranks start from DIFFERENT weights ...
        model = Toy()
        broadcast_parameters(model)                         # ... and must end up with rank 0's
        grads = FlatGradients(model.parameters())
        grads.zero_()
        loss = rel_l2_sum(model(shard_batch(x, rank, world)), shard_batch(y, rank, world))
        loss.backward()
        handle = grads.all_reduce()
        assert handle is None or handle.wait() is not False
        if rank == 0:
            out.put({"flat": grads.flat.clone(), "params": [p.detach().clone() for p in model.parameters()],
                     "views_alias_flat": all(p.grad.data_ptr() >= grads.flat.data_ptr() for p in model.parameters())})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sum_allreduce_equals_global_batch_gradient():
    world = 2
    torch.manual_seed(7)
    x, y = torch.randn(6, 6, 8), torch.randn(6, 5, 8)
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, x, y, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = out.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process reference: rank 0's weights (seed 100), the WHOLE batch, summed loss
    torch.manual_seed(100)
    model = Toy()
    for p, q in zip(model.parameters(), res["params"]):
        assert torch.equal(p.detach(), q)                   # broadcast_parameters delivered rank 0's weights
    grads = FlatGradients(model.parameters())
    grads.zero_()
    rel_l2_sum(model(x), y).backward()
    assert res["views_alias_flat"]
    assert torch.allclose(res["flat"], grads.flat, rtol=1e-5, atol=1e-7)


def _worker_overlap(rank, world, port, out):
    """Sink slots exchanged as the backward fills them (``_fnm_on_ready`` stages a slot, ``_fnm_flush`` starts the
    all-reduce of the staged, adjacent slots), the rest after:
    the result must be the plain flat SUM.  The spectral weight is stored mode-major like the drop-in modules' own."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(3)
        model = Toy()
        model.w = torch.nn.Parameter(model.w.detach().permute(2, 0, 1).contiguous().permute(1, 2, 0))   # mode-major storage
        model.w2 = torch.nn.Parameter(model.w.detach().clone(memory_format=torch.preserve_format))      # a layer's second weight
        broadcast_parameters(model)                          # storage-order view of a non-contiguous parameter
        grads = FlatGradients(model.parameters(), direct_spectral=True, overlap=True)
        assert model.w.grad.stride() == model.w.stride() and grads.head < grads.flat.numel()
        grads.zero_()
        grads.flat.copy_(torch.arange(grads.flat.numel(), dtype=torch.float32) * (rank + 1))
        sink, sink2 = model.w._fnm_grad_sink, model.w2._fnm_grad_sink
        sink2._fnm_on_ready()                                # the backward kernel would do this after writing a slot ...
        sink._fnm_on_ready()
        assert len(grads._staged) == 2 and not grads._pending
        sink._fnm_flush()                                    # ... and this once the layer's slots are staged: the two
        assert len(grads._pending) == 1 and not grads._staged    # adjacent slots travel as ONE all-reduce
        assert [id(q) for q in grads._pending[0][1]] == [id(model.w), id(model.w2)]
        grads.all_reduce()
        assert not grads._pending
        if rank == 0:
            out.put((grads.flat.clone(), list(grads.offsets)))
    finally:
        dist.destroy_process_group()


class _SinkFn(torch.autograd.Function):
    """Stand-in for ops.FourierLayerFn on CPU: y = sum(x) * (|w|^2 + |w2|^2).sum(); backward writes dW straight into the
    weights' gradient sinks, marks them ready, flushes (one merged all-reduce) and hands autograd no gradient for them."""

    @staticmethod
    def forward(ctx, x, w, w2):
        ctx.save_for_backward(x, w, w2)
        ctx.sinks = (w._fnm_grad_sink, w2._fnm_grad_sink)
        return x.sum() * ((w.abs() ** 2).sum() + (w2.abs() ** 2).sum())

    @staticmethod
    def backward(ctx, g):
        x, w, w2 = ctx.saved_tensors
        for sink, q in zip(ctx.sinks, (w, w2)):
            sink.copy_(2 * q.detach() * x.sum() * g)
            sink._fnm_on_ready()
        ctx.sinks[0]._fnm_flush()
        return g * ((w.abs() ** 2).sum() + (w2.abs() ** 2).sum()) * torch.ones_like(x), None, None


class _RecordingSGD:
    """Optimizer stand-in with Adam's partial-step interface: records which parameters each call updated."""
    supports_partial_step = True

    def __init__(self, params):
        self.params, self.calls = list(params), []

    def step(self, closure=None, params=None, continued=False):
        ps = self.params if params is None else list(params)
        self.calls.append(([id(p) for p in ps], continued))
        with torch.no_grad():
            for p in ps:
                p.sub_(0.01 * p.grad)


def _worker_train_step(rank, world, port, out):
    """parallel.train_step with the overlapped exchange and the grouped update, end to end on CPU: the spectral slots are
    exchanged from inside backward (merged), the rest after it; every parameter group is updated once, in the order its
    exchange was issued, with the SUM of the ranks' gradients."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(3)
        model = Toy()
        model.w2 = torch.nn.Parameter(model.w.detach().clone() * 0.5)
        scale = torch.nn.Parameter(torch.ones(3))                         # a real parameter that goes through autograd
        params = [scale, model.w, model.w2]
        grads = FlatGradients(params, direct_spectral=True, overlap=True)
        opt = _RecordingSGD(params)
        before = [p.detach().clone() for p in params]
        x = torch.full((4,), float(rank + 1))
        net = lambda inp: _SinkFn.apply(inp * scale.sum(), model.w, model.w2)
        fnm_b200.parallel.train_step(net, opt, grads, lambda o, y: o, x, None)
        assert not grads._pending and not grads._staged
        if rank == 0:
            out.put({"calls": opt.calls, "ids": [id(p) for p in params], "before": before,
                     "after": [p.detach().clone() for p in params], "grads": [p.grad.detach().clone() for p in params]})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_train_step_overlapped_exchange_and_grouped_update():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_train_step, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = out.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    scale_id, w_id, w2_id = res["ids"]
    # two update calls: the merged spectral group first (its exchange was issued inside backward), then the rest
    assert res["calls"] == [([w_id, w2_id], False), ([scale_id], True)]
    # gradients are the SUM over ranks (x = 1 on rank 0, 2 on rank 1: sum(x) * scale.sum() = 12 and 24)
    w0, w20 = res["before"][1], res["before"][2]
    assert torch.allclose(res["grads"][1], 2 * w0 * (12.0 + 24.0), rtol=1e-5)
    assert torch.allclose(res["grads"][2], 2 * w20 * (12.0 + 24.0), rtol=1e-5)
    q = ((w0.abs() ** 2).sum() + (w20.abs() ** 2).sum())
    assert torch.allclose(res["grads"][0], torch.full((3,), float(q) * (4.0 + 8.0)), rtol=1e-5)
    for b, a, g in zip(res["before"], res["after"], res["grads"]):
        assert torch.allclose(a, b - 0.01 * g, rtol=1e-6, atol=1e-7)      # every parameter updated exactly once


@pytest.mark.timeout(120)
def test_two_rank_overlapped_slot_exchange_equals_flat_sum():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_overlap, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    flat, offsets = out.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = torch.arange(flat.numel(), dtype=torch.float32) * 3          # rank 0 held 1x, rank 1 held 2x
    for off, n in offsets:                                                # every parameter's slot (alignment gaps are unused)
        assert torch.equal(flat[off:off + n], want[off:off + n])


def test_shard_batch_and_flat_layout():
    x = torch.arange(24.0).reshape(8, 3)
    assert torch.equal(shard_batch(x, 1, 4), x[2:4])
    with pytest.raises(ValueError):
        shard_batch(x, 0, 3)
    m = Toy()
    g = FlatGradients(m.parameters())
    n = sum(p.numel() * (2 if p.is_complex() else 1) for p in m.parameters())
    assert g.flat.numel() >= n and all(off % 64 == 0 for off, _ in g.offsets)
    g.flat.fill_(1.0)
    assert all(torch.all(torch.view_as_real(p.grad) == 1) if p.is_complex() else torch.all(p.grad == 1)
               for p in m.parameters())
    g.zero_()
    assert float(g.flat.abs().sum()) == 0.0
    assert fnm_b200.parallel.GraphedTrainStep is not None
