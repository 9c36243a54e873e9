"""Two-rank NCCL data-parallel parity worker (launched by test_two_gpu_nccl_data_parallel_matches_single_rank_global_batch
through torch.distributed.run; also runnable by hand:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dp_worker.py).

Every rank trains the same small FNO2d on its shard of a global batch with FlatGradients(direct_spectral=True,
overlap=True) + Adam(capturable=True) -- first eagerly, then replayed from a CUDA graph -- and rank 0 compares the
all-reduced gradients and the parameters after two steps with a single-rank run on the WHOLE batch.
Reference loop: /root/reference/advection_diffusion/func_to_func/train_fno.py:187-197 (loss summed over the batch)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch                      # noqa: E402
import torch.distributed as dist  # noqa: E402

import fnm_b200                   # noqa: E402
from fnm_b200.parallel import FlatGradients, GraphedTrainStep, broadcast_parameters, rel_l2_sum, shard_batch, train_step  # noqa: E402


def rel(a, b):
    a = torch.view_as_real(a) if a.is_complex() else a
    b = torch.view_as_real(b) if b.is_complex() else b
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()


def build(seed):
    torch.manual_seed(seed)
    return fnm_b200.FNO2d(4, 4, 32, n_layers=3, width_final=32)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(77)
    X = torch.randn(4 * world, 1, 24, 24, generator=g).to(dev)
    Y = torch.randn(4 * world, 1, 24, 24, generator=g).to(dev)
    xs, ys = shard_batch(X, rank, world).contiguous(), shard_batch(Y, rank, world).contiguous()

    # single-rank truth: the whole batch, no exchange
    truth = build(5).to(dev)
    opt_t = fnm_b200.Adam(truth.parameters(), lr=1e-3, weight_decay=1e-4)
    g_t = FlatGradients(truth.parameters(), direct_spectral=True)
    grads_truth = None
    for _ in range(2):
        g_t.zero_()
        rel_l2_sum(truth(X), Y).backward()
        if grads_truth is None:
            grads_truth = {k: p.grad.detach().clone() for k, p in truth.named_parameters()}
        opt_t.step()

    worst = 0.0
    graphs = []
    for how in ("eager", "graph"):
        model = build(100 + rank).to(dev)                        # ranks start from different weights ...
        with torch.no_grad():
            if rank == 0:
                model.load_state_dict(build(5).state_dict())
        broadcast_parameters(model)                              # ... and continue from rank 0's (= the truth's)
        opt = fnm_b200.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)
        grads = FlatGradients(model.parameters(), direct_spectral=True, overlap=True)
        first = None
        if how == "eager":
            for _ in range(2):
                train_step(model, opt, grads, rel_l2_sum, xs, ys)
                if first is None:
                    first = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        else:
            step = GraphedTrainStep(model, opt, grads, rel_l2_sum, xs, ys, warmup=3, sm_reserve=8)    # with the SM reservation of the bench
            graphs.append(step)
            for _ in range(2):
                step(xs, ys)
                if first is None:
                    torch.cuda.synchronize()
                    first = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        torch.cuda.synchronize()
        for k, v in first.items():
            e = rel(v, grads_truth[k])
            worst = max(worst, e)
            assert e < 1e-5, f"{how}: all-reduced gradient {k} differs from the global-batch gradient: {e:.3e}"
        for (k, a), b in zip(model.state_dict().items(), truth.state_dict().values()):
            e = rel(a, b)
            worst = max(worst, e)
            assert e < 1e-5, f"{how}: parameter {k} after 2 steps differs from the single-rank run: {e:.3e}"
        # replicas stay identical
        flat = torch.cat([torch.view_as_real(p.detach()).reshape(-1) if p.is_complex() else p.detach().reshape(-1)
                          for p in model.parameters()])
        other = flat.clone()
        dist.broadcast(other, 0)
        assert torch.equal(flat, other), f"{how}: rank {rank} diverged from rank 0"
    for s in graphs:                                             # graphs first, then the communicator: a clean shutdown
        s.close()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print(f"DP_PARITY_OK world={world} worst_rel={worst:.3e}")


if __name__ == "__main__":
    main()
