"""Kernel timeline of ONE data-parallel cfg5 train step on rank 0 (torch.profiler / CUPTI; compute and NCCL kernels with
their streams), to see how much of the gradient exchange runs under the backward pass:
python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dp_timeline.py OUT.txt
Environment: FNM_DP_SM_RESERVE, FNM_BENCH_OVERLAP (as bench.py), FNM_TL_GRAPH=0 for the eager step, FNM_TL_BATCH."""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import fnm_b200
from fnm_b200.parallel import FlatGradients, GraphedTrainStep, rel_l2_sum, train_step

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/dp_timeline.txt"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
graph = os.environ.get("FNM_TL_GRAPH", "1") == "1"
batch = int(os.environ.get("FNM_TL_BATCH", "32"))
reserve = int(os.environ.get("FNM_DP_SM_RESERVE", "0")) if world > 1 else 0
torch.manual_seed(0)
workload = os.environ.get("FNM_TL_WORKLOAD", "cfg5")            # any bench.py workload (single-GPU timelines of the small configs)
import bench as _bench
_desc, _family, _kw, _wbatch, _xshape, _ = _bench.WORKLOADS[workload]
if "FNM_TL_BATCH" not in os.environ:
    batch = _wbatch
model = _bench.build_model(_family, _kw).to(dev)
fnm_b200.parallel.broadcast_parameters(model)
opt = fnm_b200.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, capturable=graph)
grads = FlatGradients(model.parameters(), direct_spectral=True, overlap=os.environ.get("FNM_BENCH_OVERLAP", "1") == "1")
x = torch.randn(batch, *_xshape, device=dev)
with torch.no_grad():
    y = torch.randn_like(model(x))
if graph:
    g = GraphedTrainStep(model, opt, grads, rel_l2_sum, x, y, warmup=3, sm_reserve=reserve)
    step = lambda: g(x, y)
else:
    g = None
    step = lambda: train_step(model, opt, grads, rel_l2_sum, x, y, sm_reserve=reserve)
for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
if world > 1:
    dist.barrier()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
if world > 1:
    dist.barrier()
if rank == 0:
    evs = [e for e in prof.profiler.kineto_results.events()
           if e.device_type() == torch.autograd.DeviceType.CUDA and not re.search("memcpy|memset", e.name(), re.I)]
    evs.sort(key=lambda e: e.start_ns())
    evs = evs[len(evs) // 2:]                                  # the second of the two profiled steps
    t0 = evs[0].start_ns()
    short = lambda n: n.replace("void ", "").replace("fnm::", "").replace("tc::", "").split("(")[0][:46]
    lines = [f"# {workload} world {world} graph {int(graph)} reserve {reserve} overlap {int(grads.overlap)}: {ms:.3f} ms/step (5 steps, events)"]
    nccl = [(e.start_ns() - t0, e.end_ns() - t0) for e in evs if "nccl" in e.name().lower()]
    span = (max(e.end_ns() for e in evs) - t0) / 1e6
    lines.append(f"# profiled step span {span:.3f} ms, {len(evs)} kernels, NCCL kernels {len(nccl)}, NCCL busy {sum(b - a for a, b in nccl) / 1e6:.3f} ms")
    lines.append("# start_us  dur_us  stream  name")
    for e in evs:
        grid = ""
        if "nccl" in e.name().lower():
            m = re.search(r'"grid":\s*\[([^\]]*)\]', e.metadata_json() or "")
            grid = f"  grid [{m.group(1)}]" if m else "  " + (e.metadata_json() or "")[:200]
        lines.append(f"{(e.start_ns() - t0) / 1e3:10.1f} {e.duration_ns() / 1e3:9.1f}  {e.device_resource_id()}  {short(e.name())}{grid}")
    with open(out_path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:3]))
if g is not None:
    g.close()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
