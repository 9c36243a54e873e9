"""All GPU kernels of one eager cfg5 train step (ours AND torch's), aggregated by name with torch.profiler."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import fnm_b200
from fnm_b200.parallel import FlatGradients, rel_l2_sum

dev = torch.device("cuda")
torch.manual_seed(0)
model = fnm_b200.FNO2d(32, 32, 128).to(dev)
opt = fnm_b200.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
grads = FlatGradients(model.parameters(), direct_spectral=True)
x = torch.randn(32, 1, 256, 256, device=dev)
y = torch.randn(32, 1, 256, 256, device=dev)


def step():
    grads.zero_()
    loss = rel_l2_sum(model(x), y)
    loss.backward()
    grads.all_reduce()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=False) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total / 1e3) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"total kernel time {tot:.2f} ms")
for k, n, ms in rows[:40]:
    print(f"{ms:8.3f} ms  x{n:<3d} {k[:110]}")

print("fills / zeros / adds by input shape:")
for e in prof.key_averages(group_by_input_shape=True):
    if e.key in ("aten::fill_", "aten::zero_", "aten::zeros", "aten::zeros_like", "aten::add_", "aten::add", "aten::copy_") and e.device_time_total > 5:
        print(f"  {e.key:18s} x{e.count:<3d} {e.device_time_total / 1e3:7.3f} ms  {e.input_shapes}")
