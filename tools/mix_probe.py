"""The three compl_mul products at config-5 shapes (one launch each per repetition): the command ncu captures."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fnm_b200._lib import check, lib

B, R, NY, Ci, Co, rpw = 32, 64, 32, 128, 128, 32
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda")
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)
Xh = torch.randn(R, NY, 2, B, Ci, device=dev)
Gh = torch.randn(R, NY, 2, B, Co, device=dev)
Oh = torch.empty(R, NY, 2, B, Co, device=dev)
w0 = torch.randn(Ci, Co, rpw, NY, 2, device=dev)
w1 = torch.randn(Ci, Co, rpw, NY, 2, device=dev)
dw0, dw1 = torch.empty_like(w0), torch.empty_like(w1)
L = lib()
ws = torch.empty(L.fnm_mix_workspace_bytes(B, R, NY, Ci, Co), dtype=torch.uint8, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for r in range(reps):
    ev[0].record()
    check(L.fnm_mix_fwd(p(Xh), p(w0), p(w1), rpw, p(Oh), B, R, NY, Ci, Co, p(ws), ws.numel(), 0, st))
    ev[1].record()
    check(L.fnm_mix_bwd_x(p(Gh), p(w0), p(w1), rpw, p(Oh), B, R, NY, Ci, Co, p(ws), ws.numel(), 0, st))
    ev[2].record()
    check(L.fnm_mix_bwd_w(p(Xh), p(Gh), p(dw0), p(dw1), rpw, B, R, NY, Ci, Co, p(ws), ws.numel(), 0, st))
    ev[3].record()
torch.cuda.synchronize()
print("fwd %.3f ms  bwd_x %.3f ms  bwd_w %.3f ms (each incl. pack / unpack)" % tuple(ev[i].elapsed_time(ev[i + 1]) for i in range(3)))
