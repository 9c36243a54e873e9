"""One launch of each tcgen05 kernel at config-5 shapes (fp32 parity mode): the command ncu captures."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fnm_b200._lib import check, lib
from fnm_b200.plan import spectral_tables

B, C_, P1, P2, m1, m2 = 32, 128, 288, 288, 32, 32
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda")
tab = spectral_tables(P1, P2, m1, m2).device(dev)
LY, R, NY = 2 * m2, 2 * m1, m2
rows = B * P1
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)
x = torch.randn(rows, P2, C_, device=dev)
g = torch.randn(rows, P2, C_, device=dev)
out, out2 = torch.empty_like(x), torch.empty_like(x)
wc = torch.randn(C_, C_, device=dev) / C_ ** 0.5
bias = torch.randn(C_, device=dev)
Z = torch.randn(rows, LY, C_, device=dev)
T = torch.empty(rows, LY, C_, device=dev)
Xh = torch.empty(R, NY, 2, B, C_, device=dev)
L = lib()
ws = torch.empty(L.fnm_conv_wgrad_workspace_bytes(rows * P2, C_, C_), dtype=torch.uint8, device=dev)
dwc, dbc = torch.empty(C_, C_, device=dev), torch.empty(C_, device=dev)
for _ in range(reps):
    check(L.fnm_pointwise_ysyn(p(x), 0, p(wc), 0, p(bias), p(Z), p(tab["by"]), None, p(out), None, rows, P2, LY, C_, C_, 0, st))
    # what the trunk runs since round 2: forward writes GELU'(z) | GELU(z) (FNM_PW_OUT_IS_GRAD = 2), backward multiplies by the
    # cached derivative (FNM_PW_MUL_IS_GRAD = 1)
    check(L.fnm_pointwise_ysyn_ex(p(x), 0, p(wc), 0, p(bias), p(Z), p(tab["by"]), None, p(out), p(out2), rows, P2, LY, C_, C_, 0, 2, None, st))
    check(L.fnm_pointwise_ysyn_ex(p(g), 0, p(wc), 1, None, p(Z), p(tab["ayT"]), p(x), p(out), None, rows, P2, LY, C_, C_, 0, 1, None, st))
    check(L.fnm_ydft(p(x), p(tab["ay"]), p(T), rows, P2, LY, C_, 0, 0, st))
    check(L.fnm_xdft(p(T), p(tab["ex"]), p(Xh), B, P1, R, NY, C_, 0, st))
    check(L.fnm_xsyn(p(Xh), p(tab["fx"]), p(Z), B, P1, R, NY, C_, 0, st))
    check(L.fnm_conv_wgrad(p(g), p(x), 0, p(dwc), p(dbc), rows * P2, C_, C_, p(ws), ws.numel(), 0, st))
    # backward y-DFT of g fused with the conv weight gradient (one read of g)
    check(L.fnm_ydft_conv_wgrad(p(g), p(x), p(tab["ay"]), p(T), p(dwc), p(dbc), rows, P2, LY, C_, p(ws), ws.numel(), 0, st))
torch.cuda.synchronize()
print("ok")
