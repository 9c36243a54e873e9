"""Accuracy probe of the 3xTF32 stages at full BASELINE sizes against float64 (run on the GPU box)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fnm_b200._lib import check, lib

dev = torch.device("cuda")
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
L = lib()


def rel(a, b):
    return ((a.double() - b).norm() / b.norm()).item()


torch.manual_seed(0)
for npos, Cc, mean in [(288 * 288 * 2, 128, 0.0), (288 * 288 * 32, 128, 0.0), (288 * 288 * 32, 128, 0.3), (1152 * 64, 64, 0.3)]:
    g = torch.randn(npos, Cc, device=dev)
    x = torch.randn(npos, Cc, device=dev) + mean
    ws = torch.empty(L.fnm_conv_wgrad_workspace_bytes(npos, Cc, Cc), dtype=torch.uint8, device=dev)
    dwc, dbc = torch.empty(Cc, Cc, device=dev), torch.empty(Cc, device=dev)
    check(L.fnm_conv_wgrad(p(g), p(x), 0, p(dwc), p(dbc), npos, Cc, Cc, p(ws), ws.numel(), 0, st))
    ref = torch.zeros(Cc, Cc, device=dev, dtype=torch.float64)
    for i in range(0, npos, 1 << 18):
        ref += g[i:i + (1 << 18)].double().t() @ x[i:i + (1 << 18)].double()
    f32 = g.t() @ x
    print(f"conv_wgrad npos={npos} C={Cc} mean={mean}: ours {rel(dwc, ref):.2e}  torch fp32 {rel(f32, ref):.2e}  db {rel(dbc, g.double().sum(0)):.2e}")

for rows, P2, LY, Cc, mean in [(64, 1152, 32, 64, 0.3), (2048, 288, 64, 128, 0.3), (2048, 288, 64, 128, 0.0)]:
    x = torch.randn(rows, P2, Cc, device=dev) + mean
    y = torch.arange(P2, device=dev, dtype=torch.float64)
    l = torch.arange(LY // 2, device=dev, dtype=torch.float64)
    ang = 2 * torch.pi * l[:, None] * y[None, :] / P2
    tab64 = torch.cat([torch.cos(ang), -torch.sin(ang)], 0)
    tab = tab64.float().contiguous()
    T = torch.empty(rows, LY, Cc, device=dev)
    check(L.fnm_ydft(p(x), p(tab), p(T), rows, P2, LY, Cc, 0, 0, st))
    ref = torch.einsum("ly,ryc->rlc", tab.double(), x.double())
    f32 = torch.einsum("ly,ryc->rlc", tab, x)
    print(f"ydft rows={rows} P2={P2} C={Cc} mean={mean}: ours {rel(T, ref):.2e}  torch fp32 {rel(f32, ref):.2e}")
    for lo in (0, 1, LY // 2 - 1):
        print(f"   mode {lo}: ours {rel(T[:, lo], ref[:, lo]):.2e}  torch fp32 {rel(f32[:, lo], ref[:, lo]):.2e}")
