"""Time xdft / xsyn (the x-axis table GEMMs) at config-5 shapes in fp32-parity and single-pass TF32 mode:
python tools/xdft_probe.py [iters]   (environment knobs of fnm_tc_tabgemm.cu apply: FNM_TAB_NO_ATMA, FNM_TAB_NO_TMA, FNM_TG_DEBUG)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fnm_b200._lib import check, lib
from fnm_b200.plan import spectral_tables

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B, C_, P1, P2, m1, m2 = 32, 128, 288, 288, 32, 32
dev = torch.device("cuda")
tab = spectral_tables(P1, P2, m1, m2).device(dev)
LY, R, NY = 2 * m2, 2 * m1, m2
rows = B * P1
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)
T = torch.randn(rows, LY, C_, device=dev)
Z = torch.empty(rows, LY, C_, device=dev)
Xh = torch.randn(R, NY, 2, B, C_, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L = lib()


def timeit(name, fn, bytes_):
    for _ in range(2):
        fn()
    ms = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1) / iters
    print(f"{name:28s} {ms:8.3f} ms   {bytes_ / 1e6 / ms:8.1f} GB/s algorithmic")


for mode, tag in ((0, "fp32"), (1, "tf32")):
    timeit(f"xdft {tag}", lambda: check(L.fnm_xdft(p(T), p(tab["ex"]), p(Xh), B, P1, R, NY, C_, mode, st)), (T.numel() + Xh.numel()) * 4)
    timeit(f"xsyn {tag}", lambda: check(L.fnm_xsyn(p(Xh), p(tab["fx"]), p(Z), B, P1, R, NY, C_, mode, st)), (Z.numel() + Xh.numel()) * 4)
