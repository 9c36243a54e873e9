"""pw2_kernel with parts of its pipeline switched off (library built with -DFNM_PW2_DEBUG_SWITCHES): which stream slows the
tensor pipe?  bits: 1 epilogue skips global stores | 2 converters skip the lo copy | 4 producer skips TMA loads | 8 no tcgen05.ld"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fnm_b200._lib import check, lib
from fnm_b200.plan import spectral_tables

B, C_, P1, P2, m1, m2 = 32, 128, 288, 288, 32, 32
dev = torch.device("cuda")
tab = spectral_tables(P1, P2, m1, m2).device(dev)
LY = 2 * m2
rows = B * P1
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
x = torch.randn(rows, P2, C_, device=dev)
out = torch.empty_like(x)
wc = torch.randn(C_, C_, device=dev) / C_ ** 0.5
bias = torch.randn(C_, device=dev)
Z = torch.randn(rows, LY, C_, device=dev)
L = lib()
for mode in (0, 1):
    for dbg in (0, 1, 2, 4, 8, 9, 3, 5, 6, 7, 15):
        os.environ["FNM_PW2_DEBUG"] = str(dbg)
        f = lambda: check(L.fnm_pointwise_ysyn(p(x), 0, p(wc), 0, p(bias), p(Z), p(tab["by"]), None, p(out), None, rows, P2, LY, C_, C_, mode, st))
        for _ in range(2):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            f()
        e1.record()
        torch.cuda.synchronize()
        print(f"mode {'tf32' if mode else '3xtf32'} debug {dbg:2d}: {e0.elapsed_time(e1) / 5:.3f} ms")
