"""A/B timing of fnm_pointwise_ysyn between two builds of the library: python tools/pw2_ab.py libA.so libB.so [cfg3|cfg5|cfg1|cfg4]"""
import ctypes as C
import sys

import torch

cfg = [a for a in sys.argv[1:] if not a.endswith(".so")][0]
B, C_, P1, P2, m2 = {"cfg5": (32, 128, 288, 288, 32), "cfg1": (20, 32, 72, 72, 12), "cfg4": (20, 64, 144, 144, 16),
                     "cfg3": (128, 32, 248, 57, 12)}[cfg]
dev = torch.device("cuda")
LY, rows = 2 * m2, B * P1
torch.manual_seed(0)
x = torch.randn(rows, P2, C_, device=dev)
out, out2 = torch.empty_like(x), torch.empty_like(x)
wc = torch.randn(C_, C_, device=dev) / C_ ** 0.5
bias = torch.randn(C_, device=dev)
Z = torch.randn(rows, LY, C_, device=dev)
by = torch.randn(P2, LY, device=dev)
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
A = x.numel() * 4
libs = [a for a in sys.argv[1:] if a.endswith(".so")]
for path in libs:
    L = C.CDLL(path)
    f = L.fnm_pointwise_ysyn
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_int64] + [C.c_int] * 5 + [C.c_void_p]
    for name, args, nb in (("fwd noact (2A)", (p(x), 0, p(wc), 0, p(bias), p(Z), p(by), None, p(out), None), 2),
                           ("fwd +act (3A)", (p(x), 0, p(wc), 0, p(bias), p(Z), p(by), None, p(out), p(out2)), 3),
                           ("bwd mul (3A)", (p(x), 0, p(wc), 1, None, p(Z), p(by), p(out2), p(out), None), 3),
                           ("bwd nomul (2A)", (p(x), 0, p(wc), 1, None, p(Z), p(by), None, p(out), None), 2),
                           ("ysyn only (LY, no x)", (None, 0, None, 0, None, p(Z), p(by), None, p(out), None), 1)):
        call = lambda: f(*args, rows, P2, LY, C_, C_, 0, st)
        for _ in range(3):
            assert call() == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{path.split('/')[-1]:22s} {cfg} {name:22s} {ms:8.4f} ms  {nb * A / 1e6 / ms:8.1f} GB/s")
