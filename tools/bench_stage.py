"""Time individual C-ABI stages at a BASELINE config's shapes (CUDA events, inputs larger than L2).
Usage: python tools/bench_stage.py [cfg5|cfg1] [iters]   -- also the command ncu captures."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fnm_b200._lib import check, lib
from fnm_b200.plan import spectral_tables

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B, C_, P1, P2, m1, m2 = {"cfg5": (32, 128, 288, 288, 32, 32), "cfg1": (20, 32, 72, 72, 12, 12),
                         "cfg4": (20, 64, 144, 144, 16, 16), "cfg3": (128, 32, 248, 57, 12, 12),
                         "cfg3b": (128, 32, 248, 64, 12, 16), "cfg3c": (128, 32, 248, 64, 12, 12), "cfg3d": (128, 64, 248, 57, 12, 12)}[cfg]
dev = torch.device("cuda")
tab = spectral_tables(P1, P2, m1, m2).device(dev)
LY, R, NY = 2 * m2, 2 * m1, m2
rows = B * P1
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)
x = torch.randn(rows, P2, C_, device=dev)
g = torch.randn(rows, P2, C_, device=dev)
out = torch.empty_like(x)
wc = torch.randn(C_, C_, device=dev) / C_ ** 0.5
bias = torch.randn(C_, device=dev)
Z = torch.randn(rows, LY, C_, device=dev)
T = torch.empty(rows, LY, C_, device=dev)
A = x.numel() * 4
L = lib()


def timeit(name, fn, bytes_):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:28s} {ms:8.3f} ms   {bytes_ / 1e6 / ms:8.1f} GB/s algorithmic")


timeit("pointwise_ysyn fwd (2A)", lambda: check(L.fnm_pointwise_ysyn(p(x), 1, p(wc), 0, p(bias), p(Z), p(tab["by"]), None, p(out), None, rows, P2, LY, C_, C_, 0, st)), 2 * A)
timeit("pointwise_ysyn fwd noact", lambda: check(L.fnm_pointwise_ysyn(p(x), 0, p(wc), 0, p(bias), p(Z), p(tab["by"]), None, p(out), None, rows, P2, LY, C_, C_, 0, st)), 2 * A)
timeit("pointwise_ysyn noact tf32", lambda: check(L.fnm_pointwise_ysyn(p(x), 0, p(wc), 0, p(bias), p(Z), p(tab["by"]), None, p(out), None, rows, P2, LY, C_, C_, 1, st)), 2 * A)
timeit("pointwise_ysyn fwd tf32", lambda: check(L.fnm_pointwise_ysyn(p(x), 1, p(wc), 0, p(bias), p(Z), p(tab["by"]), None, p(out), None, rows, P2, LY, C_, C_, 1, st)), 2 * A)
out2 = torch.empty_like(x)
timeit("pointwise_ysyn fwd +act out (3A)", lambda: check(L.fnm_pointwise_ysyn(p(x), 0, p(wc), 0, p(bias), p(Z), p(tab["by"]), None, p(out), p(out2), rows, P2, LY, C_, C_, 0, st)), 3 * A)
timeit("pointwise_ysyn bwd tf32", lambda: check(L.fnm_pointwise_ysyn(p(g), 0, p(wc), 1, None, p(Z), p(tab["ayT"]), p(x), p(out), None, rows, P2, LY, C_, C_, 1, st)), 3 * A)
timeit("pointwise_ysyn bwd (3A)", lambda: check(L.fnm_pointwise_ysyn(p(g), 0, p(wc), 1, None, p(Z), p(tab["ayT"]), p(x), p(out), None, rows, P2, LY, C_, C_, 0, st)), 3 * A)
os.environ["FNM_PW2_CORR"] = "bf16"
from fnm_b200.plan import bf16_table_pair
tab16 = {k: bf16_table_pair(tab[k]) for k in ("by", "ayT")}
timeit("pw_ysyn fwd noact, bf16 corr", lambda: check(L.fnm_pointwise_ysyn(p(x), 0, p(wc), 0, p(bias), p(Z), p(tab["by"]), None, p(out), None, rows, P2, LY, C_, C_, 0, st)), 2 * A)
timeit("pw_ysyn bwd (3A), bf16 corr", lambda: check(L.fnm_pointwise_ysyn(p(g), 0, p(wc), 1, None, p(Z), p(tab["ayT"]), p(x), p(out), None, rows, P2, LY, C_, C_, 0, st)), 3 * A)
timeit("pw_ysyn fwd noact, bf16 corr +by16", lambda: check(L.fnm_pointwise_ysyn_ex(p(x), 0, p(wc), 0, p(bias), p(Z), p(tab["by"]), None, p(out), None, rows, P2, LY, C_, C_, 0, 0, p(tab16["by"]), st)), 2 * A)
timeit("pw_ysyn bwd cached grad, bf16 corr +by16", lambda: check(L.fnm_pointwise_ysyn_ex(p(g), 0, p(wc), 1, None, p(Z), p(tab["ayT"]), p(x), p(out), None, rows, P2, LY, C_, C_, 0, 1, p(tab16["ayT"]), st)), 3 * A)
del os.environ["FNM_PW2_CORR"]
timeit("pw_ysyn fwd grad|act out (3A)", lambda: check(L.fnm_pointwise_ysyn_ex(p(x), 0, p(wc), 0, p(bias), p(Z), p(tab["by"]), None, p(out), p(out2), rows, P2, LY, C_, C_, 0, 2, None, st)), 3 * A)
timeit("pw_ysyn bwd cached grad (3A)", lambda: check(L.fnm_pointwise_ysyn_ex(p(g), 0, p(wc), 1, None, p(Z), p(tab["ayT"]), p(x), p(out), None, rows, P2, LY, C_, C_, 0, 1, None, st)), 3 * A)
timeit("ydft (A + T)", lambda: check(L.fnm_ydft(p(x), p(tab["ay"]), p(T), rows, P2, LY, C_, 1, 0, st)), A + T.numel() * 4)
timeit("ydft noact (A + T)", lambda: check(L.fnm_ydft(p(x), p(tab["ay"]), p(T), rows, P2, LY, C_, 0, 0, st)), A + T.numel() * 4)
timeit("ydft noact tf32", lambda: check(L.fnm_ydft(p(x), p(tab["ay"]), p(T), rows, P2, LY, C_, 0, 1, st)), A + T.numel() * 4)
Xh = torch.empty(R, NY, 2, B, C_, device=dev)
timeit("xdft (T + Xh)", lambda: check(L.fnm_xdft(p(T), p(tab["ex"]), p(Xh), B, P1, R, NY, C_, 0, st)), (T.numel() + Xh.numel()) * 4)
timeit("xsyn (Xh + Z)", lambda: check(L.fnm_xsyn(p(Xh), p(tab["fx"]), p(Z), B, P1, R, NY, C_, 0, st)), (Z.numel() + Xh.numel()) * 4)
w0 = torch.randn(C_, C_, m1, m2, dtype=torch.complex64, device=dev) / C_
w1 = torch.randn(C_, C_, m1, m2, dtype=torch.complex64, device=dev) / C_
w0r, w1r = torch.view_as_real(w0), torch.view_as_real(w1)
mws = torch.empty(max(L.fnm_mix_workspace_bytes(B, R, NY, C_, C_), 16), dtype=torch.uint8, device=dev)
Oh = torch.empty_like(Xh)
WB = w0.numel() * 16
timeit("mix_fwd (W + 2 spectra)", lambda: check(L.fnm_mix_fwd(p(Xh), p(w0r), p(w1r), m1, p(Oh), B, R, NY, C_, C_, p(mws), mws.numel(), 0, st)), WB + 2 * Xh.numel() * 4)
timeit("mix_bwd_x", lambda: check(L.fnm_mix_bwd_x(p(Xh), p(w0r), p(w1r), m1, p(Oh), B, R, NY, C_, C_, p(mws), mws.numel(), 0, st)), WB + 2 * Xh.numel() * 4)
dw0, dw1 = torch.empty_like(w0r), torch.empty_like(w1r)
timeit("mix_bwd_w", lambda: check(L.fnm_mix_bwd_w(p(Xh), p(Oh), p(dw0), p(dw1), m1, B, R, NY, C_, C_, p(mws), mws.numel(), 0, st)), WB + 2 * Xh.numel() * 4)
ws = torch.empty(L.fnm_conv_wgrad_workspace_bytes(rows * P2, C_, C_), dtype=torch.uint8, device=dev)
dwc, dbc = torch.empty(C_, C_, device=dev), torch.empty(C_, device=dev)
timeit("conv_wgrad noact (2A)", lambda: check(L.fnm_conv_wgrad(p(g), p(x), 0, p(dwc), p(dbc), rows * P2, C_, C_, p(ws), ws.numel(), 0, st)), 2 * A)
timeit("ydft + conv_wgrad fused (2A + T)", lambda: check(L.fnm_ydft_conv_wgrad(p(g), p(x), p(tab["ay"]), p(T), p(dwc), p(dbc), rows, P2, LY, C_, p(ws), ws.numel(), 0, st)), 2 * A + T.numel() * 4)
timeit("conv_wgrad (2A)", lambda: check(L.fnm_conv_wgrad(p(g), p(x), 1, p(dwc), p(dbc), rows * P2, C_, C_, p(ws), ws.numel(), 0, st)), 2 * A)
y = torch.empty_like(x)
timeit("torch copy (2A) reference", lambda: y.copy_(x), 2 * A)
