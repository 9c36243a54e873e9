"""Per-role timeline of CTA 0 of pw2_kernel at config-5 shapes (library built with -DFNM_PW2_TRACE).
events: 0 mma:acc_empty ok | 1 mma:stage full | 2 mma:raw issued | 3 mma:lo tile ready | 4 mma:lo issued
        5 conv:start | 6 conv:done | 7 epi:wait acc | 8 epi:acc ready | 9 epi:done | 10 tma:stage free (issue)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fnm_b200._lib import LIB_PATH, check, lib
from fnm_b200.plan import spectral_tables

variant = sys.argv[1] if len(sys.argv) > 1 else "fwd"
B, C_, P1, P2, m1, m2 = 32, 128, 288, 288, 32, 32
dev = torch.device("cuda")
tab = spectral_tables(P1, P2, m1, m2).device(dev)
LY = 2 * m2
rows = B * P1
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
x = torch.randn(rows, P2, C_, device=dev)
out, out2 = torch.empty_like(x), torch.empty_like(x)
wc = torch.randn(C_, C_, device=dev) / C_ ** 0.5
bias = torch.randn(C_, device=dev)
Z = torch.randn(rows, LY, C_, device=dev)
L = lib()
B16 = lambda k: p(tab[k]) if k in tab else None          # present with FNM_PW2_CORR=bf16
mode = 1 if variant == "tf32" else 0
if len(sys.argv) > 2:
    os.environ["FNM_PW2_DEBUG"] = sys.argv[2]
for _ in range(3):
    if variant == "act":
        check(L.fnm_pointwise_ysyn_ex(p(x), 0, p(wc), 0, p(bias), p(Z), p(tab["by"]), None, p(out), p(out2), rows, P2, LY, C_, C_, mode, 0, B16("by16"), st))
    elif variant == "bwd":
        check(L.fnm_pointwise_ysyn_ex(p(x), 0, p(wc), 1, None, p(Z), p(tab["ayT"]), p(out2), p(out), None, rows, P2, LY, C_, C_, mode, 0, B16("ayT16"), st))
    else:
        check(L.fnm_pointwise_ysyn_ex(p(x), 0, p(wc), 0, p(bias), p(Z), p(tab["by"]), None, p(out), None, rows, P2, LY, C_, C_, mode, 0, B16("by16"), st))
torch.cuda.synchronize()
raw = C.CDLL(LIB_PATH)
if not hasattr(raw, "fnm_debug_pw2_trace"):
    print("library built without -DFNM_PW2_TRACE: kernels ran, no timeline")
    sys.exit(0)
NT, NE = 96, 24
buf = (C.c_longlong * (NT * NE))()
assert raw.fnm_debug_pw2_trace(buf, NT * NE) == 0
ev = [[buf[t * NE + e] for e in range(NE)] for t in range(NT)]
t00 = ev[0][10]
print(variant, "times in clocks relative to the first TMA issue; per tile: tma | mma acc_ok, full, raw_issued, lo_ready, lo_issued | conv start, done | epi wait, ready, done")
for t in range(40, 72):
    r = [v - t00 if v else -1 for v in ev[t]]
    print(f"{t:3d}: {r[10]:7d} | {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d} {r[4]:7d} | {r[5]:7d} {r[6]:7d} | {r[7]:7d} {r[8]:7d} {r[9]:7d}")
per = (ev[80][4] - ev[40][4]) / 40.0
print("steady-state clocks per tile:", per)
for name, a_, b_ in (("mma tile start -> raw issued", 1, 2), ("mma raw issued -> lo issued", 2, 4), ("conv work", 5, 6), ("epi wait", 7, 8), ("epi work", 8, 9)):
    d = [ev[t][b_] - ev[t][a_] for t in range(40, 80)]
    print(f"{name:30s} mean {sum(d) / len(d):8.1f}")
d = [ev[t + 1][1] - ev[t][4] for t in range(40, 80)]
print(f"{'mma lo issued -> next tile':30s} mean {sum(d) / len(d):8.1f}")
