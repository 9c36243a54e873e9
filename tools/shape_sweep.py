"""Random-shape sweep of two chained fused Fourier layers (forward + all gradients) against the CPU oracle:
python tools/shape_sweep.py [cases] [seed].  Prints every case; exits 1 if one exceeds the 1e-5 gate."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_parity import check_fused_layers

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for i in range(n):
    C = rng.choice([8, 12, 16, 20, 32, 40, 48, 64, 96, 128, 128, 128])          # 128: the TMA / MN-major kernels
    B = rng.choice([1, 2, 3, 4, 5, 7, 20, 33])
    if rng.random() < 0.3:
        Pn = rng.randint(9, 200)
        P, m = (Pn,), (rng.randint(1, Pn // 2),)
    else:
        P1, P2 = rng.randint(6, 48), rng.randint(6, 72)
        P, m = (P1, P2), (rng.randint(1, P1 // 2), rng.randint(1, P2 // 2))
    try:
        check_fused_layers(B, C, P, m, seed=100 + i)
        print(f"ok   B={B} C={C} P={P} m={m}", flush=True)
    except AssertionError as e:
        bad += 1
        print(f"FAIL B={B} C={C} P={P} m={m}: {str(e).splitlines()[0][:160]}", flush=True)
    except Exception as e:                                   # noqa: BLE001
        bad += 1
        print(f"ERR  B={B} C={C} P={P} m={m}: {type(e).__name__}: {str(e)[:160]}", flush=True)
print(f"{n - bad} of {n} cases within 1e-5")
sys.exit(1 if bad else 0)
