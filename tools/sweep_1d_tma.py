"""Random 1-D shapes inside the envelope of the TMA-fed y-DFT (C in {32, 64, 128}, 8 <= 2 m <= 64): few long rows take the
K-segmented kernel + the summing thin x-transform, row lengths that are no multiple of 4 the padded tables.
python tools/sweep_1d_tma.py [cases] [seed]"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_parity import check_fused_layers

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for i in range(n):
    C = rng.choice([32, 64, 128])
    B = rng.choice([1, 2, 3, 5, 9, 20, 37, 80])
    P = rng.randint(64, 1300)
    m = rng.randint(4, 32)
    try:
        check_fused_layers(B, C, (P,), (m,), seed=200 + i)
        print(f"ok   B={B} C={C} P={P} m={m}", flush=True)
    except AssertionError as e:
        bad += 1
        print(f"FAIL B={B} C={C} P={P} m={m}: {str(e).splitlines()[0][:160]}", flush=True)
    except Exception as e:                                   # noqa: BLE001
        bad += 1
        print(f"ERR  B={B} C={C} P={P} m={m}: {type(e).__name__}: {str(e)[:160]}", flush=True)
print(f"{n - bad} of {n} cases within 1e-5")
sys.exit(1 if bad else 0)
