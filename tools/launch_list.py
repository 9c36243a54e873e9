"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` log by kernel: python tools/launch_list.py log.csv [header text]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ni, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0.0, 0])
for r in data:
    if len(r) <= vi:
        continue
    ms = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    agg[r[ni]][0] += ms
    agg[r[ni]][1] += 1
tot = sum(v[0] for v in agg.values())
for line in sys.argv[2:]:
    print("# " + line)
print(f"# total {tot:.1f} ms over {sum(v[1] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0]:10.2f} ms {v[1]:6d} launches {100 * v[0] / tot:5.1f}%  {k[:150]}")
