"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (+grid)."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
by_grid = "--grid" in sys.argv
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9}.get(unit, 1)
    name = r["Kernel Name"].split("(")[0]
    key = (name, r["Grid Size"], r["Block Size"]) if by_grid else (name,)
    rows.append((key, ns))
agg = defaultdict(lambda: [0, 0.0])
for k, ns in rows:
    agg[k][0] += 1
    agg[k][1] += ns
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e6:.3f} ms over {len(rows)} launches")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(40)]:
    print(f"{ns/1e6:9.3f} ms {100*ns/tot:5.1f}%  n={n:5d} avg={ns/n/1e3:8.1f} us  {' '.join(k)}")
