"""Per-kernel summary (launches, ms, share, DRAM bytes per launch) of an ncu launch list captured with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv.
Usage: python scripts/kernel_summary.py launches.csv out.json ["source description"]"""
import csv
import json
import sys
from collections import defaultdict

path, out = sys.argv[1], sys.argv[2]
with open(path, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
per = defaultdict(lambda: {"launches": 0, "ns": 0.0, "dram": 0.0})
ids = set()
for r in csv.DictReader(lines):
    name = r["Kernel Name"].split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        ns = v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
        per[name]["ns"] += ns
        per[name]["launches"] += 1
    elif m.startswith("dram__bytes"):
        per[name]["dram"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
tot = sum(k["ns"] for k in per.values())
gemm = {n: k for n, k in per.items() if "conv_gemm_kernel" in n}
res = {"total_ms": tot / 1e6, "launches": sum(k["launches"] for k in per.values()),
       "conv_gemm_dram_bytes_per_launch": sum(k["dram"] for k in gemm.values()) / max(1, sum(k["launches"] for k in gemm.values())),
       "conv_gemm_share": sum(k["ns"] for k in gemm.values()) / tot,
       "kernels": {n: {"launches": k["launches"], "ms": k["ns"] / 1e6, "share": k["ns"] / tot,
                       "dram_bytes_per_launch": k["dram"] / k["launches"]}
                   for n, k in sorted(per.items(), key=lambda kv: -kv[1]["ns"])},
       "source": sys.argv[3] if len(sys.argv) > 3 else "ncu --metrics gpu__time_duration.sum,dram__bytes_{read,write}.sum --clock-control none --cache-control none, "
                 "one 512^2 batch-8 5-step decode (scripts/profile_decode.py 5 8 all)"}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: res[k] for k in ("total_ms", "launches", "conv_gemm_dram_bytes_per_launch", "conv_gemm_share")}))
