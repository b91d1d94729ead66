"""Sum an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name: python tools/launch_summary.py file.csv [skip_first_n]"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((r["Kernel Name"], v))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = defaultdict(lambda: [0, 0.0])
for k, v in rows:
    k = re.sub(r"\(.*", "", k)
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in rows)
print("%d launches, %.3f ms" % (len(rows), tot))
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %5d %10.3f ms %5.1f %%" % (k[:70], n, v, 100 * v / tot))
