"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v))
agg = defaultdict(lambda: [0, 0.0])
for name, us in rows:
    key = name.replace("(anonymous namespace)::", "").replace("void ", "")
    key = re.sub(r"\(.*", "", key)
    key = re.sub(r"<(unnamed)>::", "", key)[:70]
    agg[key][0] += 1
    agg[key][1] += us
total = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {total / 1e3:.1f} ms total (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':70s} {'launches':>8s} {'ms':>9s} {'share':>7s} {'us/launch':>10s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k:70s} {n:8d} {us / 1e3:9.2f} {100 * us / total:6.1f}% {us / n:10.1f}")
