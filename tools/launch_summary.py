"""Summarise an ncu launch list (gpu__time_duration.sum per launch) by kernel."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    m = re.search(r"kernel_entry(?:_maps)?<(?:swiftly::)?(\w+)<([^>]*)", name)
    key = f"{m.group(1)}<{m.group(2)}>" if m else name.split("(")[0][:60]
    tot[key][0] += 1
    tot[key][1] += val * scale
total = sum(v[1] for v in tot.values())
print(f"{'kernel':70s} {'launches':>9s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s}")
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} {n:9d} {ms:10.2f} {ms / n:9.4f} {100 * ms / total:6.1f}%")
print(f"{'TOTAL':70s} {sum(v[0] for v in tot.values()):9d} {total:10.2f}")
