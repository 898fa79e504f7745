"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import collections
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if unit in ("nsecond", "ns"):
        v /= 1000.0
    elif unit in ("msecond", "ms"):
        v *= 1000.0
    elif unit in ("second", "s"):
        v *= 1e6
    d = tot.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += v
total = sum(v[1] for v in tot.values())
print(f"{'kernel':70s} {'launches':>9s} {'total_us':>12s} {'avg_us':>9s} {'share':>7s}")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:9d} {t:12.1f} {t / n:9.2f} {100 * t / total:6.1f}%")
print(f"{'TOTAL':70s} {sum(v[0] for v in tot.values()):9d} {total:12.1f}")
