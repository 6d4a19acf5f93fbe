"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = defaultdict(lambda: [0, 0.0])
for r in csv.DictReader(rows):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] in ("us", "usecond"):
        v *= 1e3
    elif r["Metric Unit"] in ("ms", "msecond"):
        v *= 1e6
    agg[name][0] += 1
    agg[name][1] += v
total = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {total / 1e6:.3f} ms total (cold-cache, serialised)")
print(f"{'share':>7} {'ms':>10} {'launches':>8}  kernel")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * ns / total:6.2f}% {ns / 1e6:10.3f} {n:8d}  {name}")
