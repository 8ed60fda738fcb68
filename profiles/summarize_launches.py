"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share.

  python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt
Times under ncu are cold-cache and serialised: compare SHARES with bench.py's CUDA-event breakdown, not absolutes.
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
n = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    n += 1
tot = sum(a[1] for a in agg.values())
print(f"# {path}: {n} launches, {tot:.1f} us total (ncu-serialised, cold cache)")
print(f"{'time_us':>10} {'share':>6} {'launches':>8}  kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1]:10.1f} {100 * a[1] / tot:5.1f}% {a[0]:8d}  {k}")
