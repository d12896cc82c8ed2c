"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total/avg time, share.
    python tools/ncu_launch_summary.py gpurun_out/launches.csv > profiles/rN_launches_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"<.*", "<>", name) if "at::" in name or "cub::" in name else name
    rows.append((name, us))
agg = defaultdict(lambda: [0, 0.0])
for n, us in rows:
    agg[n][0] += 1
    agg[n][1] += us
tot = sum(v[1] for v in agg.values())
print("%d launches, %.1f us total (ncu per-launch times: cold cache, serialised)" % (len(rows), tot))
print("%-70s %7s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %7d %12.1f %10.2f %6.1f%%" % (n[:70], c, us, us / c, 100 * us / tot))
