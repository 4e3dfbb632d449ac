#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python tools/launch_summary.py gpurun_out/launches.csv [first_launch_id] > profiles/rN_launches_summary.csv
(per-launch times are cold-cache and serialised: compare SHARES, not absolutes)"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = OrderedDict()
for r in rows:
    if int(r[0]) < first:
        continue
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("wm::", "").strip()
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
    ns = float(r[14].replace(",", ""))
    if r[13] == "us":
        ns *= 1e3
    elif r[13] == "ms":
        ns *= 1e6
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(a[1] for a in agg.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "total_ms", "share_pct"])
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    w.writerow([k, n, f"{t / 1e6:.3f}", f"{100 * t / tot:.1f}"])
w.writerow(["TOTAL", sum(a[0] for a in agg.values()), f"{tot / 1e6:.3f}", "100.0"])
