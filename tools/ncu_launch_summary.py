#!/usr/bin/env python
"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv` launch list.
usage: ncu_launch_summary.py X.csv "<command that was profiled>" > X.txt"""
import csv, re, sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], newline="")) if len(r) >= 15 and r[0].isdigit()]
tot = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").strip()
    name = re.sub(r"cub::CUB_\d+_NS::", "cub::", name)[:90]
    val, unit = float(r[14].replace(",", "")), r[13]
    us = val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1; t[1] += us
total = sum(v[1] for v in tot.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
print("(per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolute times, with bench.py)")
print(f"{len(rows)} launches listed, {total / 1e3:.2f} ms of kernel time in total\n")
print(f"{'launches':>8} {'total us':>12} {'avg us':>10} {'share':>7}  kernel")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:8d} {us:12.1f} {us / n:10.1f} {100 * us / total:6.1f}%  {name}")
