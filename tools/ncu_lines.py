#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: ncu_lines.py dump.csv [topN]  -> file:line, warp-instructions executed, stall samples, source text"""
import csv, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = []
cur_file = None
hdr = None
with open(path, newline="") as f:
    for r in csv.reader(f):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]; hdr = None; continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        # line-level rows have Address == '-'
        if r[2] != "-":
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        def num(k):
            try: return float(d.get(k, "0") or 0)
            except ValueError: return 0.0
        rows.append((cur_file, ln, num("Instructions Executed"), num("Warp Stall Sampling (All Samples)"),
                     num("Thread Instructions Executed"), r[1].strip()[:110]))
tot_i = sum(x[2] for x in rows); tot_s = sum(x[3] for x in rows)
print(f"total warp-inst {tot_i:.0f}  stall samples {tot_s:.0f}")
print("--- by instructions executed")
for x in sorted(rows, key=lambda x: -x[2])[:top]:
    print(f"{x[0]}:{x[1]:<5d} inst {100*x[2]/tot_i:5.1f}%  stall {100*x[3]/max(tot_s,1):5.1f}%  thr/inst {x[4]/max(x[2],1):4.1f}  {x[5]}")
print("--- by stall samples")
for x in sorted(rows, key=lambda x: -x[3])[:top]:
    print(f"{x[0]}:{x[1]:<5d} inst {100*x[2]/tot_i:5.1f}%  stall {100*x[3]/max(tot_s,1):5.1f}%  {x[5]}")
