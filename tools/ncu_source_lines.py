#!/usr/bin/env python
"""Per-source-line totals (instructions executed, stall samples) from `ncu --page source --csv --print-source cuda,sass`.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_source_lines.py [top_n]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
c_line, c_src, c_inst, c_samp = hdr.index("Line No"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
lines, tot_i, tot_s = [], 0, 0
cur_file = ""
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        if r and r[0] == "File Path" or (len(r) >= 2 and r[0] in ("File", "File Path")):
            cur_file = r[-1]
        continue
    if r[c_line] in ("", "-"):   # a SASS row under a source line
        continue
    try:
        inst, samp = int(r[c_inst]), int(r[c_samp])
    except ValueError:
        continue
    st = {k: int(r[i]) for k, i in stall.items() if r[i].isdigit() and int(r[i])}
    lines.append((inst, samp, r[c_line], r[c_src].strip()[:110], st))
    tot_i += inst
    tot_s += samp
print(f"total instructions {tot_i}, samples {tot_s}")
for inst, samp, ln, src, st in sorted(lines, key=lambda x: -x[1])[:top]:
    tops = ",".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100.0 * samp / max(tot_s, 1):5.1f}% samp {100.0 * inst / max(tot_i, 1):5.1f}% inst  L{ln:>5}  {src}   [{tops}]")
