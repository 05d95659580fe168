#!/usr/bin/env python
"""Summarise an ncu capture of the scoring kernel into profiles/: a text file with the metrics the roofline argument
rests on and a small JSON that bench.py reads for roofline.traffic.
usage: ncu_summary.py <file.ncu-rep> <profiles/prefix> [n_decisions]"""
import csv, io, json, subprocess, sys

rep, prefix = sys.argv[1], sys.argv[2]
n_dec = int(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "sm__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
lines, js = [], []
for v in rows[2:]:
    d = dict(zip(h, v))
    u = dict(zip(h, units))
    rec = {}
    for k in KEYS:
        if k in d:
            lines.append(f"{k} = {d[k]} {u.get(k, '')}".rstrip())
            rec[k] = d[k]
    stalls = sorted(((float(d[k]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for k in h if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and d[k]),
                    reverse=True)[:8]
    lines.append("warps stalled per issue (top): " + ", ".join(f"{n}={x:.2f}" for x, n in stalls))
    def to_bytes(val, unit):
        val = float(val)
        return val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    rd = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]); wr = to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
    t_us = float(d["gpu__time_duration.sum"]) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u["gpu__time_duration.sum"], 1)
    lines.append(f"dram bytes per launch = {rd + wr:.0f} B (read {rd:.0f} + write {wr:.0f}); {(rd + wr) / t_us / 1e3:.1f} GB/s over the launch (profiled: cold caches, serialised)")
    if n_dec:
        lines.append(f"= {(rd + wr) / n_dec:.1f} B per decision over {n_dec} decisions (algorithmic: 1312 B + the 32-byte decision record)")
    lines.append("---")
    js.append({"kernel": d.get("Kernel Name", "")[:60], "dram_bytes_per_launch": rd + wr, "n_decisions": n_dec, "time_us": t_us})
open(prefix + ".txt", "w").write("\n".join(lines) + "\n")
json.dump(js[0] if js else {}, open(prefix + ".json", "w"), indent=1)
print("\n".join(lines))
