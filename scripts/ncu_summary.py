#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics per kernel + top stall instructions from the source page."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "smsp__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__cycles_active.avg", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__maximum_warps_per_active_cycle_pct", "sm__ctas_launched.sum"]
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:75s} {r[i]} {units[i]}")
    print("---")
    break
if len(sys.argv) > 2 and sys.argv[2] == "--all":
    r = rows[2]
    for h, u, v in zip(hdr, units, r):
        print(f"{h:90s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = None; out = []; k = 0
for r in rows:
    if len(r) > 2 and r[0] == "Address":
        hdr = r; k += 1; continue
    if hdr is None or len(r) < len(hdr) - 2 or k != 1:
        continue
    out.append(dict(zip(hdr, r)))
tot = sum(int(d["# Samples"]) for d in out)
print("total samples", tot, "instructions", len(out), "warp-instr executed", sum(int(d["Instructions Executed"]) for d in out))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(d[h]) for d in out) for h in stalls}
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
for d in sorted(out, key=lambda d: -int(d["# Samples"]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    s = sorted(((int(d[h]), h) for h in stalls), reverse=True)[:2]
    print(d["# Samples"], d["Instructions Executed"], d["Source"].strip()[:90], s)
