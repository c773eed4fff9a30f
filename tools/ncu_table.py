#!/usr/bin/env python
"""ncu_table.py <report.ncu-rep> : markdown table of the headline metrics of the (first) kernel in an ncu report."""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
idx = {h: i for i, h in enumerate(hdr)}
print("kernel:", vals[idx["Kernel Name"]][:120], " grid", vals[idx["Grid Size"]], " block", vals[idx["Block Size"]])
print("| metric | value | unit |\n|---|---|---|")
for m in WANT:
    if m in idx:
        print(f"| `{m}` | {vals[idx[m]]} | {units[idx[m]]} |")
