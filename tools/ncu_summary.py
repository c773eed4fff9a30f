#!/usr/bin/env python
"""ncu_summary.py REPORT.ncu-rep [title] -> markdown table of the per-kernel metrics the judge asks for (duration, DRAM bytes and GB/s,
DRAM %, tensor-pipe %, issue %, L2 hit, registers, smem), one row per captured launch.  Runs `ncu -i ... --page raw --csv` (no GPU needed)."""
import csv, subprocess, sys
rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
def col(name):
    for i, h in enumerate(hdr):
        if h == name or h.endswith("." + name) or h.endswith(name):
            return i
    return None
want = [("time us", "gpu__time_duration.sum"), ("dram rd MB", "dram__bytes_read.sum"), ("dram wr MB", "dram__bytes_write.sum"),
        ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
        ("regs", "launch__registers_per_thread"), ("block", "launch__block_size"), ("grid", "launch__grid_size"), ("smem KB", "launch__shared_mem_per_block_dynamic")]
units = rows[1]
def val(r, name):
    i = col(name)
    if i is None or not r[i]:
        return None
    v = float(r[i].replace(",", ""))
    u = units[i]
    if name.startswith("dram__bytes"):
        v *= {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)
    if name.startswith("gpu__time"):
        v *= {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
    return v
print(f"### {title}\n")
print("| kernel | " + " | ".join(w[0] for w in want) + " | GB/s |")
print("|---|" + "---|" * (len(want) + 1))
ki = col("Kernel Name")
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[ki].replace("void ", "").replace("<unnamed>::", "")
    name = name[:name.index("(")] if "(" in name else name
    vals = [val(r, w[1]) for w in want]
    t, rd, wr = vals[0], vals[1], vals[2]
    gbs = (rd + wr) * 1e6 / (t * 1e-6) / 1e9 if t and rd is not None and wr is not None else None
    fmt = lambda v: "-" if v is None else (f"{v:.0f}" if v >= 100 else f"{v:.1f}")
    print(f"| `{name}` | " + " | ".join(fmt(v) for v in vals) + f" | {fmt(gbs)} |")
