#!/usr/bin/env python3
"""Compact summary of one .ncu-rep (first kernel): headline metrics, stall breakdown and the
instructions with the most stall samples.  usage: ncu_summary.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct"]
print("kernel:", m.get("Kernel Name"))
for k in keys:
    if k in m:
        print(f"  {k:78s} {m[k]}")
print("  stalls per issue (smsp__average_warps_issue_stalled_*_per_issue_active):")
st = [(float(v), h.split("stalled_")[1].split("_per_issue")[0]) for h, v in m.items()
      if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
print("   ", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True) if v >= 0.05))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print(f"  total samples {tot}; top instructions by samples:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:topn]:
    reasons = sorted(((int(r[ix[h]]), h) for h in hdr if h.startswith("stall_") and "Not Issued" not in h), reverse=True)[:2]
    print(f"    {r[ix['# Samples']]:>6s}  x{r[ix['Instructions Executed']]:>9s}  {r[1].strip()[:58]:58s} " + ", ".join(f"{h[6:]} {v}" for v, h in reasons))
