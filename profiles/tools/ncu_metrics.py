#!/usr/bin/env python3
"""Reads a committed `ncu -i report.ncu-rep --page raw --csv` export (header row, unit row, one value row per
kernel) and returns metrics WITH their units applied — the step where round 1 read 2.24 Gbyte as 2.24 MB.
usage (CLI): ncu_metrics.py file.csv [metric ...]"""
import csv
import glob
import os
import sys

_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
          "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def load(path, row=0):
    """dict: metric name -> value in base units (bytes, seconds; everything else as printed)."""
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    hdr, units, vals = rows[0], rows[1], rows[2 + row]
    out = {}
    for h, u, v in zip(hdr, units, vals):
        try:
            x = float(v.replace(",", ""))
        except ValueError:
            out[h] = v
            continue
        out[h] = x * _SCALE.get(u, 1.0)
        out[h + "#unit"] = u
    return out


def latest(pattern):
    """The newest committed export matching `pattern` under profiles/ (names carry the round: r1o < r2a < r2b)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = sorted(glob.glob(os.path.join(here, pattern)))
    return files[-1] if files else None


def dram_bytes(path):
    m = load(path)
    return m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]


if __name__ == "__main__":
    m = load(sys.argv[1])
    for k in (sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                               "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]):
        print(k, m.get(k), m.get(k + "#unit"))
