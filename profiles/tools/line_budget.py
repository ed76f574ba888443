#!/usr/bin/env python3
"""Executed warp-instructions per CUDA source line of one kernel capture (ncu --set full --import-source on, built with
-lineinfo), as instructions per unit of work.  usage: [SRC_REV=<git rev of the captured build>] line_budget.py
report.ncu-rep units [top_n]   (`units` = warp-level units of work in the launch, e.g. vec-group passes =
solves*path_steps*vec_groups/32).  An instruction inlined from a helper is attributed to the helper's line."""
import collections
import csv
import io
import os
import subprocess
import sys

rep, units = sys.argv[1], float(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur, iE, iS = None, None, None
agg, samp = collections.Counter(), collections.Counter()
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2:
        if r[0] == "File Path":
            cur = r[1]
        continue
    if r and r[0] == "Line No":
        iE, iS = r.index("Instructions Executed"), r.index("# Samples")
        continue
    try:
        ln, e, s = int(r[0]), int(r[iE]), int(r[iS])
    except (ValueError, IndexError, TypeError):
        continue
    agg[(cur, ln)] += e
    samp[(cur, ln)] += s
tot, ts = sum(agg.values()), sum(samp.values())
print(f"total warp-instructions {tot} = {tot / units:.1f} per unit; stall samples {ts}")
cache = {}
REV = os.environ.get("SRC_REV")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def line(f, ln):
    if f not in cache:
        rel = f.split("/repo/", 1)[-1]
        if REV:
            cache[f] = subprocess.run(["git", "-C", ROOT, "show", f"{REV}:{rel}"], capture_output=True, text=True).stdout.splitlines()
        else:
            local = os.path.join(ROOT, rel)
            cache[f] = open(local).read().splitlines() if os.path.exists(local) else []
    return cache[f][ln - 1].strip()[:100] if 0 < ln <= len(cache[f]) else ""
print("file:line | instr/unit | % of samples | source")
for (f, ln), e in sorted(agg.items(), key=lambda kv: -kv[1])[:topn]:
    print(f"{os.path.basename(f)}:{ln} | {e / units:.2f} | {100 * samp[(f, ln)] / ts:.2f} | {line(f, ln)}")
