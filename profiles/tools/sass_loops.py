#!/usr/bin/env python3
"""Static view of the hot loops of a kernel: for every backward branch in the SASS of one
function (cuobjdump -sass), prints the loop length and its opcode histogram.
usage: sass_loops.py file.o kernel-substring [min_len] [must-contain-opcode]"""
import collections
import re
import subprocess
import sys

obj, kern = sys.argv[1], sys.argv[2]
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 60
must = sys.argv[4] if len(sys.argv) > 4 else None
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, ins = None, []
for line in txt.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None or kern not in cur:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_idx = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)* (?:[UP!]*\w+, )?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_idx:
            loops.append((addr_idx[tgt], i))
fp64 = ("DFMA", "DMUL", "DADD", "DSETP")
for s, e in loops:
    n = e - s + 1
    if n < min_len:
        continue
    ops = []
    for _, t in ins[s:e + 1]:
        w = t.split()
        op = w[1] if w[0].startswith("@") else w[0]
        ops.append(op.split(".")[0])
    if must and must not in ops:
        continue
    h = collections.Counter(ops)
    nf = sum(h[k] for k in fp64)
    print(f"loop {ins[s][0]:#x}..{ins[e][0]:#x}: {n} instr, fp64 {nf}, ST {h['ST']+h['STG']}, LDS {h['LDS']}, UMOV {h['UMOV']}, LDL/STL {h['LDL']+h['STL']}")
    print("   ", ", ".join(f"{k} {v}" for k, v in h.most_common()))
