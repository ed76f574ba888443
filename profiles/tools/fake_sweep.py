"""Lane-group width / cut sweep of ONE rank's share of an N-rank sharded solve on a single GPU (trace build,
HDP_TRACE_FAKE_RANKS=r/N; see trace_fake.py).  Prints the median device time and phases per (workload, group_size,
unit_oversub).  Timings only: the exchange is left out, the value is rank r's share."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, hdpde_b200 as H
import bench
h = H.Handle(0)
W = bench.strong_workloads(H)
cases = [a.split(":") for a in sys.argv[1:]]
for c in cases:
    name = c[0]; kw = {}
    if len(c) > 1 and c[1]: kw["group_size"] = int(c[1])
    if len(c) > 2 and c[2]: kw["unit_oversub"] = float(c[2])
    prob, alg = W[name]
    rows = []
    for r in range(10):
        u, st = H._ml_picard_batch(prob, alg, n_solves=1, seed=31, handle=h, **kw)
        if r >= 3: rows.append((st["device_ms"], st["phase_ms"]))
    rows.sort(key=lambda t: t[0])
    med = rows[len(rows) // 2]
    print(name, kw, "median %.4f ms" % med[0], {k: round(v, 4) for k, v in med[1].items()}, "units", st["n_units"], "G", st["group_size"], flush=True)
