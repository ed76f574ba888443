"""One-GPU single-solve latency of the strong-scaling workloads (bench.strong_workloads): median and minimum of
`device_ms` over REPS solves, with the phase times of the median run.  Run once per library build
(HDPDE_B200_LIB=...) for an A/B on one box."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, hdpde_b200 as H
import bench
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 15
h = H.Handle(0)
print("lib", os.environ.get("HDPDE_B200_LIB", "default"), flush=True)
for name, (prob, alg) in bench.strong_workloads(H).items():
    rows = []
    for r in range(reps + 3):
        u, st = H._ml_picard_batch(prob, alg, n_solves=1, seed=31, handle=h)
        if r >= 3: rows.append((st["device_ms"], st["phase_ms"]))
    rows.sort(key=lambda t: t[0])
    med = rows[len(rows) // 2]
    print(name, repr(float(u[0])), "median %.4f min %.4f ms" % (med[0], rows[0][0]), {k: round(v, 4) for k, v in med[1].items()},
          "launches", st["n_launches"], "G", st["group_size"], flush=True)
