import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, hdpde_b200 as H
import bench
h = H.Handle(0)
W = bench.strong_workloads(H)
for name in W:
    prob, alg = W[name]
    for r in range(4):
        if r == 3: os.environ["HDP_PHASE_TRACE"] = "1"
        u, st = H._ml_picard_batch(prob, alg, n_solves=1, seed=31, handle=h)
        os.environ.pop("HDP_PHASE_TRACE", None)
    print(name, float(u[0]), round(st['device_ms'], 3), {k: round(v, 3) for k, v in st['phase_ms'].items()}, st['n_units'], 'G', st['group_size'], flush=True)
