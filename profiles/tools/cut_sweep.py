import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, hdpde_b200 as H
import bench
h = H.Handle(0)
W = bench.strong_workloads(H)
def run(name, **kw):
    prob, alg = W[name]
    best = None
    for r in range(4):
        u, st = H._ml_picard_batch(prob, alg, n_solves=1, seed=31, handle=h, **kw)
        if r > 0 and (best is None or st['device_ms'] < best['device_ms']): best = st
    ph = {k: round(v, 3) for k, v in best['phase_ms'].items()}
    print(name, kw, float(u[0]), round(best['device_ms'], 3), ph, best['n_units'], 'G', best['group_size'], 'L', best['n_launches'], flush=True)
for t in (0, 400000, 800000, 1100000, 3000000, 6000000):
    run("c4_nonlocal_fisher_d10_L5_one_solve", target_units=t)
for g in (0, 1, 2, 4, 8):
    for t in (0, 200000, 800000, 2000000):
        run("allen_cahn_d100_M8_L6_one_solve", group_size=g, target_units=t)
for g in (0, 2, 4, 8, 16):
    run("allen_cahn_d100_M8_L5_one_solve", group_size=g)
