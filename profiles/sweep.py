"""Throughput sweep over the BASELINE configurations (single GPU): python profiles/sweep.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import hdpde_b200 as H  # noqa: E402

h = H.Handle(0)
QUICK = os.environ.get("SWEEP_QUICK") == "1"  # one batch size per configuration


def run(name, prob, alg, n, reps=3, **kw):
    best_dev, best_wall = 1e30, 1e30
    for r in range(reps + 1):
        t0 = time.perf_counter()
        u, st = H._ml_picard_batch(prob, alg, n_solves=n, seed=100 + r * n, handle=h, **kw)
        w = (time.perf_counter() - t0) * 1e3
        if r > 0:
            best_dev, best_wall = min(best_dev, st["device_ms"]), min(best_wall, w)
    ps = st["path_steps"] * n
    print(f"{name:34s} n={n:6d} G={st['group_size']:2d} units={st['n_units']:7d} launches={st['n_launches']:2d} "
          f"dev={best_dev:9.3f} ms wall={best_wall:9.3f} ms  {n / best_dev * 1e3:12.1f} solves/s "
          f"{ps / best_dev * 1e3 / 1e6:10.1f} Mpath-steps/s  u={u.mean():.6g}", flush=True)


d = 10
c1 = H.ParabolicPDEProblem(0.0, 0.1, np.zeros(d), (0.0, 0.5), g=H.SumSq(), f=H.FZero())
for n in ((4096,) if QUICK else (1, 64, 4096, 65536)):
    run("C1 heat d=10 L=4", c1, H.MLP(), n)
d = 100
c2 = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                           neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
for n in ((512, 4096) if QUICK else (1, 8, 64, 512, 4096, 16384)):
    run("C2 Allen-Cahn d=100 Neumann L=4", c2, H.MLP(), n)
d = 10
c4 = H.PIDEProblem(0.0, 0.1, np.zeros(d), (0.0, 0.5), H.ExpSumSq(1.0), H.FNLFisher(),
                   neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
a4 = H.MLP(M=4, L=5, K=10, mc_sample=H.UniformSampling(np.full(d, -0.5), np.full(d, 0.5)))
for n in ((1,) if QUICK else (1, 4)):
    run("C4 nonlocal Fisher d=10 L=5 K=10", c4, a4, n)
rm = H.PIDEProblem(0.0, 0.1, np.zeros(d), (0.0, 0.5), H.GaussDensity(1e-2), H.FNLReplicatorMutator(1.0))
run("C4b replicator-mutator d=10 L=4", rm, H.MLP(M=4, L=4, K=10, mc_sample=H.NormalSampling(1.0)), 4)
d = 1000
c5 = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                           neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
for n in ((64,) if QUICK else (1, 64, 256)):
    run("C5 Allen-Cahn d=1000 Neumann L=5", c5, H.MLP(M=4, L=5, K=10), n)
run("C5b Allen-Cahn d=1000 L=5 M=8", c5, H.MLP(M=8, L=5, K=10), 1)
h.close()
