import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hdpde_b200 as H
h = H.Handle(0)
def run(name, prob, alg, n, gs):
    out = []
    for g in gs:
        best = 1e30
        for r in range(4):
            u, st = H._ml_picard_batch(prob, alg, n_solves=n, seed=100 + r * n, handle=h, group_size=g)
            if r: best = min(best, st["device_ms"])
        out.append(f"G={g}:{best:.3f}ms/{st['n_units']}u")
    print(f"{name:28s} n={n:5d}  " + "  ".join(out), flush=True)
d = 100
c2 = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0), neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
for n in (1, 8, 64, 256, 512, 2048):
    run("C2 d=100 L=4", c2, H.MLP(), n, (1, 2, 4, 8, 32))
d = 1000
c5 = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0), neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
for n in (1, 4, 16, 64):
    run("C5 d=1000 L=5", c5, H.MLP(M=4, L=5, K=10), n, (8, 16, 32))
run("C5b d=1000 L=5 M=8", c5, H.MLP(M=8, L=5, K=10), 1, (8, 16, 32))
d = 10
c1 = H.ParabolicPDEProblem(0.0, 0.1, np.zeros(d), (0.0, 0.5), g=H.SumSq(), f=H.FZero())
for n in (1, 16, 64, 512):
    run("C1 d=10 L=4", c1, H.MLP(), n, (1, 2, 4, 8))
c4 = H.PIDEProblem(0.0, 0.1, np.zeros(d), (0.0, 0.5), H.ExpSumSq(1.0), H.FNLFisher(), neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
run("C4 d=10 L=5 nonlocal", c4, H.MLP(M=4, L=5, K=10, mc_sample=H.UniformSampling(np.full(d, -0.5), np.full(d, 0.5))), 1, (1, 2, 4))
run("C4 d=10 L=4 nonlocal", c4, H.MLP(M=4, L=4, K=10, mc_sample=H.UniformSampling(np.full(d, -0.5), np.full(d, 0.5))), 1, (1, 2, 4, 8))
