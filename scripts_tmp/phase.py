import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hdpde_b200 as H
h = H.Handle(0)
d = 1000
c5 = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0), neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
for g in (8, 16, 32):
    for tu in (0, 30000, 10000):
        os.environ["HDP_MLP_GROUP"] = str(g)
        for r in range(3):
            u, st = H._ml_picard_batch(c5, H.MLP(M=8, L=5, K=10), n_solves=1, seed=100 + r, handle=h, target_units=tu)
        print(g, tu, st["n_units"], st["n_sites"], st["n_nodes"], round(st["device_ms"], 3), {k: round(v, 3) for k, v in st["phase_ms"].items()}, flush=True)
