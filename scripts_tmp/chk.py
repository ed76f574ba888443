import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hdpde_b200 as H
h = H.Handle(0)
D = 100
prob2 = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(D), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                              neumann_bc=(np.full(D, -0.5), np.full(D, 0.5)))
u, st = H._ml_picard_batch(prob2, H.MLP(), n_solves=1024, seed=50, handle=h)
c, st2 = H._ml_picard_batch(prob2, H.MLP(), n_solves=4, seed=50, handle=h)
print(st["group_size"], st["n_units"], st2["group_size"], st2["n_units"])
print(u[:4], c, (u[:4] - c) / c)
for g in (1, 2, 4, 8, 16, 32):
    os.environ["HDP_MLP_GROUP"] = str(g)
    for tu in (0, 1, 100):
        c2, s3 = H._ml_picard_batch(prob2, H.MLP(), n_solves=4, seed=50, handle=h, target_units=tu)
        print(g, tu, s3["n_units"], np.max(np.abs(c2 - c) / np.abs(c)), s3["bounces"])
