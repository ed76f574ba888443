"""Small workloads for compute-sanitizer (memcheck / racecheck / synccheck): the smoke test of __graft_entry__, one
non-local MLP solve with reflection at lane-group widths 1, 4, 8 and 32 (shuffles, per-thread frames, fix-up pass,
sampler walk), one local d=100 solve on the long-row kernel, a sharded-style cut (many units), `sde_loop!` with the
fused sampler epilogue, an EM ensemble and `_reflect`.
    compute-sanitizer --tool memcheck  python profiles/sanitize_workload.py
    compute-sanitizer --tool racecheck python profiles/sanitize_workload.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import __graft_entry__ as ge  # noqa: E402
import hdpde_b200 as H  # noqa: E402

ge.smoke()
h = H.Handle(0)
d = 5
nl = H.PIDEProblem(0.0, 0.6, np.zeros(d), (0.0, 0.5), H.ExpSumSq(0.25), H.FNLAllenCahnDiff(), neumann_bc=(-0.5, 0.5))
alg = H.MLP(M=3, L=3, K=3, mc_sample=H.UniformSampling(-0.5, 0.5))
ref = None
for G in (1, 4, 8, 32):
    for target in (1, 200):
        u, st = H._ml_picard_batch(nl, alg, n_solves=3, seed=7, handle=h, group_size=G, target_units=target, std_err=True)
        ref = u if ref is None else ref
        assert np.allclose(u, ref, rtol=1e-12), (G, target, u, ref)
        print("nonlocal G", G, "target", target, u, st["n_units"], st["bounces"], flush=True)
d = 100
c2 = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0), neumann_bc=(-0.5, 0.5))
for G, n in ((1, 8), (4, 2), (32, 1)):
    u, st = H._ml_picard_batch(c2, H.MLP(M=4, L=3, K=10), n_solves=n, seed=3, handle=h, group_size=G)
    print("C2-like L=3 G", G, u.mean(), st["n_units"], flush=True)
y1 = np.zeros((d, 64), dtype=np.float32, order="F")
y0 = np.empty_like(y1, order="F")
z = np.empty((d, 64, 2), dtype=np.float32, order="F")
p32 = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d, dtype=np.float32), (0.0, 1.0), g=H.SumSq(), f=H.FZero(), neumann_bc=(-1.0, 1.0))
H.sde_loop(p32, y0, y1, 10, 0.1, seed=1, mc_sample=H.NormalSampling(0.1, True), z=z, handle=h)
y3 = np.zeros((7, 9, 3), order="F")
H.sde_loop_epochs(H.ParabolicPDEProblem(0.0, 1.0, np.zeros(7), (0.0, 1.0), g=H.SumSq(), f=H.FZero(), neumann_bc=(-1.0, 1.0)),
                  np.empty_like(y3, order="F"), y3, 4, 0.1, seed=2, handle=h)
H.em_ensemble(H.ParabolicPDEProblem(H.LinearCoef(0.02), H.LinearCoef(0.2), np.zeros(6), (0.0, 0.5), g=H.SumSq(), f=H.FZero()),
              np.full(6, 100.0), 0.1, trajectories=50, saveat=True, seed=4, handle=h)
a = np.zeros((9, 33), order="F")
H._reflect(a, np.asfortranarray(a + np.random.default_rng(0).standard_normal(a.shape)), -0.5, 0.5, return_bounces=True, handle=h)
h.close()
print("sanitize workload: OK")
