"""Small fixed workloads for ncu captures (see profiles/README.md).
usage: python profiles/prof_workload.py mlp [n_solves] | c1 [n_solves] | sde [B] | ens [B] | c4 | c5"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import hdpde_b200 as H  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "mlp"
h = H.Handle(0)
if what == "mlp":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    d = 100
    prob = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                 neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
    for it in range(3):
        u, st = H._ml_picard_batch(prob, H.MLP(), n_solves=n, seed=11 + it, handle=h)
        print(it, u.mean(), st)
elif what == "c1":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    d = 10
    prob = H.ParabolicPDEProblem(0.0, 0.1, np.zeros(d), (0.0, 0.5), g=H.SumSq(), f=H.FZero())
    for it in range(3):
        u, st = H._ml_picard_batch(prob, H.MLP(), n_solves=n, seed=11 + it, handle=h)
        print(it, u.mean(), st)
elif what == "c4":
    d = 10
    prob = H.PIDEProblem(0.0, 0.1, np.zeros(d), (0.0, 0.5), H.ExpSumSq(1.0), H.FNLFisher(),
                         neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
    Lv = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    alg = H.MLP(M=4, L=Lv, K=10, mc_sample=H.UniformSampling(np.full(d, -0.5), np.full(d, 0.5)))
    tu = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    for it in range(2):
        u, st = H._ml_picard_batch(prob, alg, n_solves=1, seed=3 + it, handle=h, target_units=tu)
        print(it, u, st)
elif what == "c5":
    d = 1000
    prob = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                 neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    for it in range(2):
        u, st = H._ml_picard_batch(prob, H.MLP(M=4, L=5, K=10), n_solves=n, seed=5 + it, handle=h)
        print(it, u.mean(), st)
elif what == "ens":
    import torch
    B = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 18)
    d, N = 100, 10
    prob = H.ParabolicPDEProblem(H.LinearCoef(0.02), H.LinearCoef(0.2), np.zeros(d), (0.0, 1.0), g=H.SumSq(), f=H.FZero())
    pb, _k = prob.to_c()
    x = torch.empty(d * B, dtype=torch.float64, device="cuda")
    traj = torch.empty(d * B * (N + 1), dtype=torch.float64, device="cuda")
    for it in range(3):
        x.fill_(100.0)
        torch.cuda.synchronize()
        rc = h.lib.hdp_em_ensemble_f64(h.h, C.byref(pb), C.c_void_p(x.data_ptr()), B, N, C.c_double(0.1), C.c_double(0.0),
                                       C.c_void_p(traj.data_ptr()), 7 + it, None, H.capi.HDP_MEM_DEVICE | H.capi.HDP_ASYNC)
        H.capi.check(h.h, rc)
        print(it, h.last_device_ms(), "ms")
else:
    import torch
    B = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 20)
    d, N = 100, 10
    prob = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d, dtype=np.float32), (0.0, 1.0), g=H.SumSq(), f=H.FZero(),
                                 neumann_bc=(-1.0, 1.0))
    pb, _k = prob.to_c()
    none, _k2 = H.NoSampling()._descriptor(d)
    y1 = torch.zeros(d * B, dtype=torch.float32, device="cuda")
    y0 = torch.empty_like(y1)
    for it in range(3):
        y1.zero_()
        torch.cuda.synchronize()
        rc = h.lib.hdp_sde_loop_f32(h.h, C.byref(pb), C.c_void_p(y0.data_ptr()), C.c_void_p(y1.data_ptr()), B, N,
                                    C.c_float(0.1), None, C.byref(none), 7 + it, None, C.byref(none), None, 0, 0,
                                    H.capi.HDP_MEM_DEVICE | H.capi.HDP_ASYNC)
        H.capi.check(h.h, rc)
        print(it, h.last_device_ms(), "ms")
h.close()
