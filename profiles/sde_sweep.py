"""`sde_loop!` throughput: the BASELINE shape (constant coefficients, symmetric box: the fast mode) and the generic mode
(Black-Scholes coefficients of test/DeepSplitting.jl:470-471 without a boundary; vector bounds), fp32 and fp64."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hdpde_b200 as H
h = H.Handle(0)
d, N = 100, 10
cases = {"const_sym": lambda dt_: (H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d, dtype=dt_), (0.0, 1.0), g=H.SumSq(), f=H.FZero(), neumann_bc=(-1.0, 1.0)), 0.0),
         "linear_nobc": lambda dt_: (H.ParabolicPDEProblem(H.LinearCoef(0.02), H.LinearCoef(0.2), np.zeros(d, dtype=dt_), (0.0, 1.0), g=H.SumSq(), f=H.FZero()), 100.0),
         "const_vecbox": lambda dt_: (H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d, dtype=dt_), (0.0, 1.0), g=H.SumSq(), f=H.FZero(), neumann_bc=(np.linspace(-1.0, -0.5, d), np.linspace(0.5, 1.0, d))), 0.0)}
which = sys.argv[1:] or list(cases)
for dtype, tdt, fn in ((np.float32, torch.float32, "hdp_sde_loop_f32"), (np.float64, torch.float64, "hdp_sde_loop_f64")):
  for name in which:
    prob, fill = cases[name](dtype)
    pb, _k = prob.to_c()
    none, _k2 = H.NoSampling()._descriptor(d)
    for B in ((8192, 65536, 1 << 20) if name == "const_sym" else (1 << 20,)):
        y1 = torch.zeros(d * B, dtype=tdt, device="cuda"); y0 = torch.empty_like(y1)
        best = None
        for r in range(12):
            y1.fill_(fill); torch.cuda.synchronize()
            dt = C.c_float(0.1) if dtype == np.float32 else C.c_double(0.1)
            rc = getattr(h.lib, fn)(h.h, C.byref(pb), C.c_void_p(y0.data_ptr()), C.c_void_p(y1.data_ptr()), B, N, dt, None, C.byref(none), 7 + r, None, C.byref(none), None, 0, H.capi.HDP_NORMAL_DEFAULT, H.capi.HDP_MEM_DEVICE | H.capi.HDP_ASYNC)
            H.capi.check(h.h, rc); ms = h.last_device_ms()
            if r >= 2: best = ms if best is None else min(best, ms)
        h.sync()
        print(fn, name, B, f"{best*1e3:.1f} us", f"{B*N*d/best/1e6:.1f} G coord-steps/s", float(y1.abs().max()), flush=True)
