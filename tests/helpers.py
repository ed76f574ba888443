"""Shared test helpers: problem zoo (the reference's own test problems), tape
construction, host-emulation driver."""
import ctypes as C

import numpy as np


def zoo(H):
    """name -> (prob, alg, dtype): the problems of test/MLP.jl, some scaled down."""
    z = {}
    d = 10
    z["heat_d10"] = (H.ParabolicPDEProblem(0.0, 0.1, np.full(d, 2.0), (0.0, 0.5), g=H.SumSq(), f=H.FZero()),
                     H.MLP(), np.float64)  # test/MLP.jl:19-47
    z["heat_neumann_d5"] = (H.ParabolicPDEProblem(0.0, 0.1, np.full(5, 0.3), (0.0, 0.5), g=H.SumSq(), f=H.FZero(),
                                                  neumann_bc=(np.full(5, -0.5), np.full(5, 0.5))),
                            H.MLP(), np.float64)  # :80-105
    z["exponential_d3"] = (H.ParabolicPDEProblem(0.0, 0.0, np.zeros(3), (0.0, 0.5), g=H.SumSq(2.0), f=H.FLinear(0.1)),
                           H.MLP(), np.float64)  # :107-134
    z["allen_cahn_d10"] = (H.ParabolicPDEProblem(0.0, 1.0, np.zeros(10), (0.3, 0.6), g=H.InvSumSq(2.0, 0.4),
                                                 f=H.FAllenCahn(-1.0)), H.MLP(), np.float64)  # :136-160
    z["allen_cahn_neumann_d5"] = (H.ParabolicPDEProblem(0.0, 1.0, np.zeros(5), (0.0, 0.5), g=H.ExpSumSq(0.25),
                                                        f=H.FAllenCahn(1.0),
                                                        neumann_bc=(np.full(5, -0.5), np.full(5, 0.5))),
                                  H.MLP(), np.float64)  # :162-186
    z["black_scholes_d20"] = (H.ParabolicPDEProblem(H.LinearCoef(0.02), H.LinearCoef(0.2), np.full(20, 100.0),
                                                    (0.0, 1.0), g=H.MinCoord(), f=H.FDefaultRisk()),
                              H.MLP(), np.float64)  # :188-230
    z["replicator_d2"] = (H.PIDEProblem(0.0, 0.1, np.zeros(2), (0.0, 0.5), H.GaussDensity(1e-2),
                                        H.FNLReplicatorMutator(1.0)),
                          H.MLP(M=3, L=4, K=3, mc_sample=H.NormalSampling(1.0)), np.float64)  # :236-286 (smaller M,K)
    z["nonlocal_ac_neumann_d5"] = (H.PIDEProblem(0.0, 0.1, np.zeros(5), (0.0, 0.5), H.ExpSumSq(0.25),
                                                 H.FNLAllenCahnDiff(),
                                                 neumann_bc=(np.full(5, -0.5), np.full(5, 0.5))),
                                   H.MLP(M=3, L=3, K=4, mc_sample=H.UniformSampling(np.full(5, -0.5), np.full(5, 0.5))),
                                   np.float64)  # :288-317 (smaller M,L,K)
    z["fisher_nonlocal_d10"] = (H.PIDEProblem(0.0, 0.1, np.zeros(10), (0.0, 0.5), H.ExpSumSq(1.0), H.FNLFisher(),
                                              neumann_bc=(np.full(10, -0.5), np.full(10, 0.5))),
                                H.MLP(M=2, L=4, K=3, mc_sample=H.UniformSampling(-0.5, 0.5)),
                                np.float64)  # docs/src/tutorials/mlp.md:45-58 (smaller)
    z["shifted_normal_d3_L5"] = (H.PIDEProblem(0.0, 0.1, np.zeros(3), (0.0, 0.5), H.ExpSumSq(1.0), H.FNLFisher()),
                                 H.MLP(M=2, L=5, K=2, mc_sample=H.NormalSampling(0.2, True)), np.float64)
    z["heat_neumann_f32_d3"] = (H.ParabolicPDEProblem(0.0, 0.1, np.full(3, 0.3, dtype=np.float32), (0.0, 0.5),
                                                      g=H.SumSq(), f=H.FZero(),
                                                      neumann_bc=(np.full(3, -0.5), np.full(3, 0.5))),
                                H.MLP(), np.float32)
    z["allen_cahn_neumann_f32_d7"] = (H.ParabolicPDEProblem(0.0, 1.0, np.zeros(7, dtype=np.float32), (0.0, 0.5),
                                                            g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                                            neumann_bc=(-0.5, 0.5)), H.MLP(M=3, L=3, K=2), np.float32)
    z["odd_d_fisher_kpp_d7"] = (H.ParabolicPDEProblem(0.0, 0.1, np.zeros(7), (0.0, 0.5), g=H.ExpSumSq(1.0),
                                                      f=H.FFisherKPP()), H.MLP(M=3, L=4, K=2), np.float64)
    return z


def random_tapes(H, prob, alg, dtype, seed=1, compat=0):
    """numpy-generated tapes in the reference's consumption order."""
    rng = np.random.default_rng(seed)
    nl = H._integrate(alg.mc_sample)
    cn = H.mlp_counts(alg.M, alg.L, alg.K, nl, compat)
    d = prob.d
    r = rng.random(cn["time_uniforms"]).astype(dtype)
    n = rng.standard_normal((cn["path_steps"], d)).astype(dtype)
    s = None
    if nl:
        if isinstance(alg.mc_sample, H.NormalSampling):
            s = rng.standard_normal((cn["sampler_calls"], d)).astype(dtype)
        else:
            s = rng.random((cn["sampler_calls"], d)).astype(dtype)
    return r, n, s


def emu_solve(H, emu, prob, alg, dtype=np.float64, seed=0, tapes=None, target_units=0, compat=0, n_solves=1, nranks=1,
              std_err=False):
    pb, keep = prob.to_c()
    cfg = H.capi.hdp_mlp_cfg()
    cfg.M, cfg.L, cfg.K = alg.M, alg.L, alg.K
    cfg.n_solves, cfg.n_x0, cfg.seed = n_solves, 1, seed
    cfg.target_units, cfg.compat_mlt_threads = target_units, compat
    sm, k2 = alg.mc_sample._descriptor(prob.d)
    cfg.sampler = sm
    ts = []
    if tapes is not None:
        ts = [np.ascontiguousarray(t, dtype=dtype) if t is not None else None for t in tapes]
        cfg.r_tape = ts[0].ctypes.data
        cfg.n_tape = ts[1].ctypes.data
        if ts[2] is not None:
            cfg.s_tape = ts[2].ctypes.data
    x = np.ascontiguousarray(prob.x, dtype=dtype)
    u = np.zeros(n_solves, dtype=dtype)
    se = np.zeros(n_solves, dtype=dtype)
    if std_err:
        cfg.std_err_out = se.ctypes.data
    st = H.capi.hdp_mlp_stats()
    fn = emu.hdp_emu_mlp_solve_sharded_f64 if dtype == np.float64 else emu.hdp_emu_mlp_solve_sharded_f32
    rc = fn(C.byref(pb), C.byref(cfg), x.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p), C.byref(st), int(nranks))
    if std_err:
        st.se = se
    return u, rc, st


def oracle_solve(H, O, prob, alg, dtype, tapes):
    pb, keep = prob.to_c()
    sm, k2 = alg.mc_sample._descriptor(prob.d)
    rng = O.Rng.tape(dtype, *tapes)
    u, info = O.ml_picard(pb, sm, alg.M, alg.L, alg.K, np.asarray(prob.x, dtype=dtype), rng)
    info["tape_pos"] = rng.tape_pos()
    return u, info


def rel_err(a, b):
    return abs(a - b) / max(abs(b), 1e-300)
