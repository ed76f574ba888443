"""`MLP` algorithm struct and `solve` — mirror of src/MLP.jl:24-100.
The recursion itself (`_ml_picard*`, src/MLP.jl:102-388) runs in the K3 kernels
of libhdpde_b200.so; this file only packs descriptors and wraps the result in the
same `PIDESolution(x, [tspan...], nothing, [g(x), usol], nothing)` (:99).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi as capi
from .problems import PIDEProblem, ParabolicPDEProblem, PIDESolution, _Problem
from .sampling import NoSampling, _integrate


class MLP:
    """MLP(; M = 4, L = 4, K = 10, mc_sample = NoSampling()) (src/MLP.jl:24-32)."""

    def __init__(self, M: int = 4, L: int = 4, K: int = 10, mc_sample=None):
        self.M, self.L, self.K = int(M), int(L), int(K)
        self.mc_sample = NoSampling() if mc_sample is None else mc_sample


def mlp_counts(M, L, K, nonlocal_=False, compat_mlt_threads=0) -> dict:
    lib = capi.load_library()
    c = capi.hdp_mlp_counts()
    rc = lib.hdp_mlp_counts_get(M, L, K, int(nonlocal_), int(compat_mlt_threads), C.byref(c))
    if rc != capi.HDP_OK:
        raise capi.HdpError(rc, "hdp_mlp_counts_get")
    return {k: getattr(c, k) for k, _ in capi.hdp_mlp_counts._fields_}


def _ml_picard_batch(prob: _Problem, alg: MLP, *, n_solves=1, x0=None, seed=0, tapes=None,
                     compat_mlt_threads=0, target_units=0, handle=None, dtype=None, group_size=0,
                     unit_oversub=0.0, std_err=False):
    """Runs n_solves independent MLP solves (seeds seed, seed+1, …) on the GPU.
    x0: None (prob.x), a d-vector, or a d×n_solves column-major matrix.
    tapes: optional (r_tape, n_tape, s_tape) in the reference's consumption order.
    group_size / unit_oversub: launch tuning (0 = automatic; results do not depend on them).
    std_err: also estimate the Monte-Carlo standard error of every solve (stats["std_err"]).
    Returns (u[n_solves], stats dict)."""
    h = handle or capi.default_handle()
    if not isinstance(prob.x0_sample, NoSampling):
        # src/MLP.jl:65-66
        raise capi.HdpError(capi.HDP_ERR_X0_SAMPLE_NOT_SUPPORTED,
                            "`MLP` algorithm can only be used with `x0_sample=NoSampling()`.")
    x = prob.x if x0 is None else np.asarray(x0)
    dt = np.dtype(dtype) if dtype is not None else (x.dtype if x.dtype in (np.float32, np.float64) else np.dtype(np.float64))
    x = np.asarray(x, dtype=dt, order="F")
    d = prob.d
    if x.ndim == 1:
        n_x0 = 1
    else:
        n_x0 = x.shape[1]
    if x.shape[0] != d:
        raise ValueError("x0 has the wrong leading dimension")
    pb, keep = prob.to_c()
    cfg = capi.hdp_mlp_cfg()
    cfg.M, cfg.L, cfg.K = alg.M, alg.L, alg.K
    cfg.n_solves = int(n_solves)
    cfg.n_x0 = int(n_x0)
    cfg.normal_mode = capi.HDP_NORMAL_DEFAULT
    cfg.compat_mlt_threads = int(compat_mlt_threads)
    cfg.target_units = int(target_units)
    cfg.group_size = int(group_size)
    cfg.unit_oversub = float(unit_oversub)
    cfg.mem_flags = capi.HDP_MEM_HOST
    cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    sm, keep2 = alg.mc_sample._descriptor(d)
    cfg.sampler = sm
    ue = prob.user_exprs()
    if ue is not None:
        cfg.user = C.pointer(ue)
    tp = []
    if tapes is not None:
        r_t, n_t, s_t = tapes
        r_t = np.ascontiguousarray(r_t, dtype=dt)
        n_t = np.ascontiguousarray(n_t, dtype=dt)
        tp = [r_t, n_t]
        cfg.r_tape = r_t.ctypes.data
        cfg.n_tape = n_t.ctypes.data
        if s_t is not None:
            s_t = np.ascontiguousarray(s_t, dtype=dt)
            tp.append(s_t)
            cfg.s_tape = s_t.ctypes.data
    u = np.empty(int(n_solves), dtype=dt)
    se = None
    if std_err:
        se = np.zeros(int(n_solves), dtype=dt)
        cfg.std_err_out = se.ctypes.data
    st = capi.hdp_mlp_stats()
    fn = h.lib.hdp_mlp_solve_f64 if dt == np.float64 else h.lib.hdp_mlp_solve_f32
    rc = fn(h.h, C.byref(pb), C.byref(cfg), capi.np_ptr(x), capi.np_ptr(u), C.byref(st))
    capi.check(h.h, rc)
    stats = {"path_steps": st.per_solve.path_steps, "time_uniforms": st.per_solve.time_uniforms,
             "sampler_calls": st.per_solve.sampler_calls, "f_evals": st.per_solve.f_evals,
             "g_evals": st.per_solve.g_evals, "bounces": st.bounces, "n_units": st.n_units,
             "n_nodes": st.n_nodes, "n_sites": st.n_sites, "device_ms": st.device_ms,
             "group_size": st.group_size, "n_launches": st.n_launches,
             "phase_ms": {"expand": st.phase_ms[0], "units": st.phase_ms[1], "allreduce": st.phase_ms[2],
                          "combine": st.phase_ms[3]},
             "exchange_bytes": st.exchange_bytes,
             "exchange_kind": ("none", "nccl", "peer")[st.exchange_kind]}
    if std_err:
        L1 = alg.L + 1
        stats.update({"std_err": se, "level_n": list(st.level_n[:L1]), "level_sum": list(st.level_sum[:L1]),
                      "level_sumsq": list(st.level_sumsq[:L1])})
    del keep, keep2, tp
    return u, stats


def _ml_picard_device(prob: _Problem, alg: MLP, x0_ptr: int, u_ptr: int, *, n_solves=1, n_x0=1, seed=0, dtype=np.float64,
                      std_err_ptr: int = 0, async_=False, group_size=0, handle=None):
    """The same solve with device-resident arguments (Julia: `CuArray`s handed over as `CuPtr`): x0 (d×n_x0),
    u (n_solves) and the optional standard errors are raw device addresses; with async_=True the call only
    enqueues (finish with handle.sync(), which also reports a start point outside the box)."""
    h = handle or capi.default_handle()
    dt = np.dtype(dtype)
    pb, keep = prob.to_c()
    cfg = capi.hdp_mlp_cfg()
    cfg.M, cfg.L, cfg.K = alg.M, alg.L, alg.K
    cfg.n_solves, cfg.n_x0 = int(n_solves), int(n_x0)
    cfg.normal_mode = capi.HDP_NORMAL_DEFAULT
    cfg.group_size = int(group_size)
    cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    cfg.mem_flags = capi.HDP_MEM_DEVICE | (capi.HDP_ASYNC if async_ else 0)
    cfg.std_err_out = int(std_err_ptr) or None
    sm, keep2 = alg.mc_sample._descriptor(prob.d)
    cfg.sampler = sm
    ue = prob.user_exprs()
    if ue is not None:
        cfg.user = C.pointer(ue)
    st = capi.hdp_mlp_stats()
    fn = h.lib.hdp_mlp_solve_f64 if dt == np.float64 else h.lib.hdp_mlp_solve_f32
    capi.check(h.h, fn(h.h, C.byref(pb), C.byref(cfg), C.c_void_p(int(x0_ptr)), C.c_void_p(int(u_ptr)), C.byref(st)))
    del keep, keep2
    return st


def solve(prob, alg, *args, multithreading=True, verbose=False, seed=None, **kw):
    """SciMLBase.solve(prob, alg::MLP; multithreading = true, verbose = false) (src/MLP.jl:43-100).

    `multithreading=True` is the reference's threaded estimator `_ml_picard_mlt`; its
    `_get_loop_num` partition depends on `Threads.nthreads()`.  Pass
    `julia_threads=NT` to reproduce that partition exactly; by default the serial
    estimator `_ml_picard` (the unbiased one) is evaluated — on the GPU the two differ
    only in which sites are sampled, not in how fast.
    """
    if not isinstance(alg, MLP):
        raise TypeError(f"solve: unsupported algorithm {type(alg).__name__} (this package accelerates MLP's path)")
    if not isinstance(prob, (PIDEProblem, ParabolicPDEProblem)):
        raise TypeError("solve: prob must be a PIDEProblem or ParabolicPDEProblem")
    nt = int(kw.pop("julia_threads", 0))
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
        h = kw.get("handle") or capi.default_handle()
        if h.info()["nccl_nranks"] > 1:
            # a solve sharded over several ranks: every rank must walk the same random tree
            from . import dist
            seed = dist.broadcast_int(seed)
    u, stats = _ml_picard_batch(prob, alg, n_solves=1, seed=seed,
                                compat_mlt_threads=nt if multithreading else 0, **kw)
    usol = u[0].item()
    return PIDESolution(prob.x, [prob.tspan[0], prob.tspan[1]], None, [prob.g(prob.x), usol], None, stats)
