"""`_reflect` and the DeepSplitting `sde_loop!` — mirrors of src/reflect.jl:8-82
and src/DeepSplitting.jl:153-166, served by the K4 / K1 kernels."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from .sampling import NoSampling


def _reflect(a, b, s, e, *, return_bounces=False, handle=None):
    """_reflect(a, b, s, e): reflection of the segment a→b on the hypercube [s, e]
    (src/reflect.jl).  a, b: d-vectors (Vector method, :8-55) or d×B column-major
    matrices (AbstractArray method, :58-82); s, e scalars or length-d vectors.
    Raises StartOutsideDomain when a is not inside (:12, :59)."""
    h = handle or capi.default_handle()
    a = np.asarray(a)
    dt = a.dtype if a.dtype in (np.float32, np.float64) else np.dtype(np.float64)
    a = np.asarray(a, dtype=dt, order="F")
    b = np.asarray(b, dtype=dt, order="F")
    if a.shape != b.shape:
        raise ValueError("a not same dim as b")  # :13, :60
    d = a.shape[0]
    B = 1 if a.ndim == 1 else int(np.prod(a.shape[1:]))
    scalar = np.ndim(s) == 0 and np.ndim(e) == 0
    if scalar:
        lo = np.array([s], dtype=dt)
        hi = np.array([e], dtype=dt)
    else:
        lo = np.ascontiguousarray(np.broadcast_to(np.asarray(s, dtype=dt), (d,)))
        hi = np.ascontiguousarray(np.broadcast_to(np.asarray(e, dtype=dt), (d,)))
    out = np.empty_like(a, order="F")
    cnt = np.empty(a.shape, dtype=np.int32, order="F") if return_bounces else None
    fn = h.lib.hdp_reflect_f64 if dt == np.float64 else h.lib.hdp_reflect_f32
    capi.check(h.h, fn(h.h, capi.np_ptr(a), capi.np_ptr(b), capi.np_ptr(out), d, B, capi.np_ptr(lo),
                       capi.np_ptr(hi), int(scalar), capi.np_ptr(cnt), capi.HDP_MEM_HOST))
    return (out, cnt) if return_bounces else out


def _set_user_coefs(h, prob):
    """hdp_set_user_exprs with the problem's μ/σ expressions (HDP_COEF_USER), cleared otherwise."""
    ue = prob.user_exprs()
    capi.check(h.h, h.lib.hdp_set_user_exprs(h.h, C.byref(ue) if ue is not None else None))
    return ue


def sde_loop(prob, y0, y1, nsteps, dt, *, ts=None, dWall=None, seed=0, mc_sample=None, z=None, handle=None):
    """sde_loop!(y0, y1, dWall) (src/DeepSplitting.jl:153-166): `nsteps` reflected
    Euler–Maruyama steps of the d×B batch.  y1 holds x0_batch on entry (:182); on
    return y0 is the state before the last step and y1 the final state.
    dWall: optional injected d×B×nsteps standard normals (the reference's array after
    randn!, :154); by default the normals are Philox draws generated in registers.
    mc_sample / z: optional fused `mc_sample!(z, y1)` (:186-189), z is d×B×K.
    ts: the time passed to μ, σ at each step (Julia's ts[N+1-i], :157); required for UserCoef kinds."""
    h = handle or capi.default_handle()
    dtp = y1.dtype
    if dtp not in (np.float32, np.float64) or y0.dtype != dtp:
        raise TypeError("y0, y1 must both be float32 or float64")
    for arr in (y0, y1):
        if arr.ndim != 2 or not arr.flags.f_contiguous:
            raise ValueError("y0, y1 must be d×B column-major (order='F') matrices")
    d, B = y1.shape
    if d != prob.d or y0.shape != y1.shape:
        raise ValueError("shape mismatch")
    pb, keep = prob.to_c()
    x0s, keep2 = prob.x0_sample._descriptor(d)
    mcs, keep3 = (mc_sample or NoSampling())._descriptor(d)
    K = 0
    if z is not None:
        if z.dtype != dtp or not z.flags.f_contiguous or z.shape[:2] != (d, B):
            raise ValueError("z must be d×B×K column-major of the same element type")
        K = int(np.prod(z.shape[2:]))
    dw = None
    if dWall is not None:
        dw = np.asarray(dWall, dtype=dtp, order="F")
        if dw.shape != (d, B, nsteps):
            raise ValueError("dWall must be d×B×nsteps")
    tsa = None if ts is None else np.ascontiguousarray(ts, dtype=dtp)
    if tsa is not None and tsa.shape != (nsteps,):
        raise ValueError("ts must hold nsteps times")
    ue = _set_user_coefs(h, prob)
    if dtp == np.float64:
        fn, dtc = h.lib.hdp_sde_loop_f64, C.c_double(dt)
    else:
        fn, dtc = h.lib.hdp_sde_loop_f32, C.c_float(dt)
    capi.check(h.h, fn(h.h, C.byref(pb), capi.np_ptr(y0), capi.np_ptr(y1), B, int(nsteps), dtc,
                       capi.np_ptr(tsa), C.byref(x0s), int(seed), capi.np_ptr(dw), C.byref(mcs),
                       capi.np_ptr(z), K, capi.HDP_NORMAL_DEFAULT, capi.HDP_MEM_HOST))
    del keep, keep2, keep3, ue
    return y0, y1


def sde_loop_epochs(prob, y0, y1, nsteps, dt, *, seed=0, mc_sample=None, z=None, device_ptrs=None, B=None, n_epochs=None,
                    dtype=None, async_=False, handle=None):
    """E epochs of `y1 .= x0_batch; sde_loop!(y0, y1, dWall)` (src/DeepSplitting.jl:181-184) in one launch: y0, y1 are
    d×B×E column-major (y1 holds every epoch's start points), z optionally d×B×K×E; slice e is what
    sde_loop(..., seed=seed + e) returns.  With device_ptrs=(y0_ptr, y1_ptr, z_ptr) the arrays are device-resident
    (give B, n_epochs, dtype) and nothing crosses PCIe."""
    h = handle or capi.default_handle()
    pb, keep = prob.to_c()
    x0s, keep2 = prob.x0_sample._descriptor(prob.d)
    mcs, keep3 = (mc_sample or NoSampling())._descriptor(prob.d)
    ue = _set_user_coefs(h, prob)
    if device_ptrs is None:
        dtp = y1.dtype
        if y1.ndim != 3 or not y1.flags.f_contiguous or y0.shape != y1.shape or y0.dtype != dtp or not y0.flags.f_contiguous:
            raise ValueError("y0, y1 must be d×B×E column-major arrays of one element type")
        d, B, n_epochs = y1.shape
        K = 0
        if z is not None:
            if z.dtype != dtp or not z.flags.f_contiguous or z.ndim != 4 or z.shape[:2] != (d, B) or z.shape[3] != n_epochs:
                raise ValueError("z must be d×B×K×E column-major")
            K = z.shape[2]
        p0, p1, pz, flags = capi.np_ptr(y0), capi.np_ptr(y1), capi.np_ptr(z), capi.HDP_MEM_HOST
    else:
        dtp = np.dtype(dtype)
        vp = lambda a: C.c_void_p(int(a)) if a else None
        p0, p1, pz = (vp(a) for a in device_ptrs)
        K = int(z or 0)  # with device pointers `z` carries K
        flags = capi.HDP_MEM_DEVICE | (capi.HDP_ASYNC if async_ else 0)
    fn, dtc = (h.lib.hdp_sde_loop_epochs_f64, C.c_double(dt)) if dtp == np.float64 else (h.lib.hdp_sde_loop_epochs_f32, C.c_float(dt))
    capi.check(h.h, fn(h.h, C.byref(pb), p0, p1, int(B), int(nsteps), dtc, None, C.byref(x0s), int(seed), int(n_epochs),
                       C.byref(mcs), pz, K, flags))
    del keep, keep2, keep3, ue
    return y0, y1


def sde_loop_device(prob, y0_ptr: int, y1_ptr: int, B: int, nsteps: int, dt, *, dtype=np.float32, seed=0, dWall_ptr: int = 0,
                    mc_sample=None, z_ptr: int = 0, K: int = 0, async_=False, handle=None):
    """`sde_loop!` on device-resident arrays (Julia: `CuArray`s handed over as `CuPtr`): y0, y1 (d×B), the optional
    injected dWall (d×B×nsteps) and z (d×B×K) are raw device addresses; nothing crosses PCIe.  With async_=True
    the call only enqueues on the handle's stream (finish with handle.sync())."""
    h = handle or capi.default_handle()
    dtp = np.dtype(dtype)
    pb, keep = prob.to_c()
    x0s, keep2 = prob.x0_sample._descriptor(prob.d)
    mcs, keep3 = (mc_sample or NoSampling())._descriptor(prob.d)
    ue = _set_user_coefs(h, prob)
    fn, dtc = (h.lib.hdp_sde_loop_f64, C.c_double(dt)) if dtp == np.float64 else (h.lib.hdp_sde_loop_f32, C.c_float(dt))
    flags = capi.HDP_MEM_DEVICE | (capi.HDP_ASYNC if async_ else 0)
    vp = lambda a: C.c_void_p(int(a)) if a else None
    capi.check(h.h, fn(h.h, C.byref(pb), vp(y0_ptr), vp(y1_ptr), int(B), int(nsteps), dtc, None, C.byref(x0s), int(seed),
                       vp(dWall_ptr), C.byref(mcs), vp(z_ptr), int(K), capi.HDP_NORMAL_DEFAULT, flags))
    del keep, keep2, keep3, ue


def _mlt_sde_loop(x2, x, s, t, prob, *, normals=None, seed=0, handle=None):
    """_mlt_sde_loop!(x2, x, s, t, prob, neumann_bc) (src/MLP.jl:406-421): one Euler–Maruyama step
    s -> t of the MLP recursion (x2 = x + (μ*dt + σ*sqrt(dt)*ξ), then `_reflect`), mutating x2.
    x, x2: d-vectors or d×B column-major matrices; normals: optional injected ξ (`randn!(x2)`)."""
    h = handle or capi.default_handle()
    dt_ = x.dtype
    if dt_ not in (np.float32, np.float64) or x2.dtype != dt_ or x2.shape != x.shape:
        raise TypeError("x, x2 must have the same float32/float64 dtype and shape")
    xa = np.asarray(x, order="F")
    if not (x2.flags.f_contiguous or x2.ndim == 1):
        raise ValueError("x2 must be column-major")
    d = x.shape[0]
    B = 1 if x.ndim == 1 else int(np.prod(x.shape[1:]))
    pb, keep = prob.to_c()
    inj = None if normals is None else np.asarray(normals, dtype=dt_, order="F")
    ue = _set_user_coefs(h, prob)
    if dt_ == np.float64:
        fn, a, b = h.lib.hdp_mlt_sde_loop_f64, C.c_double(s), C.c_double(t)
    else:
        fn, a, b = h.lib.hdp_mlt_sde_loop_f32, C.c_float(s), C.c_float(t)
    capi.check(h.h, fn(h.h, C.byref(pb), capi.np_ptr(x2), capi.np_ptr(xa), B, a, b, int(seed), capi.np_ptr(inj),
                       capi.HDP_MEM_HOST))
    del keep, ue
    return x2


def em_ensemble(prob, x0, dt, *, trajectories=None, saveat=False, normals=None, seed=0, handle=None):
    """`solve(EnsembleProblem(SDEProblem(μ, σ, x0, tspan)), EM(); dt, trajectories, adaptive=false)` — the
    pre-simulation stage of NNKolmogorov (src/NNKolmogorov.jl:81-112, final states only) and NNStopping
    (src/NNStopping.jl:90-99, `saveat = dt`).  x0: a d-vector (every trajectory starts there; give
    `trajectories`) or a d×B column-major matrix (trajectory b starts at x0[:, b], the `prob_func` of
    NNKolmogorov).  Returns the d×B final states, or with saveat=True the d×(N+1)×B array `Array(sim)`.
    normals: optional injected d×B×N standard normals."""
    h = handle or capi.default_handle()
    x0 = np.asarray(x0)
    dtp = x0.dtype if x0.dtype in (np.float32, np.float64) else np.dtype(np.float64)
    if x0.ndim == 1:
        if trajectories is None:
            raise ValueError("trajectories is required when x0 is a vector")
        x = np.asfortranarray(np.repeat(np.asarray(x0, dtype=dtp)[:, None], int(trajectories), axis=1))
    else:
        x = np.array(x0, dtype=dtp, order="F")
    d, B = x.shape
    if d != prob.d:
        raise ValueError("shape mismatch")
    t0, t1 = prob.tspan
    nsteps = int(round((t1 - t0) / dt))
    pb, keep = prob.to_c()
    inj = None
    if normals is not None:
        inj = np.asarray(normals, dtype=dtp, order="F")
        if inj.shape != (d, B, nsteps):
            raise ValueError("normals must be d×B×nsteps")
    traj = np.empty((d, nsteps + 1, B), dtype=dtp, order="F") if saveat else None
    ue = _set_user_coefs(h, prob)
    if dtp == np.float64:
        fn, dtc, t0c = h.lib.hdp_em_ensemble_f64, C.c_double(dt), C.c_double(t0)
    else:
        fn, dtc, t0c = h.lib.hdp_em_ensemble_f32, C.c_float(dt), C.c_float(t0)
    capi.check(h.h, fn(h.h, C.byref(pb), capi.np_ptr(x), B, nsteps, dtc, t0c, capi.np_ptr(traj), int(seed),
                       capi.np_ptr(inj), capi.HDP_MEM_HOST))
    del keep, ue
    return traj if saveat else x
