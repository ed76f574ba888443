"""ctypes wrapper of the CPU oracle (oracle/hdp_oracle.c) — TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import this module, and only as the checker / CPU baseline.
The product package (`highdimpde.jl_b200/`) never does.

Descriptors: the oracle reads the same POD structs as the C ABI
(include/hdpde.h: hdp_problem, hdp_sampler); pass `ctypes.byref(struct)`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libhdp_oracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("hdp_oracle.c", "hdp_oracle_impl.h", "hdp_oracle.h")]
    src.append(os.path.join(_HERE, "..", "include", "hdpde.h"))
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src if os.path.exists(s))):
        return LIB_PATH
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, i, i64p, u64 = C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_uint64
        L.hdp_oracle_rng_tape.restype = vp
        L.hdp_oracle_rng_tape.argtypes = [i, vp, u64, vp, u64, vp, u64]
        L.hdp_oracle_rng_tape_pos.argtypes = [vp, C.POINTER(u64)]
        L.hdp_oracle_rng_philox.restype = vp
        L.hdp_oracle_rng_philox.argtypes = [u64, i, i]
        L.hdp_oracle_rng_xoshiro.restype = vp
        L.hdp_oracle_rng_xoshiro.argtypes = [u64, i, i]
        L.hdp_oracle_rng_free.argtypes = [vp]
        L.hdp_oracle_rng_fill.argtypes = [vp, i, vp, i]
        L.hdp_oracle_philox4x32_10.argtypes = [vp, vp, vp]
        L.hdp_oracle_philox4x32.argtypes = [vp, vp, i, vp]
        L.hdp_oracle_philox_rounds.restype = i
        L.hdp_oracle_draw_f64.restype = C.c_double
        L.hdp_oracle_draw_f64.argtypes = [u64, i, u64, u64, i]
        L.hdp_oracle_draw_f32.restype = C.c_float
        L.hdp_oracle_draw_f32.argtypes = [u64, i, u64, u64, i]
        L.hdp_oracle_mlp_counts.argtypes = [i, i, i, i, i64p]
        for sfx in ("_f64", "_f32"):
            real = C.c_double if sfx == "_f64" else C.c_float
            getattr(L, "hdp_oracle_reflect_vector" + sfx).argtypes = [vp, vp, vp, vp, i, vp, vp, i64p]
            getattr(L, "hdp_oracle_reflect_array" + sfx).argtypes = [vp, vp, vp, vp, i, i, i, vp, vp, vp, C.POINTER(C.c_int32)]
            getattr(L, "hdp_oracle_mc_sample" + sfx).argtypes = [vp, vp, vp, vp, i, i, i]
            f = getattr(L, "hdp_oracle_get_loop_num" + sfx)
            f.restype = C.c_int64
            f.argtypes = [i, C.c_int64, i, i]
            getattr(L, "hdp_oracle_ml_picard" + sfx).argtypes = [vp, vp, i, i, i, vp, vp, vp, i64p]
            getattr(L, "hdp_oracle_ml_picard_mlt" + sfx).argtypes = [vp, vp, i, i, i, vp, i, C.POINTER(vp), i, vp, i64p]
            getattr(L, "hdp_oracle_sde_loop" + sfx).argtypes = [vp, vp, vp, i, i, real, vp, vp, vp, vp, vp]
        L.hdp_oracle_bench_solves_f64.argtypes = [vp, vp, i, i, i, vp, i, i, u64, vp, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _sfx(dt):
    dt = np.dtype(dt)
    if dt == np.float64:
        return "_f64"
    if dt == np.float32:
        return "_f32"
    raise TypeError("float32 or float64 only")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Rng:
    """Random source consumed in the reference's order (see hdp_oracle.h)."""

    def __init__(self, ptr, keep=()):
        self.ptr = ptr
        self._keep = keep

    @staticmethod
    def tape(dtype, r, n, s=None):
        dt = np.dtype(dtype)
        r = np.ascontiguousarray(r, dtype=dt).ravel()
        n = np.ascontiguousarray(n, dtype=dt).ravel()
        s = np.ascontiguousarray(s if s is not None else np.zeros(0), dtype=dt).ravel()
        ptr = lib().hdp_oracle_rng_tape(int(dt == np.float32), _p(r), r.size, _p(n), n.size, _p(s), s.size)
        return Rng(ptr, (r, n, s))

    @staticmethod
    def philox(seed, dtype=np.float64, sampler_normal=False):
        return Rng(lib().hdp_oracle_rng_philox(int(seed), int(np.dtype(dtype) == np.float32), int(sampler_normal)))

    @staticmethod
    def xoshiro(seed, dtype=np.float64, sampler_normal=False):
        return Rng(lib().hdp_oracle_rng_xoshiro(int(seed), int(np.dtype(dtype) == np.float32), int(sampler_normal)))

    def tape_pos(self):
        out = (C.c_uint64 * 4)()
        lib().hdp_oracle_rng_tape_pos(self.ptr, out)
        return {"r": out[0], "n": out[1], "s": out[2], "overrun": bool(out[3])}

    def fill(self, what: int, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.float64)
        lib().hdp_oracle_rng_fill(self.ptr, what, _p(out), n)
        return out

    def __del__(self):
        try:
            if self.ptr:
                lib().hdp_oracle_rng_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().hdp_oracle_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def philox4x32(ctr, key, rounds=None):
    """Philox4x32-R; rounds=None: the library's round count (HDP_PHILOX_ROUNDS, include/hdpde.h)."""
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().hdp_oracle_philox4x32(_p(c), _p(k), philox_rounds() if rounds is None else int(rounds), _p(out))
    return out


def philox_rounds():
    return int(lib().hdp_oracle_philox_rounds())


def draw(dtype, seed, stream, ord_, i, normal):
    if np.dtype(dtype) == np.float64:
        return lib().hdp_oracle_draw_f64(seed, stream, ord_, i, int(normal))
    return lib().hdp_oracle_draw_f32(seed, stream, ord_, i, int(normal))


def philox_tape(dtype, seed, stream, ord0, n_ord, width, normal):
    """CPU twin of hdp_tape_fill_* (libm transcendentals; not bit-equal to the device's)."""
    out = np.empty((n_ord, width), dtype=dtype)
    for o in range(n_ord):
        for i in range(width):
            out[o, i] = draw(dtype, seed, stream, ord0 + o, i, normal)
    return out


def mlp_counts(M, L, K, nonlocal_):
    out = (C.c_int64 * 5)()
    lib().hdp_oracle_mlp_counts(M, L, K, int(nonlocal_), out)
    return {"path_steps": out[0], "time_uniforms": out[1], "sampler_calls": out[2], "f_evals": out[3], "g_evals": out[4]}


def get_loop_num(M, num, thread_id, nthreads):
    return lib().hdp_oracle_get_loop_num_f64(M, num, thread_id, nthreads)


def reflect_vector(a, b, s, e, return_bounces=False):
    """_reflect, Vector method (src/reflect.jl:8-55)."""
    a = np.ascontiguousarray(a)
    dt = a.dtype
    b = np.ascontiguousarray(b, dtype=dt)
    d = a.shape[0]
    s = np.ascontiguousarray(np.broadcast_to(np.asarray(s, dtype=dt), (d,)))
    e = np.ascontiguousarray(np.broadcast_to(np.asarray(e, dtype=dt), (d,)))
    out = np.empty_like(a)
    cnt = np.zeros(d, dtype=np.int32)
    tot = C.c_int64(0)
    rc = getattr(lib(), "hdp_oracle_reflect_vector" + _sfx(dt))(_p(a), _p(b), _p(s), _p(e), d, _p(out), _p(cnt), C.byref(tot))
    if rc != 0:
        raise ValueError(f"oracle reflect_vector rc={rc} (a not in hypercube)")
    return (out, cnt) if return_bounces else out


def reflect_array(a, b, s, e, return_info=False):
    """_reflect, AbstractArray method (src/reflect.jl:58-82); a, b d×B column-major."""
    a = np.asarray(a)
    dt = a.dtype
    a = np.asfortranarray(a)
    b = np.asfortranarray(b, dtype=dt)
    d = a.shape[0]
    B = 1 if a.ndim == 1 else int(np.prod(a.shape[1:]))
    scalar = np.ndim(s) == 0 and np.ndim(e) == 0
    if scalar:
        sv, ev = np.array([s], dtype=dt), np.array([e], dtype=dt)
    else:
        sv = np.ascontiguousarray(np.broadcast_to(np.asarray(s, dtype=dt), (d,)))
        ev = np.ascontiguousarray(np.broadcast_to(np.asarray(e, dtype=dt), (d,)))
    out = np.empty_like(a, order="F")
    cnt = np.zeros(a.shape, dtype=np.int32, order="F")
    col = np.zeros(B, dtype=np.int32)
    it = C.c_int32(0)
    rc = getattr(lib(), "hdp_oracle_reflect_array" + _sfx(dt))(_p(a), _p(b), _p(sv), _p(ev), int(scalar), d, B,
                                                                 _p(out), _p(cnt), _p(col), C.byref(it))
    if rc != 0:
        raise ValueError(f"oracle reflect_array rc={rc} (a not in hypercube)")
    return (out, cnt, col, it.value) if return_info else out


def mc_sample(sampler_struct, raw, y1=None):
    """Applies a sampler (hdp_sampler struct) to raw draws; raw is d×B×K."""
    raw = np.asfortranarray(raw)
    dt = raw.dtype
    d = raw.shape[0]
    B = raw.shape[1] if raw.ndim > 1 else 1
    K = int(np.prod(raw.shape[2:])) if raw.ndim > 2 else 1
    z = np.empty_like(raw, order="F")
    y = None if y1 is None else np.asfortranarray(y1, dtype=dt)
    rc = getattr(lib(), "hdp_oracle_mc_sample" + _sfx(dt))(C.byref(sampler_struct), _p(raw), _p(z), _p(y), d, B, K)
    if rc != 0:
        raise ValueError(f"oracle mc_sample rc={rc}")
    return z


def ml_picard(pb_struct, sampler_struct, M, L, K, x, rng: Rng):
    """solve(prob, MLP(M,L,K,mc_sample); multithreading=false) → `_ml_picard` (src/MLP.jl:102-208)."""
    x = np.ascontiguousarray(x)
    dt = x.dtype
    u = np.zeros(1, dtype=dt)
    counts = (C.c_int64 * 6)()
    rc = getattr(lib(), "hdp_oracle_ml_picard" + _sfx(dt))(C.byref(pb_struct), C.byref(sampler_struct), M, L, K,
                                                            _p(x), rng.ptr, _p(u), counts)
    info = {"path_steps": counts[0], "time_uniforms": counts[1], "sampler_calls": counts[2],
            "f_evals": counts[3], "g_evals": counts[4], "bounces": counts[5], "rc": rc}
    return u[0].item(), info


def ml_picard_mlt(pb_struct, sampler_struct, M, L, K, x, nthreads, rngs, parallel=False):
    """`_ml_picard_mlt` with Threads.nthreads() == nthreads (src/MLP.jl:218-404).
    rngs: list of nthreads+1 Rng (index 0 = main task); pass the same Rng object
    nthreads+1 times to consume one shared sequential tape (parallel must be False)."""
    x = np.ascontiguousarray(x)
    dt = x.dtype
    u = np.zeros(1, dtype=dt)
    counts = (C.c_int64 * 6)()
    arr = (C.c_void_p * (nthreads + 1))(*[r.ptr for r in rngs])
    rc = getattr(lib(), "hdp_oracle_ml_picard_mlt" + _sfx(dt))(C.byref(pb_struct), C.byref(sampler_struct), M, L, K,
                                                                _p(x), nthreads, arr, int(parallel), _p(u), counts)
    info = {"path_steps": counts[0], "time_uniforms": counts[1], "sampler_calls": counts[2],
            "f_evals": counts[3], "g_evals": counts[4], "bounces": counts[5], "rc": rc}
    return u[0].item(), info


def mlt_sde_loop(pb_struct, x, s, t, normals):
    """_mlt_sde_loop!(x2, x, s, t, prob, neumann_bc) (src/MLP.jl:406-421) for one d-vector."""
    x = np.ascontiguousarray(x)
    dt = x.dtype
    xi = np.ascontiguousarray(normals, dtype=dt)
    out = np.empty_like(x)
    real = C.c_double if dt == np.float64 else C.c_float
    fn = getattr(lib(), "hdp_oracle_mlt_sde_loop" + _sfx(dt))
    fn.argtypes = [C.c_void_p, C.c_void_p, real, real, C.c_void_p, C.c_void_p]
    rc = fn(C.byref(pb_struct), _p(x), real(s), real(t), _p(xi), _p(out))
    if rc != 0:
        raise ValueError(f"oracle mlt_sde_loop rc={rc}")
    return out


def em_ensemble(pb_struct, x0, nsteps, dt, t0, xi, saveat=False):
    """EM ensemble (src/NNKolmogorov.jl:81-112, src/NNStopping.jl:90-99; StochasticDiffEq's EM step).
    x0: d×B; xi: d×B×nsteps standard normals.  Returns (x_end d×B, traj d×(nsteps+1)×B or None)."""
    x = np.array(x0, order="F", copy=True)
    dtp = x.dtype
    d, B = x.shape
    z = np.asfortranarray(xi, dtype=dtp)
    traj = np.empty((d, nsteps + 1, B), dtype=dtp, order="F") if saveat else None
    real = C.c_double if dtp == np.float64 else C.c_float
    fn = getattr(lib(), "hdp_oracle_em_ensemble" + _sfx(dtp))
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, real, real, C.c_void_p, C.c_void_p]
    rc = fn(C.byref(pb_struct), _p(x), B, nsteps, real(dt), real(t0), _p(traj), _p(z))
    if rc != 0:
        raise ValueError(f"oracle em_ensemble rc={rc}")
    return x, traj


def sde_loop(pb_struct, y1, nsteps, dt, dWall, x0_sampler_struct=None, x0_raw=None, ts=None):
    """sde_loop!(y0, y1, dWall) (src/DeepSplitting.jl:153-166). Returns (y0, y1, iters_per_step)."""
    y1 = np.array(y1, order="F", copy=True)
    dtp = y1.dtype
    y0 = np.empty_like(y1, order="F")
    d, B = y1.shape
    dW = np.asfortranarray(dWall, dtype=dtp)
    its = np.zeros(max(nsteps, 1), dtype=np.int32)
    xr = None if x0_raw is None else np.asfortranarray(x0_raw, dtype=dtp)
    real = C.c_double if dtp == np.float64 else C.c_float
    tsa = None if ts is None else np.ascontiguousarray(ts, dtype=dtp)
    rc = getattr(lib(), "hdp_oracle_sde_loop" + _sfx(dtp))(
        C.byref(pb_struct), _p(y0), _p(y1), B, nsteps, real(dt), _p(tsa),
        C.byref(x0_sampler_struct) if x0_sampler_struct is not None else None, _p(xr), _p(dW), _p(its))
    if rc != 0:
        raise ValueError(f"oracle sde_loop rc={rc}")
    return y0, y1, its


def bench_solves(pb_struct, sampler_struct, M, L, K, x, n_solves, nthreads, seed=0):
    """n_solves independent serial `_ml_picard` solves on nthreads host threads.
    Returns (u[n_solves], seconds)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.zeros(n_solves, dtype=np.float64)
    sec = C.c_double(0)
    rc = lib().hdp_oracle_bench_solves_f64(C.byref(pb_struct), C.byref(sampler_struct), M, L, K, _p(x), n_solves,
                                           nthreads, seed, _p(u), C.byref(sec))
    if rc != 0:
        raise ValueError(f"oracle bench_solves rc={rc}")
    return u, sec.value


_user_libs = []


def set_user_kinds(mu=None, sigma=None, g=None, f=None):
    """Compile the C twins of the CUDA-C user expressions (hdp_user_exprs) with gcc and register them
    with the oracle (HDP_*_USER kinds).  Expressions use the same variable names as on the device."""
    import hashlib
    import tempfile
    src = ["#include <math.h>", "#define fmax2(a,b) ((a)>(b)?(a):(b))"]
    src.append("double u_mu(double x, int i, double t, int d, const double* p) { (void)x;(void)i;(void)t;(void)d;(void)p; return (%s); }" % (mu or "0"))
    src.append("double u_sigma(double x, int i, double t, int d, const double* p) { (void)x;(void)i;(void)t;(void)d;(void)p; return (%s); }" % (sigma or "0"))
    src.append("double u_g(double ssq, double mn, int d, const double* p) { (void)ssq;(void)mn;(void)d;(void)p; return (%s); }" % (g or "0"))
    src.append("double u_f(double vy, double vz, double ssq_y, double ssq_z, double t, int d, const double* p) { (void)vy;(void)vz;(void)ssq_y;(void)ssq_z;(void)t;(void)d;(void)p; return (%s); }" % (f or "0"))
    code = "\n".join(src)
    tag = hashlib.sha1(code.encode()).hexdigest()[:16]
    d = os.path.join(tempfile.gettempdir(), "hdp_oracle_user")
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, f"u_{tag}.so")
    if not os.path.exists(so):
        cfile = os.path.join(d, f"u_{tag}.c")
        with open(cfile, "w") as fh:
            fh.write(code)
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, cfile, "-lm"], check=True)
    ul = C.CDLL(so)
    _user_libs.append(ul)
    L = lib()
    L.hdp_oracle_set_user_fns.argtypes = [C.c_void_p] * 4
    L.hdp_oracle_set_user_fns(C.cast(ul.u_mu, C.c_void_p), C.cast(ul.u_sigma, C.c_void_p),
                              C.cast(ul.u_g, C.c_void_p), C.cast(ul.u_f, C.c_void_p))
