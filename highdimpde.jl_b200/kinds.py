"""Function *kinds*: callable descriptors standing in for the Julia closures
μ(x,p,t), σ(x,p,t), g(x), f(...) of `PIDEProblem` / `ParabolicPDEProblem`
(src/HighDimPDE.jl:56-59).  A closure cannot cross a C ABI onto the GPU, so each
function used by the reference's tests and tutorials is an enumerated kind plus a
parameter block (include/hdpde.h).  Every kind is also an ordinary callable
(numpy), mirroring the reference signature, so `g(x)` in
`PIDESolution(x, ts, nothing, [g(x), usol], nothing)` (src/MLP.jl:99) still works.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import _capi as capi


# ------------------------------ μ, σ ------------------------------------
@dataclass(frozen=True)
class ConstCoef:
    """μ(x,p,t) = c  /  σ(x,p,t) = c   (test/MLP.jl:24-25)"""
    c: float
    kind = capi.HDP_COEF_CONST

    def __call__(self, x, p=None, t=None):
        return self.c

    def params(self):
        return [self.c, 0.0]


@dataclass(frozen=True)
class LinearCoef:
    """μ(x,p,t) = c*x  /  σ(x,p,t) = c*x   (test/MLP.jl:191-192)"""
    c: float
    kind = capi.HDP_COEF_LINEAR

    def __call__(self, x, p=None, t=None):
        return self.c * np.asarray(x)

    def params(self):
        return [self.c, 0.0]


@dataclass(frozen=True)
class QuadraticCoef:
    """σ(x,p,t) = c * x .^ 2   (test/DeepSplitting.jl:355)"""
    c: float
    kind = capi.HDP_COEF_QUADRATIC

    def __call__(self, x, p=None, t=None):
        return self.c * np.asarray(x) ** 2

    def params(self):
        return [self.c, 0.0]


_PY_ENV = {"exp": math.exp, "log": math.log, "sqrt": math.sqrt, "fabs": abs, "abs": abs, "pow": pow, "sin": math.sin,
           "cos": math.cos, "tanh": math.tanh, "fmax": max, "fmin": min, "max": max, "min": min, "floor": math.floor,
           "erf": math.erf, "M_PI": math.pi}


def _py_eval(expr, **names):
    """Best-effort host evaluation of a CUDA-C expression (same syntax as Python for the arithmetic subset);
    `a ? b : c` is not supported here — the GPU is the implementation, this only serves `g(x)` in PIDESolution."""
    return eval(expr.replace("&&", " and ").replace("||", " or "), {"__builtins__": {}}, {**_PY_ENV, **names})


@dataclass(frozen=True)
class UserCoef:
    """μ(x,p,t)[i] / σ(x,p,t)[i] as a CUDA-C expression over `x` (the coordinate), `i`, `t` (the time the
    reference passes: the end of the step in MLP, ts[N+1-i] in sde_loop!, t_k in an EM ensemble), `d`,
    `p[0..1]` — compiled with NVRTC into the MLP / sde_loop! / ensemble kernels."""
    expr: str
    p: tuple = ()
    kind = capi.HDP_COEF_USER

    def __call__(self, x, p=None, t=0.0):
        x = np.asarray(x, dtype=float)
        return np.array([_py_eval(self.expr, x=float(v), i=i, t=t, d=x.shape[0], p=list(self.p) + [0, 0])
                         for i, v in enumerate(x)])

    def params(self):
        return (list(self.p) + [0.0, 0.0])[:2]


def as_coef(v):
    if isinstance(v, (ConstCoef, LinearCoef, QuadraticCoef, UserCoef)):
        return v
    if isinstance(v, (int, float, np.floating)):
        return ConstCoef(float(v))
    raise TypeError(
        "mu/sigma must be a ConstCoef/LinearCoef kind (or a number): Python/Julia closures cannot run on the GPU")


# ------------------------------ g ---------------------------------------
@dataclass(frozen=True)
class SumSq:
    """g(x) = sum(x.^2) + c   (test/MLP.jl:28, :118)"""
    c: float = 0.0
    kind = capi.HDP_G_SUMSQ

    def __call__(self, x):
        return float(np.sum(np.asarray(x) ** 2) + self.c)

    def params(self):
        return [self.c, 0, 0, 0]


@dataclass(frozen=True)
class InvSumSq:
    """g(x) = 1/(a + b*sum(x.^2))   (test/MLP.jl:146)"""
    a: float = 2.0
    b: float = 0.4
    kind = capi.HDP_G_INV_SUMSQ

    def __call__(self, x):
        return float(1.0 / (self.a + self.b * np.sum(np.asarray(x) ** 2)))

    def params(self):
        return [self.a, self.b, 0, 0]


@dataclass(frozen=True)
class ExpSumSq:
    """g(x) = exp(-c*sum(x.^2))   (test/MLP.jl:174; docs/src/tutorials/mlp.md:17)"""
    c: float = 1.0
    kind = capi.HDP_G_EXP_SUMSQ

    def __call__(self, x):
        return float(np.exp(-self.c * np.sum(np.asarray(x) ** 2)))

    def params(self):
        return [self.c, 0, 0, 0]


@dataclass(frozen=True)
class MinCoord:
    """g(x) = minimum(x)   (test/MLP.jl:199)"""
    kind = capi.HDP_G_MIN

    def __call__(self, x):
        return float(np.min(np.asarray(x)))

    def params(self):
        return [0, 0, 0, 0]


@dataclass(frozen=True)
class GaussDensity:
    """g(x) = (2π)^(-d/2) ss0^(-d/2) exp(-0.5*sum(x.^2)/ss0)   (test/MLP.jl:265)"""
    ss0: float
    kind = capi.HDP_G_GAUSS

    def __call__(self, x):
        x = np.asarray(x)
        d = x.shape[0]
        return float((2 * math.pi) ** (-d / 2) * self.ss0 ** (-d / 2) * np.exp(-0.5 * np.sum(x ** 2) / self.ss0))

    def params(self):
        return [self.ss0, 0, 0, 0]


@dataclass(frozen=True)
class UserG:
    """g(x) as a CUDA-C expression over `ssq` (= sum(x.^2)), `mn` (= minimum(x)), `d`, `p[0..3]`."""
    expr: str
    p: tuple = ()
    kind = capi.HDP_G_USER

    def __call__(self, x):
        x = np.asarray(x, dtype=float)
        return float(_py_eval(self.expr, ssq=float(np.sum(x ** 2)), mn=float(np.min(x)), d=x.shape[0],
                              p=list(self.p) + [0] * 4))

    def params(self):
        return (list(self.p) + [0.0] * 4)[:4]


# ------------------------------ f ---------------------------------------
class _F:
    nonlocal_ = False


@dataclass(frozen=True)
class FZero(_F):
    """f(y, v_y, ∇v_y, p, t) = 0 .* v_y   (test/MLP.jl:35)"""
    kind = capi.HDP_F_ZERO

    def __call__(self, y, v_y, *a):
        return 0.0 * v_y

    def params(self):
        return [0.0] * 8


@dataclass(frozen=True)
class FLinear(_F):
    """f = r * v_y   (test/MLP.jl:122)"""
    r: float
    kind = capi.HDP_F_LINEAR

    def __call__(self, y, v_y, *a):
        return self.r * v_y

    def params(self):
        return [self.r] + [0.0] * 7


@dataclass(frozen=True)
class FAllenCahn(_F):
    """f = c * (v_y - v_y^3)   (test/MLP.jl:148 with c=-1, :176 with c=+1)"""
    c: float = 1.0
    kind = capi.HDP_F_ALLEN_CAHN

    def __call__(self, y, v_y, *a):
        return self.c * (v_y - v_y ** 3)

    def params(self):
        return [self.c] + [0.0] * 7


@dataclass(frozen=True)
class FDefaultRisk(_F):
    """Black–Scholes with default risk, f = -(1-δ) Q(v) v - R v   (test/MLP.jl:201-215)"""
    delta: float = 2.0 / 3.0
    R: float = 0.02
    vh: float = 50.0
    vl: float = 70.0
    gh: float = 0.2
    gl: float = 0.02
    kind = capi.HDP_F_DEFAULT_RISK

    def __call__(self, y, v_y, *a):
        u = v_y
        Q = (u < self.vh) * self.gh + (u >= self.vl) * self.gl + ((u >= self.vh) * (u < self.vl)) * (
            ((self.gh - self.gl) / (self.vh - self.vl)) * (u - self.vh) + self.gh)
        return -(1.0 - self.delta) * Q * v_y - self.R * v_y

    def params(self):
        return [self.delta, self.R, self.vh, self.vl, self.gh, self.gl, 0.0, 0.0]


@dataclass(frozen=True)
class FFisherKPP(_F):
    """f = max(0,v) (1 - max(0,v))   (docs/src/tutorials/mlp.md:20)"""
    kind = capi.HDP_F_FISHER_KPP

    def __call__(self, y, v_y, *a):
        m = max(0.0, v_y)
        return m * (1 - m)

    def params(self):
        return [0.0] * 8


class _FNL(_F):
    nonlocal_ = True


@dataclass(frozen=True)
class FNLAllenCahnDiff(_FNL):
    """f(y,z,v_y,v_z,…) = a(v_y) - a(v_z), a(u) = u - u^3   (test/MLP.jl:304)"""
    kind = capi.HDP_F_NL_ALLEN_CAHN_DIFF

    def __call__(self, y, z, v_y, v_z, *a):
        return (v_y - v_y ** 3) - (v_z - v_z ** 3)

    def params(self):
        return [0.0] * 8


@dataclass(frozen=True)
class FNLFisher(_FNL):
    """f = max(0,v_x) (1 - max(0,v_y))   (docs/src/tutorials/mlp.md:52)"""
    kind = capi.HDP_F_NL_FISHER

    def __call__(self, y, z, v_y, v_z, *a):
        return max(0.0, v_y) * (1 - max(0.0, v_z))

    def params(self):
        return [0.0] * 8


@dataclass(frozen=True)
class FNLReplicatorMutator(_FNL):
    """f = max(0,v_y) (m(y) - max(0,v_z) m(z) (2π)^(d/2) σs^d exp(0.5 sum(z.^2)/σs²)),
    m(x) = -0.5 sum(x.^2)   (test/MLP.jl:266-274)"""
    sigma_sampling: float = 1.0
    kind = capi.HDP_F_NL_REPLICATOR

    def __call__(self, y, z, v_y, v_z, *a):
        y = np.asarray(y)
        z = np.asarray(z)
        d = y.shape[0]
        m = lambda x: -0.5 * np.sum(x ** 2)
        ss = self.sigma_sampling
        return max(0.0, v_y) * (m(y) - max(0.0, v_z) * m(z) * (2 * math.pi) ** (d / 2) * ss ** d *
                                np.exp(0.5 * np.sum(z ** 2) / ss ** 2))

    def params(self):
        return [self.sigma_sampling] + [0.0] * 7


@dataclass(frozen=True)
class UserF(_F):
    """Local f(y, v_y, ∇v_y, p, t) as a CUDA-C expression over `vy`, `ssq_y` (= sum(y.^2)), `t`, `d`, `p[0..7]`."""
    expr: str
    p: tuple = ()
    kind = capi.HDP_F_USER

    def __call__(self, y, v_y, *a):
        y = np.asarray(y, dtype=float)
        return _py_eval(self.expr, vy=v_y, vz=0.0, ssq_y=float(np.sum(y ** 2)), ssq_z=0.0, t=0.0, d=y.shape[0],
                        p=list(self.p) + [0] * 8)

    def params(self):
        return (list(self.p) + [0.0] * 8)[:8]


@dataclass(frozen=True)
class UserFNL(_FNL):
    """Non-local f(y, z, v_y, v_z, …) as a CUDA-C expression over `vy`, `vz`, `ssq_y`, `ssq_z`, `t`, `d`, `p[0..7]`."""
    expr: str
    p: tuple = ()
    kind = capi.HDP_F_NL_USER

    def __call__(self, y, z, v_y, v_z, *a):
        y, z = np.asarray(y, dtype=float), np.asarray(z, dtype=float)
        return _py_eval(self.expr, vy=v_y, vz=v_z, ssq_y=float(np.sum(y ** 2)), ssq_z=float(np.sum(z ** 2)), t=0.0,
                        d=y.shape[0], p=list(self.p) + [0] * 8)

    def params(self):
        return (list(self.p) + [0.0] * 8)[:8]


G_KINDS = (SumSq, InvSumSq, ExpSumSq, MinCoord, GaussDensity, UserG)
F_KINDS = (FZero, FLinear, FAllenCahn, FDefaultRisk, FFisherKPP, FNLAllenCahnDiff, FNLFisher, FNLReplicatorMutator,
           UserF, UserFNL)
