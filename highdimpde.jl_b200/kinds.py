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


def as_coef(v):
    if isinstance(v, (ConstCoef, LinearCoef)):
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


G_KINDS = (SumSq, InvSumSq, ExpSumSq, MinCoord, GaussDensity)
F_KINDS = (FZero, FLinear, FAllenCahn, FDefaultRisk, FFisherKPP, FNLAllenCahnDiff, FNLFisher, FNLReplicatorMutator)
