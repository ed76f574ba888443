"""Problem / solution containers — the drop-in surface of the reference
(src/HighDimPDE.jl:53-420), reduced to the fields the Monte-Carlo hot path reads.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from . import _capi as capi
from . import kinds as K
from .sampling import AbstractSampling, NoSampling


class _Problem:
    nonlocal_f = False

    def __init__(self, mu, sigma, x0, tspan, g, f, *, p=None, x0_sample=None, neumann_bc=None):
        self.mu = K.as_coef(mu)
        self.sigma = K.as_coef(sigma)
        self.x = np.asarray(x0)
        if self.x.dtype not in (np.float32, np.float64):
            self.x = self.x.astype(np.float64)
        self.x = np.ascontiguousarray(self.x)
        if self.x.ndim != 1:
            raise ValueError("x0 must be a vector")
        self.tspan = (float(tspan[0]), float(tspan[1]))
        if not isinstance(g, K.G_KINDS):
            raise TypeError("g must be one of the kinds in hdpde_b200.kinds (closures cannot run on the GPU)")
        if not isinstance(f, K.F_KINDS):
            raise TypeError("f must be one of the kinds in hdpde_b200.kinds (closures cannot run on the GPU)")
        self.g = g
        self.f = f
        self.p = p
        self.x0_sample = NoSampling() if x0_sample is None else x0_sample
        if not isinstance(self.x0_sample, AbstractSampling):
            raise TypeError("x0_sample must be a sampling struct")
        self.neumann_bc = neumann_bc
        if neumann_bc is not None:
            lo, hi = neumann_bc
            # type-consistency asserts of the reference (src/HighDimPDE.jl:115-122)
            if np.ndim(lo) != np.ndim(hi):
                raise ValueError("neumann_bc bounds must have the same shape")
            if np.ndim(lo) == 1 and (len(lo) != self.x.shape[0] or len(hi) != self.x.shape[0]):
                raise ValueError("neumann_bc vectors must have length d")

    @property
    def d(self) -> int:
        return int(self.x.shape[0])

    def user_exprs(self):
        """hdp_user_exprs for the *_USER kinds of this problem, or None."""
        ue = capi.hdp_user_exprs()
        used = False
        for field, obj in (("mu", self.mu), ("sigma", self.sigma), ("g", self.g), ("f", self.f)):
            if isinstance(obj, (K.UserCoef, K.UserG, K.UserF, K.UserFNL)):
                setattr(ue, field, obj.expr.encode())
                used = True
        return ue if used else None

    def check_user_kinds(self, dtype=np.float64) -> str:
        """Compile the user expressions (NVRTC, no GPU needed); raises HdpError with the compiler log."""
        ue = self.user_exprs()
        if ue is None:
            return ""
        lib = capi.load_library()
        pb, keep = self.to_c()
        log = C.create_string_buffer(1 << 16)
        rc = lib.hdp_user_kinds_check(C.byref(pb), C.byref(ue), int(np.dtype(dtype) == np.float32), log, len(log))
        msg = log.value.decode("utf-8", "replace")
        if rc != capi.HDP_OK:
            raise capi.HdpError(rc, msg)
        return msg

    def check_user_coefs(self, dtype=np.float64) -> str:
        """The same for the sde_loop! / EM-ensemble kernels (user μ, σ only)."""
        ue = self.user_exprs()
        if ue is None:
            return ""
        lib = capi.load_library()
        pb, keep = self.to_c()
        log = C.create_string_buffer(1 << 16)
        rc = lib.hdp_user_coefs_check(C.byref(pb), C.byref(ue), int(np.dtype(dtype) == np.float32), log, len(log))
        msg = log.value.decode("utf-8", "replace")
        if rc != capi.HDP_OK:
            raise capi.HdpError(rc, msg)
        return msg

    def to_c(self):
        """hdp_problem descriptor (+ the numpy arrays it points to, which must stay alive)."""
        pb = capi.hdp_problem()
        pb.d = self.d
        pb.mu_kind = self.mu.kind
        pb.sigma_kind = self.sigma.kind
        pb.g_kind = self.g.kind
        pb.f_kind = self.f.kind
        pb.t0, pb.t1 = self.tspan
        pb.mu_p[:] = self.mu.params()
        pb.sigma_p[:] = self.sigma.params()
        pb.g_p[:] = [float(v) for v in self.g.params()]
        pb.f_p[:] = [float(v) for v in self.f.params()]
        keep = []
        if self.neumann_bc is None:
            pb.has_bc = 0
            pb.bc_is_scalar = 1
        else:
            lo, hi = self.neumann_bc
            pb.has_bc = 1
            if np.ndim(lo) == 0:
                pb.bc_is_scalar = 1
                pb.lo_scalar, pb.hi_scalar = float(lo), float(hi)
            else:
                lo = np.ascontiguousarray(lo, dtype=np.float64)
                hi = np.ascontiguousarray(hi, dtype=np.float64)
                pb.bc_is_scalar = 0
                pb.lo = lo.ctypes.data_as(C.POINTER(C.c_double))
                pb.hi = hi.ctypes.data_as(C.POINTER(C.c_double))
                keep = [lo, hi]
        return pb, keep


class PIDEProblem(_Problem):
    """PIDEProblem(μ, σ, x0, tspan, g, f; p, x0_sample, neumann_bc) (src/HighDimPDE.jl:53-135):
    `f` is the 8-argument non-local form f(x, y, v_x, v_y, ∇v_x, ∇v_y, p, t)."""
    nonlocal_f = True

    def __init__(self, mu, sigma, x0, tspan, g, f, **kw):
        super().__init__(mu, sigma, x0, tspan, g, f, **kw)


class ParabolicPDEProblem(_Problem):
    """ParabolicPDEProblem(μ, σ, x0, tspan; g, f, p, x0_sample, neumann_bc)
    (src/HighDimPDE.jl:160-303): `f` is the 5-argument local form f(x, v_x, ∇v_x, p, t)."""

    def __init__(self, mu, sigma, x0, tspan, *, g, f, **kw):
        super().__init__(mu, sigma, x0, tspan, g, f, **kw)
        if getattr(self.f, "nonlocal_", False):
            raise TypeError("ParabolicPDEProblem takes a local f kind; use PIDEProblem for non-local f")


@dataclass
class PIDESolution:
    """PIDESolution(x0, ts, losses, usol, ufuns) (src/HighDimPDE.jl:386-410)."""
    x0: Any
    ts: Any
    losses: Any
    us: Any
    ufuns: Any
    stats: dict = field(default_factory=dict)
