"""Monte-Carlo samplers — mirror of src/MCSample.jl:22-94.

`UniformSampling(a, b)`, `NormalSampling(σ, shifted)` and `NoSampling()` are
callable structs that mutate their first argument, exactly like the reference's
call operators (:28-33, :63-71, :86); the draws come from the K2 kernel.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi


class AbstractSampling:
    def _descriptor(self, d: int):
        raise NotImplementedError


class MCSampling(AbstractSampling):
    pass


def _call_mc_sample(desc, keep, x_mc: np.ndarray, x, seed, handle, inject=None):
    if x_mc.dtype not in (np.float32, np.float64):
        raise TypeError("x_mc must be float32 or float64")
    if not x_mc.flags.f_contiguous and x_mc.ndim > 1:
        raise ValueError("x_mc must be column-major (order='F'), as Julia arrays are")
    h = handle or capi.default_handle()
    d = x_mc.shape[0]
    if x_mc.ndim == 1:
        B, K = 1, 1
    elif x_mc.ndim == 2:
        B, K = x_mc.shape[1], 1
    else:
        B, K = x_mc.shape[1], int(np.prod(x_mc.shape[2:]))
    fn = h.lib.hdp_mc_sample_f64 if x_mc.dtype == np.float64 else h.lib.hdp_mc_sample_f32
    y1 = None
    if x is not None:
        y1 = np.asarray(x, dtype=x_mc.dtype, order="F")
        if y1.size != d * B:
            raise ValueError("reference points must be d×B")
    inj = None if inject is None else np.asarray(inject, dtype=x_mc.dtype, order="F")
    capi.check(h.h, fn(h.h, capi.np_ptr(x_mc), capi.np_ptr(y1), d, B, K, C.byref(desc), int(seed),
                       capi.np_ptr(inj), capi.HDP_MEM_HOST))
    return x_mc


_seed_counter = [0x5EED]


def _next_seed():
    _seed_counter[0] += 1
    return _seed_counter[0]


class UniformSampling(MCSampling):
    """UniformSampling(a, b): uniform in the hypercube [a, b]^d (src/MCSample.jl:22-33).
    a, b scalars or length-d vectors."""

    def __init__(self, a, b):
        self.a = a
        self.b = b

    def _descriptor(self, d: int):
        s = capi.hdp_sampler()
        s.kind = capi.HDP_SAMPLER_UNIFORM
        keep = []
        if np.ndim(self.a) == 0 and np.ndim(self.b) == 0:
            s.bounds_are_scalar = 1
            s.a_scalar, s.b_scalar = float(self.a), float(self.b)
        else:
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(self.a, dtype=np.float64), (d,)))
            b = np.ascontiguousarray(np.broadcast_to(np.asarray(self.b, dtype=np.float64), (d,)))
            s.bounds_are_scalar = 0
            s.a = a.ctypes.data_as(C.POINTER(C.c_double))
            s.b = b.ctypes.data_as(C.POINTER(C.c_double))
            keep = [a, b]
        return s, keep

    def __call__(self, x_mc, x=None, *, seed=None, handle=None, inject=None):
        # the second positional argument is ignored (kwargs... slurp, :28)
        desc, keep = self._descriptor(x_mc.shape[0])
        return _call_mc_sample(desc, keep, x_mc, None, _next_seed() if seed is None else seed, handle, inject)


class NormalSampling(MCSampling):
    """NormalSampling(σ, shifted=false) (src/MCSample.jl:55-71)."""

    def __init__(self, sigma, shifted: bool = False):
        self.sigma = float(sigma)
        self.shifted = bool(shifted)

    def _descriptor(self, d: int):
        s = capi.hdp_sampler()
        s.kind = capi.HDP_SAMPLER_NORMAL
        s.shifted = int(self.shifted)
        s.bounds_are_scalar = 1
        s.sigma = self.sigma
        return s, []

    def __call__(self, x_mc, x=None, *, seed=None, handle=None, inject=None):
        desc, keep = self._descriptor(x_mc.shape[0])
        return _call_mc_sample(desc, keep, x_mc, x if self.shifted else None,
                               _next_seed() if seed is None else seed, handle, inject)


class NoSampling(AbstractSampling):
    """NoSampling()(x...) = nothing (src/MCSample.jl:84-86)."""

    def _descriptor(self, d: int):
        s = capi.hdp_sampler()
        s.kind = capi.HDP_SAMPLER_NONE
        s.bounds_are_scalar = 1
        return s, []

    def __call__(self, *x, **kw):
        return None


def _integrate(mc) -> bool:
    """src/MCSample.jl:88-94"""
    return not isinstance(mc, NoSampling)
