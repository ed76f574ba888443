"""hdpde_b200 — B200-native (sm_100a) Monte-Carlo hot path of SciML/HighDimPDE.jl.

Host-side mirror of the reference interface for the accelerated path
(`PIDEProblem` / `ParabolicPDEProblem`, `MLP`, `solve`, the samplers, `_reflect`,
`sde_loop!`) over the C ABI of libhdpde_b200.so (include/hdpde.h).  The directory
is named `highdimpde.jl_b200`; import it through the `hdpde_b200` shim at the
repo root.  There is no CPU fallback: without the CUDA library and a B200 the
calls raise.
"""
from . import _capi as capi
from ._capi import Handle, HdpError, StartOutsideDomain, default_handle, load_library, set_default_handle
from . import kinds
from .kinds import (ConstCoef, LinearCoef, QuadraticCoef, SumSq, InvSumSq, ExpSumSq, MinCoord, GaussDensity, FZero, FLinear,
                    FAllenCahn, FDefaultRisk, FFisherKPP, FNLAllenCahnDiff, FNLFisher, FNLReplicatorMutator,
                    UserCoef, UserG, UserF, UserFNL)
from .sampling import UniformSampling, NormalSampling, NoSampling, _integrate
from .problems import PIDEProblem, ParabolicPDEProblem, PIDESolution
from .mlp import MLP, solve, mlp_counts, _ml_picard_batch, _ml_picard_device
from .sde import _reflect, sde_loop, sde_loop_device, sde_loop_epochs, _mlt_sde_loop, em_ensemble

__all__ = [n for n in dir() if not n.startswith("__")]
