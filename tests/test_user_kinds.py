"""User-defined kinds (HDP_*_USER): CUDA-C expressions compiled with NVRTC into
the K3 kernels.  CPU: the expressions compile (NVRTC needs no GPU) and bad ones are rejected with
the compiler's message; the oracle's gcc twin of the same expressions equals the built-in kinds.
GPU: the NVRTC-built kernels reproduce the built-in kinds and the oracle on injected tapes."""
import numpy as np
import pytest

from helpers import oracle_solve, random_tapes, rel_err

BS_F = ("-(1.0-p[0])*((vy<p[2])*p[4] + (vy>=p[3])*p[5] + ((vy>=p[2])&&(vy<p[3]))*"
        "(((p[4]-p[5])/(p[2]-p[3]))*(vy-p[2])+p[4]))*vy - p[1]*vy")
BS_P = (2 / 3, 0.02, 50.0, 70.0, 0.2, 0.02)


def _bs_user(H, d=4):
    return H.ParabolicPDEProblem(H.UserCoef("p[0]*x", (0.02,)), H.UserCoef("p[0]*x", (0.2,)), np.full(d, 100.0),
                                 (0.0, 1.0), g=H.UserG("mn"), f=H.UserF(BS_F, BS_P))


def _bs_builtin(H, d=4):
    return H.ParabolicPDEProblem(H.LinearCoef(0.02), H.LinearCoef(0.2), np.full(d, 100.0), (0.0, 1.0),
                                 g=H.MinCoord(), f=H.FDefaultRisk())


def test_user_expressions_compile_without_a_gpu(H):
    assert _bs_user(H).check_user_kinds(np.float64) is not None
    assert _bs_user(H).check_user_kinds(np.float32) is not None
    nl = H.PIDEProblem(0.0, 0.1, np.zeros(3), (0.0, 0.5), H.UserG("exp(-ssq)"),
                       H.UserFNL("fmax(0.0, vy)*(1.0 - fmax(0.0, vz)) + 0.0*ssq_z"))
    nl.check_user_kinds()
    # built-in kinds need no compilation
    assert _bs_builtin(H).check_user_kinds() == ""


def test_bad_user_expressions_are_rejected_with_the_compiler_log(H):
    bad = H.ParabolicPDEProblem(0.0, 0.1, np.zeros(3), (0.0, 0.5), g=H.UserG("exp(-ssqq)"), f=H.FZero())
    with pytest.raises(H.HdpError) as e:
        bad.check_user_kinds()
    assert e.value.code == H.capi.HDP_ERR_BAD_ARG and "ssqq" in str(e.value)
    for inj in ("ssq; asm volatile(\"trap;\")", "ssq }", "ssq\n#include <x>", ""):
        p = H.ParabolicPDEProblem(0.0, 0.1, np.zeros(3), (0.0, 0.5), g=H.UserG(inj), f=H.FZero())
        with pytest.raises(H.HdpError) as e:
            p.check_user_kinds()
        assert e.value.code == H.capi.HDP_ERR_BAD_ARG


def test_user_kinds_are_callable_on_the_host(H):
    assert H.UserG("exp(-p[0]*ssq) + mn", (0.25,))(np.array([1.0, 2.0])) == pytest.approx(np.exp(-1.25) + 1.0)
    assert H.UserF("p[0]*(vy - vy*vy*vy)", (2.0,))(np.zeros(2), 0.5) == pytest.approx(2.0 * (0.5 - 0.125))


def test_oracle_user_twin_equals_builtin_kinds(H, O):
    user, ref = _bs_user(H), _bs_builtin(H)
    O.set_user_kinds("p[0]*x", "p[0]*x", "mn", BS_F)
    alg = H.MLP(M=3, L=3, K=2)
    tapes = random_tapes(H, ref, alg, np.float64, seed=9)
    u1, _ = oracle_solve(H, O, user, alg, np.float64, tapes)
    u2, _ = oracle_solve(H, O, ref, alg, np.float64, tapes)
    assert u1 == u2


@pytest.mark.gpu
def test_gpu_user_kinds_reproduce_builtin_and_oracle(H, O, handle):
    user, ref = _bs_user(H, 20), _bs_builtin(H, 20)
    alg = H.MLP(M=4, L=3, K=3)
    tapes = random_tapes(H, ref, alg, np.float64, seed=4)
    O.set_user_kinds("p[0]*x", "p[0]*x", "mn", BS_F)
    uo, info = oracle_solve(H, O, user, alg, np.float64, tapes)
    for target in (1, 60):
        uu, su = H._ml_picard_batch(user, alg, tapes=tapes, target_units=target, handle=handle)
        ub, sb = H._ml_picard_batch(ref, alg, tapes=tapes, target_units=target, handle=handle)
        assert rel_err(float(uu[0]), uo) <= 1e-12
        assert rel_err(float(uu[0]), float(ub[0])) <= 1e-12
    # RNG mode, batch
    uu, _ = H._ml_picard_batch(user, alg, n_solves=64, seed=5, handle=handle)
    ub, _ = H._ml_picard_batch(ref, alg, n_solves=64, seed=5, handle=handle)
    assert np.max(np.abs(uu - ub) / np.abs(ub)) <= 1e-12


@pytest.mark.gpu
def test_gpu_user_kinds_time_dependent_and_nonlocal(H, O, handle):
    d = 6
    bc = (np.full(d, -0.5), np.full(d, 0.5))
    # σ depends on the time the reference passes (the END of the step), g and f on sum(x^2) / u
    prob = H.PIDEProblem(H.UserCoef("p[0]*(0.5 - x)", (0.1,)), H.UserCoef("p[0]*(1.0 + t)", (0.2,)), np.zeros(d),
                         (0.0, 0.5), H.UserG("exp(-p[0]*ssq)", (0.25,)),
                         H.UserFNL("(vy - vy*vy*vy) - (vz - vz*vz*vz) + p[0]*(ssq_y - ssq_z)", (0.01,)), neumann_bc=bc)
    alg = H.MLP(M=3, L=3, K=3, mc_sample=H.UniformSampling(*bc))
    O.set_user_kinds(prob.mu.expr, prob.sigma.expr, prob.g.expr, prob.f.expr)
    tapes = random_tapes(H, prob, alg, np.float64, seed=6)
    uo, info = oracle_solve(H, O, prob, alg, np.float64, tapes)
    assert info["rc"] == 0
    for target in (1, 500):
        u, st = H._ml_picard_batch(prob, alg, tapes=tapes, target_units=target, handle=handle)
        assert rel_err(float(u[0]), uo) <= 1e-12 and st["bounces"] == info["bounces"]
    # fp32 build of the same expressions
    prob32 = H.PIDEProblem(prob.mu, prob.sigma, np.zeros(d, dtype=np.float32), (0.0, 0.5), prob.g, prob.f, neumann_bc=bc)
    t32 = random_tapes(H, prob32, alg, np.float32, seed=6)
    uo32, _ = oracle_solve(H, O, prob32, alg, np.float32, t32)
    u32, _ = H._ml_picard_batch(prob32, alg, tapes=t32, handle=handle, dtype=np.float32)
    assert rel_err(float(u32[0]), uo32) <= 1e-5


@pytest.mark.gpu
def test_gpu_user_kind_errors(H, handle):
    bad = H.ParabolicPDEProblem(0.0, 0.1, np.zeros(3), (0.0, 0.5), g=H.UserG("exp(-ssqq)"), f=H.FZero())
    with pytest.raises(H.HdpError) as e:
        H._ml_picard_batch(bad, H.MLP(L=2), handle=handle)
    assert "ssqq" in str(e.value)
    # user coefficients are an MLP feature: the DeepSplitting loop refuses them
    prob = H.ParabolicPDEProblem(H.UserCoef("x"), 1.0, np.zeros(4), (0.0, 1.0), g=H.SumSq(), f=H.FZero())
    y1 = np.zeros((4, 8), order="F")
    with pytest.raises(H.HdpError):
        H.sde_loop(prob, np.empty_like(y1, order="F"), y1, 2, 0.1, handle=handle)
