"""Committed golden fixtures (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py):
CPU — the oracle and the host-emulated product code reproduce them; GPU — the CUDA path does."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import golden_problems  # noqa: E402
from helpers import emu_solve, oracle_solve, rel_err  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
TOL = {np.float64: 1e-12, np.float32: 1e-5}


def _tapes(name):
    return G[f"mlp_{name}_r"], G[f"mlp_{name}_n"], (G[f"mlp_{name}_s"] if f"mlp_{name}_s" in G else None)


@pytest.mark.parametrize("name", list(golden_problems().keys()))
def test_oracle_and_emulation_reproduce_golden_mlp(H, O, emu, name):
    prob, alg, dt = golden_problems()[name]
    u, info = oracle_solve(H, O, prob, alg, dt, _tapes(name))
    assert u == G[f"mlp_{name}_u"][0] and info["bounces"] == int(G[f"mlp_{name}_bounces"])  # bit-equal
    for target in (0, 50):
        ue, rc, st = emu_solve(H, emu, prob, alg, dt, tapes=_tapes(name), target_units=target)
        assert rc == 0 and rel_err(float(ue[0]), float(G[f"mlp_{name}_u"][0])) <= TOL[dt]
        assert st.bounces == int(G[f"mlp_{name}_bounces"])


def test_oracle_reproduces_golden_elementwise(H, O):
    for tag in ("f64", "f32"):
        a, b, lo, hi = (G[f"reflect_{tag}_{k}"] for k in ("a", "b", "lo", "hi"))
        arr, cnt, col, it = O.reflect_array(a, b, lo, hi, return_info=True)
        assert np.array_equal(arr, G[f"reflect_{tag}_array"]) and np.array_equal(cnt, G[f"reflect_{tag}_count"])
        assert it == int(G[f"reflect_{tag}_iters"])
        vec = np.stack([O.reflect_vector(a[:, j].copy(), b[:, j].copy(), lo, hi) for j in range(a.shape[1])], axis=1)
        assert np.array_equal(vec, G[f"reflect_{tag}_vector"])
        dt = a.dtype
        prob = H.ParabolicPDEProblem(0.05, 1.0, np.zeros(5, dtype=dt), (0.0, 1.0), g=H.SumSq(), f=H.FZero(), neumann_bc=(-1.0, 1.0))
        y0, y1, its = O.sde_loop(prob.to_c()[0], np.zeros((5, 16), dtype=dt, order="F"), 6, 0.1, G[f"sde_{tag}_dW"])
        assert np.array_equal(y0, G[f"sde_{tag}_y0"]) and np.array_equal(y1, G[f"sde_{tag}_y1"])
    d = 4
    assert np.array_equal(O.mc_sample(H.UniformSampling(G["mc_a"], G["mc_b"])._descriptor(d)[0], G["mc_raw_u"]), G["mc_uniform"])
    assert np.array_equal(O.mc_sample(H.NormalSampling(0.3, True)._descriptor(d)[0], G["mc_raw_n"], G["mc_y1"]), G["mc_normal_shifted"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(golden_problems().keys()))
def test_gpu_reproduces_golden_mlp(H, handle, name):
    prob, alg, dt = golden_problems()[name]
    for target in (1, 40, 0):
        u, st = H._ml_picard_batch(prob, alg, tapes=_tapes(name), target_units=target, handle=handle, dtype=dt)
        assert rel_err(float(u[0]), float(G[f"mlp_{name}_u"][0])) <= TOL[dt]
        assert st["bounces"] == int(G[f"mlp_{name}_bounces"])


@pytest.mark.gpu
def test_gpu_reproduces_golden_elementwise(H, handle):
    for tag, tol in (("f64", 1e-12), ("f32", 4e-6)):
        a, b, lo, hi = (G[f"reflect_{tag}_{k}"] for k in ("a", "b", "lo", "hi"))
        out, cnt = H._reflect(a, b, lo, hi, return_bounces=True, handle=handle)
        assert np.max(np.abs(out - G[f"reflect_{tag}_array"])) <= tol and np.max(np.abs(out - G[f"reflect_{tag}_vector"])) <= tol
        assert np.array_equal(cnt, G[f"reflect_{tag}_count"])
        dt = a.dtype
        prob = H.ParabolicPDEProblem(0.05, 1.0, np.zeros(5, dtype=dt), (0.0, 1.0), g=H.SumSq(), f=H.FZero(), neumann_bc=(-1.0, 1.0))
        y1 = np.zeros((5, 16), dtype=dt, order="F")
        y0 = np.empty_like(y1, order="F")
        H.sde_loop(prob, y0, y1, 6, 0.1, dWall=G[f"sde_{tag}_dW"], handle=handle)
        t2 = 1e-12 if tag == "f64" else 1e-5
        assert np.max(np.abs(y1 - G[f"sde_{tag}_y1"])) <= t2 and np.max(np.abs(y0 - G[f"sde_{tag}_y0"])) <= t2
    d, B, K = 4, 9, 3
    z = np.empty((d, B, K), order="F")
    H.UniformSampling(G["mc_a"], G["mc_b"])(z, inject=G["mc_raw_u"], handle=handle)
    assert np.array_equal(z, G["mc_uniform"])
    H.NormalSampling(0.3, True)(z, G["mc_y1"], inject=G["mc_raw_n"], handle=handle)
    assert np.array_equal(z, G["mc_normal_shifted"])


# ------------------------------ golden_v2: EM ensembles, state-dependent coefficients ------------------------------
from make_golden_v2 import ensemble_cases, prob_of, sde_cases  # noqa: E402

G2 = np.load(os.path.join(HERE, "golden", "golden_v2.npz"))
DT = {"f64": np.float64, "f32": np.float32}


def test_oracle_reproduces_golden_v2(H, O):
    for tag, dt in DT.items():
        for name, (mu, sg, d, B, N, tspan) in ensemble_cases().items():
            h = (tspan[1] - tspan[0]) / N
            xe, traj = O.em_ensemble(prob_of(mu, sg, d, dt, tspan).to_c()[0], G2[f"ens_{tag}_{name}_x0"], N, dt(h),
                                     dt(tspan[0]), G2[f"ens_{tag}_{name}_xi"], saveat=True)
            assert np.array_equal(traj, G2[f"ens_{tag}_{name}_traj"]) and np.array_equal(xe, traj[:, -1, :])
        for name, (mu, sg, bc) in sde_cases().items():
            y0, y1, _ = O.sde_loop(prob_of(mu, sg, 5, dt, (0.0, 1.0), bc).to_c()[0], G2[f"sde_{tag}_{name}_start"], 6, 0.1,
                                   G2[f"sde_{tag}_{name}_dW"])
            assert np.array_equal(y0, G2[f"sde_{tag}_{name}_y0"]) and np.array_equal(y1, G2[f"sde_{tag}_{name}_y1"])


@pytest.mark.gpu
def test_gpu_reproduces_golden_v2(H, handle):
    for tag, dt in DT.items():
        for name, (mu, sg, d, B, N, tspan) in ensemble_cases().items():
            h = (tspan[1] - tspan[0]) / N
            traj = H.em_ensemble(prob_of(mu, sg, d, dt, tspan), G2[f"ens_{tag}_{name}_x0"], h, normals=G2[f"ens_{tag}_{name}_xi"],
                                 saveat=True, handle=handle)
            assert np.array_equal(traj, G2[f"ens_{tag}_{name}_traj"])  # the same IEEE operations in the same order
        for name, (mu, sg, bc) in sde_cases().items():
            y1 = np.array(G2[f"sde_{tag}_{name}_start"], order="F")
            y0 = np.empty_like(y1, order="F")
            H.sde_loop(prob_of(mu, sg, 5, dt, (0.0, 1.0), bc), y0, y1, 6, 0.1, dWall=G2[f"sde_{tag}_{name}_dW"], handle=handle)
            tol = TOL[dt] if bc is not None else 0.0
            assert np.max(np.abs(y1 - G2[f"sde_{tag}_{name}_y1"])) <= tol and np.max(np.abs(y0 - G2[f"sde_{tag}_{name}_y0"])) <= tol
