"""Generates tests/golden/*.npz: inputs (including the random tapes) and the CPU oracle's
outputs for small instances of every function on the hot path.  The reference itself
cannot run here (pure Julia, no toolchain), so these vectors pin the ORACLE — any later
edit of oracle/ that changes a bit fails tests/test_golden.py — and give the CUDA path
fixed inputs with known answers.  Re-run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import hdpde_b200 as H  # noqa: E402  (descriptors only; no GPU needed)
from oracle import oracle as O  # noqa: E402
from helpers import oracle_solve, random_tapes  # noqa: E402


def golden_problems():
    """name -> (prob, alg, dtype); small trees (M=3, L<=3) so the tapes stay a few KB."""
    z = {}
    z["heat_d5"] = (H.ParabolicPDEProblem(0.0, 0.1, np.full(5, 2.0), (0.0, 0.5), g=H.SumSq(), f=H.FZero()),
                    H.MLP(M=3, L=3, K=2), np.float64)
    z["allen_cahn_neumann_d6"] = (H.ParabolicPDEProblem(0.0, 1.0, np.zeros(6), (0.0, 0.5), g=H.ExpSumSq(0.25),
                                                        f=H.FAllenCahn(1.0), neumann_bc=(np.full(6, -0.5), np.full(6, 0.5))),
                                  H.MLP(M=3, L=3, K=2), np.float64)
    z["black_scholes_d4"] = (H.ParabolicPDEProblem(H.LinearCoef(0.02), H.LinearCoef(0.2), np.full(4, 100.0), (0.0, 1.0),
                                                   g=H.MinCoord(), f=H.FDefaultRisk()), H.MLP(M=3, L=3, K=2), np.float64)
    z["nonlocal_uniform_neumann_d3"] = (H.PIDEProblem(0.0, 0.1, np.zeros(3), (0.0, 0.5), H.ExpSumSq(0.25),
                                                      H.FNLAllenCahnDiff(), neumann_bc=(np.full(3, -0.5), np.full(3, 0.5))),
                                        H.MLP(M=2, L=3, K=2, mc_sample=H.UniformSampling(np.full(3, -0.5), np.full(3, 0.5))),
                                        np.float64)
    z["replicator_normal_d2"] = (H.PIDEProblem(0.0, 0.1, np.zeros(2), (0.0, 0.5), H.GaussDensity(1e-2),
                                               H.FNLReplicatorMutator(1.0)),
                                 H.MLP(M=2, L=3, K=2, mc_sample=H.NormalSampling(1.0)), np.float64)
    z["fisher_shifted_normal_d3"] = (H.PIDEProblem(0.0, 0.1, np.zeros(3), (0.0, 0.5), H.ExpSumSq(1.0), H.FNLFisher()),
                                     H.MLP(M=2, L=3, K=2, mc_sample=H.NormalSampling(0.2, True)), np.float64)
    z["heat_neumann_f32_d3"] = (H.ParabolicPDEProblem(0.0, 0.1, np.full(3, 0.3, dtype=np.float32), (0.0, 0.5), g=H.SumSq(),
                                                      f=H.FZero(), neumann_bc=(np.full(3, -0.5), np.full(3, 0.5))),
                                H.MLP(M=3, L=3, K=2), np.float32)
    return z


def main():
    rng = np.random.default_rng(20261019)
    out = {}
    # ---- _reflect ----
    for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
        d, B = 6, 40
        lo = np.linspace(-1.0, -0.5, d).astype(dt)
        hi = np.linspace(0.5, 1.0, d).astype(dt)
        a = np.asfortranarray(np.clip((lo[:, None] + (hi - lo)[:, None] * rng.random((d, B))).astype(dt), lo[:, None], hi[:, None]))
        b = np.asfortranarray((a + 1.5 * rng.standard_normal((d, B))).astype(dt))
        arr, cnt, col, it = O.reflect_array(a, b, lo, hi, return_info=True)
        vec = np.stack([O.reflect_vector(a[:, j].copy(), b[:, j].copy(), lo, hi) for j in range(B)], axis=1)
        out.update({f"reflect_{tag}_a": a, f"reflect_{tag}_b": b, f"reflect_{tag}_lo": lo, f"reflect_{tag}_hi": hi,
                    f"reflect_{tag}_array": arr, f"reflect_{tag}_vector": np.asfortranarray(vec),
                    f"reflect_{tag}_count": cnt, f"reflect_{tag}_iters": np.int32(it)})
    # ---- samplers ----
    d, B, K = 4, 9, 3
    ru = np.asfortranarray(rng.random((d, B, K)))
    rn = np.asfortranarray(rng.standard_normal((d, B, K)))
    y1 = np.asfortranarray(rng.standard_normal((d, B)))
    sa, sb = np.linspace(-2, -1, d), np.linspace(1, 3, d)
    out.update({"mc_raw_u": ru, "mc_raw_n": rn, "mc_y1": y1, "mc_a": sa, "mc_b": sb,
                "mc_uniform": O.mc_sample(H.UniformSampling(sa, sb)._descriptor(d)[0], ru),
                "mc_normal": O.mc_sample(H.NormalSampling(0.3, False)._descriptor(d)[0], rn, y1),
                "mc_normal_shifted": O.mc_sample(H.NormalSampling(0.3, True)._descriptor(d)[0], rn, y1)})
    # ---- sde_loop! ----
    for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
        d, B, N = 5, 16, 6
        dW = rng.standard_normal((d, B, N)).astype(dt)
        prob = H.ParabolicPDEProblem(0.05, 1.0, np.zeros(d, dtype=dt), (0.0, 1.0), g=H.SumSq(), f=H.FZero(), neumann_bc=(-1.0, 1.0))
        y0, y1o, its = O.sde_loop(prob.to_c()[0], np.zeros((d, B), dtype=dt, order="F"), N, 0.1, dW)
        out.update({f"sde_{tag}_dW": dW, f"sde_{tag}_y0": y0, f"sde_{tag}_y1": y1o, f"sde_{tag}_iters": its})
    # ---- MLP ----
    for name, (prob, alg, dt) in golden_problems().items():
        r, n, s = random_tapes(H, prob, alg, dt, seed=abs(hash(name)) % 1000 + 1)
        # make the tapes independent of Python's hash randomisation
        r, n, s = random_tapes(H, prob, alg, dt, seed=sum(map(ord, name)))
        u, info = oracle_solve(H, O, prob, alg, dt, (r, n, s))
        assert info["rc"] == 0
        out.update({f"mlp_{name}_r": r, f"mlp_{name}_n": n, f"mlp_{name}_u": np.array([u], dtype=dt),
                    f"mlp_{name}_bounces": np.int64(info["bounces"])})
        if s is not None:
            out[f"mlp_{name}_s"] = s
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
