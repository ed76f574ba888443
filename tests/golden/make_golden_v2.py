"""Generates tests/golden/golden_v2.npz: fixed inputs and the CPU oracle's outputs for the rows added after
golden_v1 — the plain Euler–Maruyama ensemble (src/NNKolmogorov.jl:81-112, src/NNStopping.jl:90-99) and
`sde_loop!` with the state-dependent coefficient kinds (test/DeepSplitting.jl:355, :470-471).  Like golden_v1 it
pins the ORACLE (the reference is pure Julia and cannot run here) and gives the CUDA path known answers.
Re-run:  python tests/golden/make_golden_v2.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import hdpde_b200 as H  # noqa: E402  (descriptors only; no GPU needed)
from oracle import oracle as O  # noqa: E402


def ensemble_cases():
    """name -> (mu, sigma, d, B, N, tspan)"""
    return {"const": (H.ConstCoef(-2.0), H.ConstCoef(2.0), 3, 10, 5, (0.0, 0.5)),      # test/NNKolmogorov.jl:13-14
            "gbm": (H.LinearCoef(0.125), H.LinearCoef(0.5), 4, 9, 6, (0.0, 0.75)),      # :38-39
            "stopping": (H.LinearCoef(0.05 - 0.1), H.LinearCoef(0.2), 3, 8, 9, (0.0, 3.0)),  # test/NNStopping.jl:16-17
            "quad": (H.ConstCoef(0.0), H.QuadraticCoef(0.4), 5, 7, 4, (0.0, 0.4))}      # test/DeepSplitting.jl:355


def prob_of(mu, sg, d, dt, tspan, bc=None):
    return H.ParabolicPDEProblem(mu, sg, np.zeros(d, dtype=dt), tspan, g=H.SumSq(), f=H.FZero(), neumann_bc=bc)


def sde_cases():
    """name -> (mu, sigma, bc)"""
    return {"quad_box": (H.ConstCoef(0.0), H.QuadraticCoef(0.4), (-1.0, 1.0)),
            "black_scholes": (H.LinearCoef(0.02), H.LinearCoef(0.2), None)}


def main():
    rng = np.random.default_rng(20261020)
    out = {}
    for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
        for name, (mu, sg, d, B, N, tspan) in ensemble_cases().items():
            x0 = np.asfortranarray(rng.uniform(0.5, 1.5, (d, B)).astype(dt))
            xi = np.asfortranarray(rng.standard_normal((d, B, N)).astype(dt))
            h = (tspan[1] - tspan[0]) / N
            xe, traj = O.em_ensemble(prob_of(mu, sg, d, dt, tspan).to_c()[0], x0, N, dt(h), dt(tspan[0]), xi, saveat=True)
            out.update({f"ens_{tag}_{name}_x0": x0, f"ens_{tag}_{name}_xi": xi, f"ens_{tag}_{name}_traj": traj})
        for name, (mu, sg, bc) in sde_cases().items():
            d, B, N = 5, 12, 6
            y1 = np.asfortranarray(rng.uniform(-0.9, 0.9, (d, B)).astype(dt))
            dW = rng.standard_normal((d, B, N)).astype(dt)
            y0, y1o, _ = O.sde_loop(prob_of(mu, sg, d, dt, (0.0, 1.0), bc).to_c()[0], y1, N, 0.1, dW)
            out.update({f"sde_{tag}_{name}_start": y1, f"sde_{tag}_{name}_dW": dW, f"sde_{tag}_{name}_y0": y0,
                        f"sde_{tag}_{name}_y1": y1o})
    np.savez_compressed(os.path.join(HERE, "golden_v2.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_v2.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
