"""Pins the CPU oracle (oracle/hdp_oracle.c) against every closed-form answer and
predicate the reference's own tests hold for the path:
test/MLP.jl, test/reflect.jl, test/MCSample.jl.  CPU only."""
import math

import numpy as np
import pytest

from helpers import zoo


def rel_error_l2(u, uanal):  # test/MLP.jl:10-16
    return math.sqrt((u - uanal) ** 2 / u ** 2) if abs(uanal) >= 10 * np.finfo(float).eps else abs(u - uanal)


def _solve(H, O, prob, alg, seed, dtype=np.float64, fast=True):
    pb, keep = prob.to_c()
    sm, k2 = alg.mc_sample._descriptor(prob.d)
    normal = isinstance(alg.mc_sample, H.NormalSampling)
    rng = (O.Rng.xoshiro if fast else O.Rng.philox)(seed, dtype, normal)
    u, info = O.ml_picard(pb, sm, alg.M, alg.L, alg.K, np.asarray(prob.x, dtype=dtype), rng)
    assert info["rc"] == 0
    return u, info


def test_philox_known_answers(O):
    # Random123 kat_vectors: philox4x32_10
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        assert [int(v) for v in O.philox4x32_10(ctr, key)] == want
        assert [int(v) for v in O.philox4x32(ctr, key, 10)] == want
    # the library draws with HDP_PHILOX_ROUNDS rounds of the same round function and key schedule:
    # R rounds followed by 10 - R more from the bumped key are the 10-round vector
    R = O.philox_rounds()
    assert R == 7
    for ctr, key, want in kat:
        mid = O.philox4x32(ctr, key, R)
        key2 = [(key[0] + R * 0x9E3779B9) & 0xFFFFFFFF, (key[1] + R * 0xBB67AE85) & 0xFFFFFFFF]
        assert [int(v) for v in O.philox4x32(mid, key2, 10 - R)] == want


def test_draw_addressing_layout(O):
    """The library's draws: Philox4x32-7 (HDP_PHILOX_ROUNDS), fixed key, counter (w0, w1, seed_lo, seed_hi) with the documented
    packing of (stream, ordinal, group) — include/hdpde.h, DESIGN.md section 3."""
    key = [0xA4093822, 0x299F31D0]
    seed = 0x0123456789ABCDEF

    def words(stream, ord_, group):
        if stream <= 2:  # TIME, STEP, SAMPLER: ordinal < 2^40, group < 2^21
            w0 = ord_ & 0xFFFFFFFF
            w1 = ((ord_ >> 32) | (group << 8) | (stream << 29)) & 0xFFFFFFFF
        else:            # SDE, X0, MC: ordinal < 2^24, group < 2^37
            w0 = group & 0xFFFFFFFF
            w1 = ((group >> 32) | (ord_ << 5) | (stream << 29)) & 0xFFFFFFFF
        return [int(v) for v in O.philox4x32([w0, w1, seed & 0xFFFFFFFF, seed >> 32], key)]

    for stream, ord_, group in [(0, 5, 0), (1, (1 << 39) + 12345, (1 << 20) + 7), (2, 77, 3), (3, (1 << 23) + 9, (1 << 36) + 11),
                                (4, 0, 123456789), (5, 31, 2)]:
        w = words(stream, ord_, group)
        # f64: two uniforms per call, u = ((w_hi >> 12) : w_lo) * 2^-52
        u0 = O.draw(np.float64, seed, stream, ord_, 2 * group, False)
        u1 = O.draw(np.float64, seed, stream, ord_, 2 * group + 1, False)
        assert u0 == (((w[0] >> 12) << 32) | w[1]) * 2.0 ** -52 and u1 == (((w[2] >> 12) << 32) | w[3]) * 2.0 ** -52
        # f32: four uniforms per call, u = (w >> 9) * 2^-23
        for e in range(4):
            assert O.draw(np.float32, seed, stream, ord_, 4 * group + e, False) == np.float32((w[e] >> 9) * 2.0 ** -23)
        # f64 Box–Muller pair from the same words: u1 = (k1 | 1) 2^-52, angle 2 pi k2 / 2^52
        k1 = (((w[0] >> 12) << 32) | w[1]) | 1
        k2 = ((w[2] >> 12) << 32) | w[3]
        r = math.sqrt(-2.0 * math.log(k1 * 2.0 ** -52))
        th = 2.0 * math.pi * (k2 * 2.0 ** -52)
        assert abs(O.draw(np.float64, seed, stream, ord_, 2 * group, True) - r * math.cos(th)) <= 1e-14
        assert abs(O.draw(np.float64, seed, stream, ord_, 2 * group + 1, True) - r * math.sin(th)) <= 1e-14


@pytest.mark.parametrize("kind", ["xoshiro", "philox"])
def test_generators_are_standard(O, kind):
    rng = getattr(O.Rng, kind)(7)
    z = rng.fill(1, 200000)
    u = np.array([rng.fill(0, 1)[0] for _ in range(20000)])
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1) < 0.02
    assert abs(np.mean(z ** 4) - 3) < 0.1
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01


@pytest.mark.parametrize("M,L,K,nl,want", [
    (4, 4, 10, False, (4916, 26480, 2304)),
    (4, 5, 10, False, (41652, 224400, 19520)),
    (4, 4, 10, True, (928116, None, None)),
    (4, 5, 10, True, (45384692, 229363600, 22457920)),
    (8, 5, 10, False, (1184968, None, None)),
])
def test_work_counts(O, H, M, L, K, nl, want):
    c = O.mlp_counts(M, L, K, nl)
    assert c["path_steps"] == want[0]
    if want[1] is not None:
        assert c["f_evals"] == want[1] and c["g_evals"] == want[2]
    # the product's planner tables agree with the oracle's recurrences
    assert H.mlp_counts(M, L, K, nl) == c


def test_counts_match_actual_consumption(O, H):
    for name, (prob, alg, dt) in zoo(H).items():
        u, info = _solve(H, O, prob, alg, 3, dt)
        c = O.mlp_counts(alg.M, alg.L, alg.K, H._integrate(alg.mc_sample))
        for k in ("path_steps", "time_uniforms", "sampler_calls", "f_evals", "g_evals"):
            assert info[k] == c[k], (name, k)


def test_get_loop_num_quirk(O):
    # M=4, L=4, NT=8 runs 260/256, 68/64, 20/16, 4/4 samples for l=0..3
    tot = [sum(O.get_loop_num(4, 4 ** (4 - l), t, 8) for t in range(1, 9)) for l in range(4)]
    assert tot == [260, 68, 20, 4]
    tot1 = [O.get_loop_num(4, 4 ** (4 - l), 1, 1) for l in range(4)]
    assert tot1 == [257, 65, 17, 4]
    assert sum(O.get_loop_num(4, 16, t, 3) for t in range(1, 4)) == 18


# ---- closed-form pins of test/MLP.jl --------------------------------------------
@pytest.mark.parametrize("d", [1, 10, 50])
def test_heat_equation(O, H, d):  # test/MLP.jl:19-47
    prob = H.ParabolicPDEProblem(0.0, 0.1, np.full(d, 2.0), (0.0, 0.5), g=H.SumSq(), f=H.FZero())
    u, _ = _solve(H, O, prob, H.MLP(), 100)
    assert rel_error_l2(u, 4.0 * d + d * 0.1 ** 2 * 0.5) < 0.1
    # averaged over seeds the estimator is unbiased: much tighter than the reference's 10 %
    us = [_solve(H, O, prob, H.MLP(), s)[0] for s in range(32)]
    assert abs(np.mean(us) - (4.0 * d + d * 0.005)) < 4 * np.std(us) / math.sqrt(32) + 1e-12


def test_heat_equation_x0_zero(O, H):  # BASELINE config C1
    prob = H.ParabolicPDEProblem(0.0, 0.1, np.zeros(10), (0.0, 0.5), g=H.SumSq(), f=H.FZero())
    us = [_solve(H, O, prob, H.MLP(), s)[0] for s in range(64)]
    assert abs(np.mean(us) - 0.05) < 4 * np.std(us) / 8


@pytest.mark.parametrize("d", [1, 3, 10])
def test_exponential(O, H, d):  # test/MLP.jl:107-134
    prob = H.ParabolicPDEProblem(0.0, 0.0, np.zeros(d), (0.0, 0.5), g=H.SumSq(2.0), f=H.FLinear(0.1))
    u, _ = _solve(H, O, prob, H.MLP(), 100)
    assert rel_error_l2(u, 2.0 * math.exp(0.1 * 0.5)) < 0.1
    assert abs(u - 2.10254) < 0.01  # 2.1027 ± 0.0005


def test_allen_cahn_literature(O, H):  # test/MLP.jl:136-160 (tolerance 0.5 in the reference)
    prob = zoo(H)["allen_cahn_d10"][0]
    us = [_solve(H, O, prob, H.MLP(), s)[0] for s in range(8)]
    assert rel_error_l2(float(np.mean(us)), 0.30879) < 0.5
    assert abs(np.mean(us) - 0.241) < 0.02  # SURVEY.md App. C restatement value


@pytest.mark.parametrize("d", [1, 2, 5])
def test_allen_cahn_neumann_not_nan(O, H, d):  # test/MLP.jl:162-186
    prob = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                 neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
    u, info = _solve(H, O, prob, H.MLP(), 100)
    assert not math.isnan(u) and info["bounces"] > 0


@pytest.mark.parametrize("d", [1, 2, 5])
def test_heat_neumann_two_runs_agree(O, H, d):  # test/MLP.jl:80-105
    prob = H.ParabolicPDEProblem(0.0, 0.1, np.full(d, 0.3), (0.0, 0.5), g=H.SumSq(), f=H.FZero(),
                                 neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
    u1, _ = _solve(H, O, prob, H.MLP(), 1)
    u2, _ = _solve(H, O, prob, H.MLP(), 2)
    assert rel_error_l2(u1, u2) < 0.1


def test_black_scholes_default_risk(O, H):  # test/MLP.jl:188-230
    prob, alg, _ = zoo(H)["black_scholes_d20"]
    u, _ = _solve(H, O, prob, alg, 100)
    assert rel_error_l2(u, 60.781) < 0.1


@pytest.mark.parametrize("d", [1, 2])
def test_replicator_mutator(O, H, d):  # test/MLP.jl:236-286
    ss0, sig, t = 1e-2, 0.1, 0.5
    MM = sig
    SSt = MM * ((MM * math.sinh(MM * t) + ss0 * math.cosh(MM * t)) / (MM * math.cosh(MM * t) + ss0 * math.sinh(MM * t)))
    uanal = (2 * math.pi) ** (-d / 2) * SSt ** (-d / 2)
    prob = H.PIDEProblem(0.0, sig, np.zeros(d), (0.0, t), H.GaussDensity(ss0), H.FNLReplicatorMutator(1.0))
    alg = H.MLP(M=4, K=10, L=3 if d > 1 else 4, mc_sample=H.NormalSampling(1.0))
    us = [_solve(H, O, prob, alg, s)[0] for s in range(3)]
    assert rel_error_l2(float(np.mean(us)), uanal) < 0.2


def test_nonlocal_allen_cahn_neumann_two_runs(O, H):  # test/MLP.jl:288-317
    d = 2
    bc = (np.full(d, -0.5), np.full(d, 0.5))
    prob = H.PIDEProblem(0.0, 0.1, np.zeros(d), (0.0, 0.5), H.ExpSumSq(0.25), H.FNLAllenCahnDiff(), neumann_bc=bc)
    alg = H.MLP(M=4, K=10, L=3, mc_sample=H.UniformSampling(*bc))
    u1, _ = _solve(H, O, prob, alg, 1)
    u2, _ = _solve(H, O, prob, alg, 2)
    assert rel_error_l2(u1, u2) < 0.1


def test_threaded_variant_matches_serial_statistically(O, H):
    """`_ml_picard_mlt` (src/MLP.jl:218-260): f = 0 makes the `_get_loop_num` bias vanish."""
    prob = H.ParabolicPDEProblem(0.0, 0.1, np.full(10, 2.0), (0.0, 0.5), g=H.SumSq(), f=H.FZero())
    pb, keep = prob.to_c()
    sm, k2 = H.NoSampling()._descriptor(10)
    for nt, par in ((1, False), (8, False), (8, True)):
        rngs = [O.Rng.xoshiro(1000 + i) for i in range(nt + 1)]
        u, info = O.ml_picard_mlt(pb, sm, 4, 4, 10, prob.x, nt, rngs, parallel=par)
        assert info["rc"] == 0 and abs(u - 40.05) / 40.05 < 0.1
    # and with f != 0 the threaded estimator is the biased one (Q4): 260/256 samples at l=0
    c1 = H.mlp_counts(4, 4, 10, False, 8)
    assert c1["path_steps"] == 5264 and c1["time_uniforms"] == 2800


# ---- test/reflect.jl ------------------------------------------------------------
def test_reflect_vector_inside_and_no_nan(O):  # test/reflect.jl:8-30
    rng = np.random.default_rng(0)
    d = 10
    X0 = np.zeros(d)
    X1 = X0 + rng.standard_normal(d)
    X11 = O.reflect_vector(X0, X1, np.full(d, -1.0), np.full(d, 1.0))
    assert np.all((X11 < 1) & (X11 > -1))
    y1 = X0.copy()
    dW = rng.standard_normal((d, 1000)).astype(np.float32)
    for i in range(1000):
        y0 = y1.copy()
        y1 = y0 + dW[:, i]
        y1 = O.reflect_vector(y0, y1, np.full(d, -1.0), np.full(d, 1.0))
    assert not np.any(np.isnan(y1))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_reflect_array_equals_vector(O, dtype):  # test/reflect.jl:36-53
    rng = np.random.default_rng(1)
    d, B = 10, 200
    a = np.zeros((d, B), dtype=dtype, order="F")
    b = np.asfortranarray((a + 2 * rng.standard_normal((d, B))).astype(dtype))
    out, cnt, col, it = O.reflect_array(a, b, dtype(-1.0), dtype(1.0), return_info=True)
    assert np.all((out < 1) & (out > -1))
    tol = 1e-12 if dtype == np.float64 else 2e-6
    for j in range(0, B, 7):
        v, cv = O.reflect_vector(a[:, j].copy(), b[:, j].copy(), dtype(-1.0), dtype(1.0), return_bounces=True)
        assert np.max(np.abs(v - out[:, j])) <= tol
        assert np.array_equal(cv, cnt[:, j])  # same wall hits per coordinate (no exact ties in random data)
    assert it == col.max()


def test_reflect_array_chain_no_nan(O):  # test/reflect.jl:55-69
    rng = np.random.default_rng(2)
    d, B = 10, 500
    y1 = np.zeros((d, B), dtype=np.float32, order="F")
    for i in range(30):
        y0 = y1.copy(order="F")
        y1 = np.asfortranarray(y0 + rng.standard_normal((d, B)).astype(np.float32))
        y1 = O.reflect_array(y0, y1, np.float32(-1.0), np.float32(1.0))
    assert not np.any(np.isnan(y1)) and np.all(np.abs(y1) <= 1)


def test_reflect_start_outside_errors(O):  # src/reflect.jl:12, :59
    with pytest.raises(ValueError):
        O.reflect_vector(np.array([2.0, 0.0]), np.array([0.0, 0.0]), -1.0, 1.0)
    with pytest.raises(ValueError):
        O.reflect_array(np.array([[2.0], [0.0]]), np.array([[0.0], [0.0]]), -1.0, 1.0)


def test_reflect_matches_per_coordinate_mirror(O):
    """The survey's claim (App. C): in an axis-aligned box the sequential billiard equals the
    per-coordinate iterated mirror b <- b - 2(b - wall)."""
    rng = np.random.default_rng(3)
    for d, sc in ((10, 0.071), (100, 0.5), (100, 0.707), (300, 0.5)):
        a = rng.uniform(-0.5, 0.5, d)
        b = a + sc * rng.standard_normal(d)
        out, cnt = O.reflect_vector(a, b, -0.5, 0.5, return_bounces=True)
        m = b.copy()
        n = np.zeros(d, dtype=int)
        for i in range(d):
            while m[i] < -0.5 or m[i] > 0.5:
                w = -0.5 if m[i] < -0.5 else 0.5
                m[i] = m[i] - 2 * (m[i] - w)
                n[i] += 1
        assert np.array_equal(n, cnt)
        assert np.max(np.abs(m - out)) <= 1e-15


# ---- test/MCSample.jl -----------------------------------------------------------
def test_samplers_shape_type_range(O, H):  # test/MCSample.jl:25-42
    rng = np.random.default_rng(4)
    for dt in (np.float64, np.float32):
        raw = np.asfortranarray(rng.random((10, 80)).astype(dt))
        sm, _ = H.UniformSampling(0.0, 1.0)._descriptor(10)
        z = O.mc_sample(sm, raw)
        assert z.shape == raw.shape and z.dtype == raw.dtype and np.all((z >= 0) & (z < 1))
        sm, _ = H.UniformSampling(-1.0, 1.0)._descriptor(10)
        z = O.mc_sample(sm, raw)
        assert np.all((z >= -1) & (z < 1)) and np.allclose(z, (raw - 0.5) * 2.0)
        n = np.asfortranarray(rng.standard_normal((10, 80, 3)).astype(dt))
        y1 = np.asfortranarray(rng.standard_normal((10, 80)).astype(dt))
        sm, _ = H.NormalSampling(2.0, False)._descriptor(10)
        assert np.array_equal(O.mc_sample(sm, n, y1), n * dt(2.0))
        sm, _ = H.NormalSampling(2.0, True)._descriptor(10)
        assert np.array_equal(O.mc_sample(sm, n, y1), n * dt(2.0) + y1[:, :, None])


def test_sde_loop_oracle_matches_numpy(O, H):
    """sde_loop! (src/DeepSplitting.jl:153-166) without bc is a plain cumulative sum."""
    rng = np.random.default_rng(5)
    d, B, N = 6, 33, 7
    dW = rng.standard_normal((d, B, N))
    prob = H.ParabolicPDEProblem(0.5, 2.0, np.zeros(d), (0.0, 1.0), g=H.SumSq(), f=H.FZero())
    pb, _ = prob.to_c()
    y0, y1, _ = O.sde_loop(pb, np.zeros((d, B), order="F"), N, 0.1, dW)
    want = np.zeros((d, B))
    prev = want
    for i in range(N):
        prev = want
        want = (prev + 0.5 * 0.1) + (2.0 * math.sqrt(0.1)) * dW[:, :, i]
    assert np.array_equal(y1, want) and np.array_equal(y0, prev)


# ------------------------------ a second, independent restatement ------------------------------
def _seq_sumsq(x):
    s = 0.0
    for v in x:
        s += float(v) * float(v)
    return s


def _closure_cases(H):
    """name -> (kinds problem, alg, python closures mu, sigma, g, f, python sampler) — the closures are the
    reference's test functions transcribed literally (test/MLP.jl, docs/src/tutorials/mlp.md)."""
    from oracle import ml_picard_py as P
    mx = lambda u: u if u > 0 else 0.0  # noqa: E731
    d = 4
    c = {}
    c["heat"] = (H.ParabolicPDEProblem(0.0, 0.1, np.full(d, 2.0), (0.0, 0.5), g=H.SumSq(), f=H.FZero()), H.MLP(M=3, L=3, K=2),
                 lambda x, p, t: 0.0, lambda x, p, t: 0.1, lambda x: _seq_sumsq(x),
                 lambda y, z, vy, vz, dy, dz, p, t: 0.0 * vy, P.NoSampling())                              # test/MLP.jl:24-35
    bc = (np.full(d, -0.5), np.full(d, 0.5))
    c["allen_cahn_neumann"] = (H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                                     neumann_bc=bc), H.MLP(M=3, L=4, K=2),
                               lambda x, p, t: 0.0, lambda x, p, t: 1.0, lambda x: math.exp(-0.25 * _seq_sumsq(x)),
                               lambda y, z, vy, vz, dy, dz, p, t: 1.0 * (vy - vy * vy * vy), P.NoSampling())  # :162-186
    c["black_scholes"] = (H.ParabolicPDEProblem(H.LinearCoef(0.02), H.LinearCoef(0.2), np.full(d, 100.0), (0.0, 1.0),
                                                g=H.MinCoord(), f=H.FDefaultRisk()), H.MLP(M=3, L=3, K=2),
                          lambda x, p, t: 0.02 * x, lambda x, p, t: 0.2 * x, lambda x: float(np.min(x)),
                          None, P.NoSampling())                                                            # :188-230
    dr = H.FDefaultRisk()

    def f_risk(y, z, vy, vz, dy, dz, p, t):                                                                # :209-215
        Q = (vy < dr.vh) * dr.gh + (vy >= dr.vl) * dr.gl + ((vy >= dr.vh) and (vy < dr.vl)) * (
            ((dr.gh - dr.gl) / (dr.vh - dr.vl)) * (vy - dr.vh) + dr.gh)
        return -(1.0 - dr.delta) * Q * vy - dr.R * vy
    c["black_scholes"] = c["black_scholes"][:5] + (f_risk,) + c["black_scholes"][6:]
    c["fisher_nonlocal_neumann"] = (H.PIDEProblem(0.0, 0.1, np.zeros(d), (0.0, 0.5), H.ExpSumSq(1.0), H.FNLFisher(), neumann_bc=bc),
                                    H.MLP(M=2, L=4, K=2, mc_sample=H.UniformSampling(np.full(d, -0.5), np.full(d, 0.5))),
                                    lambda x, p, t: 0.0, lambda x, p, t: 0.1, lambda x: math.exp(-_seq_sumsq(x)),
                                    lambda y, z, vy, vz, dy, dz, p, t: mx(vy) * (1.0 - mx(vz)),
                                    P.UniformSampling(np.full(d, -0.5), np.full(d, 0.5)))                  # tutorials/mlp.md:45-58
    c["allen_cahn_diff_shifted_normal"] = (H.PIDEProblem(0.0, 0.1, np.zeros(d), (0.0, 0.5), H.ExpSumSq(0.25), H.FNLAllenCahnDiff()),
                                           H.MLP(M=2, L=4, K=2, mc_sample=H.NormalSampling(0.2, True)),
                                           lambda x, p, t: 0.0, lambda x, p, t: 0.1, lambda x: math.exp(-0.25 * _seq_sumsq(x)),
                                           lambda y, z, vy, vz, dy, dz, p, t: (vy - vy * vy * vy) - (vz - vz * vz * vz),
                                           P.NormalSampling(0.2, True))                                    # test/MLP.jl:288-317
    return c


@pytest.mark.parametrize("name", ["heat", "allen_cahn_neumann", "black_scholes", "fisher_nonlocal_neumann",
                                  "allen_cahn_diff_shifted_normal"])
def test_c_oracle_equals_independent_python_restatement(O, H, name):
    """oracle/ml_picard_py.py was written from src/MLP.jl, src/reflect.jl and src/MCSample.jl on its own and runs
    real closures; the C oracle (enumerated kinds) must give the same bits on the same tapes, consume exactly the
    same number of random numbers and count the same wall hits."""
    from helpers import oracle_solve, random_tapes
    from oracle import ml_picard_py as P
    prob, alg, mu, sigma, g, f, sampler = _closure_cases(H)[name]
    tapes = random_tapes(H, prob, alg, np.float64, seed=sum(map(ord, name)))
    want, info = oracle_solve(H, O, prob, alg, np.float64, tapes)
    pp = P.Problem(mu, sigma, g, f, prob.tspan, neumann_bc=prob.neumann_bc)
    tape = P.Tape(*tapes)
    got, st = P.solve(pp, alg.M, alg.L, alg.K, prob.x, sampler, tape)
    assert got == want, (got, want)                                   # bit for bit
    assert st["bounces"] == info["bounces"]
    assert tape.ir == len(tapes[0]) and tape.in_ == len(tapes[1])     # every draw consumed, in the same order
    if tapes[2] is not None:
        assert tape.is_ == len(tapes[2])


@pytest.mark.parametrize("nt", [1, 3, 8])
@pytest.mark.parametrize("name", ["allen_cahn_neumann", "fisher_nonlocal_neumann"])
def test_c_oracle_threaded_variant_equals_independent_python_restatement(O, H, name, nt):
    """`_ml_picard_mlt` / `_ml_picard_call` / `_get_loop_num` (src/MLP.jl:218-404, quirk Q4), tasks run one after the
    other on one shared tape: both restatements give the same bits and consume the same draws."""
    from helpers import random_tapes
    from oracle import ml_picard_py as P
    prob, alg, mu, sigma, g, f, sampler = _closure_cases(H)[name]
    tapes = random_tapes(H, prob, alg, np.float64, seed=nt + sum(map(ord, name)), compat=nt)
    pb, _k = prob.to_c()
    sm, _k2 = alg.mc_sample._descriptor(prob.d)
    rng = O.Rng.tape(np.float64, *tapes)
    want, info = O.ml_picard_mlt(pb, sm, alg.M, alg.L, alg.K, prob.x, nt, [rng] * (nt + 1), parallel=False)
    assert info["rc"] == 0
    tape = P.Tape(*tapes)
    got, st = P.solve_mlt(P.Problem(mu, sigma, g, f, prob.tspan, neumann_bc=prob.neumann_bc), alg.M, alg.L, alg.K, prob.x,
                          sampler, tape, nt)
    assert got == want and st["bounces"] == info["bounces"] and st["steps"] == info["path_steps"]
    assert tape.ir == len(tapes[0]) and tape.in_ == len(tapes[1])
    for M, num, tid, n in [(4, 256, 1, 8), (4, 4, 3, 8), (4, 16, 2, 3), (4, 1, 1, 1), (4, 4, 1, 1), (3, 27, 5, 4)]:
        assert P.get_loop_num(M, num, tid, n) == O.get_loop_num(M, num, tid, n)
