"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI,
against the CPU oracle on identical inputs.  Tolerances are north_star's: bit-exact
reflection decisions (wall-hit counts) and indexing, <=1e-12 relative (fp64) / <=1e-5
(fp32) on identical injected Brownian increments and sample buffers."""
import math
import os

import numpy as np
import pytest

from helpers import oracle_solve, random_tapes, rel_err, zoo

pytestmark = pytest.mark.gpu
TOL = {np.float64: 1e-12, np.float32: 1e-5}


def _names():
    import hdpde_b200 as H
    return list(zoo(H).keys())


# ------------------------------ K4: _reflect ------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("d,B", [(10, 1), (10, 1001), (100, 257), (7, 33)])
def test_reflect_matches_oracle(H, O, handle, dtype, d, B):
    rng = np.random.default_rng(d * 1000 + B)
    lo = np.linspace(-1.0, -0.5, d).astype(dtype)
    hi = np.linspace(0.5, 1.5, d).astype(dtype)
    a = np.asfortranarray((lo[:, None] + (hi - lo)[:, None] * rng.random((d, B))).astype(dtype))
    a = np.clip(a, lo[:, None], hi[:, None])
    b = np.asfortranarray((a + 2.0 * rng.standard_normal((d, B))).astype(dtype))
    out, cnt = H._reflect(a, b, lo, hi, return_bounces=True, handle=handle)
    assert np.all((out >= lo[:, None]) & (out <= hi[:, None]))
    # Array method of the reference
    oa, ocnt, col, it = O.reflect_array(a, b, lo, hi, return_info=True)
    tol = 1e-12 if dtype == np.float64 else 4e-6
    assert np.max(np.abs(out - oa)) <= tol
    assert np.array_equal(cnt, ocnt)  # wall hits per coordinate: exact
    # Vector method of the reference on a few columns
    for j in range(0, B, max(1, B // 5)):
        ov, cv = O.reflect_vector(a[:, j].copy(), b[:, j].copy(), lo, hi, return_bounces=True)
        assert np.max(np.abs(out[:, j] - ov)) <= tol
        assert np.array_equal(cnt[:, j], cv)


def test_reflect_reference_predicates(H, handle):  # test/reflect.jl:36-69
    rng = np.random.default_rng(0)
    d, B = 10, 10000
    y1 = np.zeros((d, B), dtype=np.float32, order="F")
    for i in range(20):
        y0 = y1.copy(order="F")
        y1 = np.asfortranarray(y0 + rng.standard_normal((d, B)).astype(np.float32))
        y1 = H._reflect(y0, y1, np.float32(-1.0), np.float32(1.0), handle=handle)
    assert not np.any(np.isnan(y1)) and np.all((y1 <= 1) & (y1 >= -1))


def test_reflect_edge_cases(H, handle):
    # start outside the box: `error("a not in hypercube")` (src/reflect.jl:12)
    with pytest.raises(H.StartOutsideDomain):
        H._reflect(np.array([2.0, 0.0]), np.array([0.0, 0.0]), -1.0, 1.0, handle=handle)
    # empty batch
    out = H._reflect(np.zeros((5, 0), order="F"), np.zeros((5, 0), order="F"), -1.0, 1.0, handle=handle)
    assert out.shape == (5, 0)
    # points exactly on the walls stay; far excursions fold back inside
    a = np.zeros(4)
    b = np.array([1.0, -1.0, 1e6 + 0.25, -3.5])
    out, cnt = H._reflect(a, b, -1.0, 1.0, return_bounces=True, handle=handle)
    assert out[0] == 1.0 and out[1] == -1.0 and abs(out[3] - 0.5) < 1e-15 and np.all(np.abs(out) <= 1)
    assert cnt[0] == 0 and cnt[1] == 0 and cnt[3] == 2 and cnt[2] == 500000


# ------------------------------ K2: samplers ------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_samplers_injected_match_oracle(H, O, handle, dtype):
    rng = np.random.default_rng(3)
    d, B, K = 10, 77, 5
    raw_u = np.asfortranarray(rng.random((d, B, K)).astype(dtype))
    raw_n = np.asfortranarray(rng.standard_normal((d, B, K)).astype(dtype))
    y1 = np.asfortranarray(rng.standard_normal((d, B)).astype(dtype))
    cases = [(H.UniformSampling(-0.5, 0.5), raw_u), (H.UniformSampling(np.linspace(-2, -1, d), np.linspace(1, 3, d)), raw_u),
             (H.NormalSampling(0.3, False), raw_n), (H.NormalSampling(0.3, True), raw_n)]
    for smp, raw in cases:
        z = np.empty((d, B, K), dtype=dtype, order="F")
        smp(z, y1, inject=raw, handle=handle)
        sm, _k = smp._descriptor(d)
        zo = O.mc_sample(sm, raw, y1)
        assert np.array_equal(z, zo)  # same IEEE operations in the same order: bit-equal


def test_samplers_reference_predicates(H, handle):  # test/MCSample.jl:5-42
    X = np.zeros((10, 8000), dtype=np.float32, order="F")
    out = H.UniformSampling(-1.0, 1.0)(X, handle=handle)
    assert out.shape == X.shape and out.dtype == X.dtype and np.all((-1.0 <= out) & (out < 1.0))
    assert abs(out.mean()) < 0.01 and abs(out.var() - 1 / 3) < 0.01
    X64 = np.zeros((10, 8000), order="F")
    out = H.NormalSampling(1.0, False)(X64, handle=handle)
    assert out.shape == X64.shape and abs(out.mean()) < 0.02 and abs(out.var() - 1) < 0.02
    ref = np.full((10, 8000), 5.0, order="F")
    out = H.NormalSampling(1.0, True)(np.zeros((10, 8000), order="F"), ref, handle=handle)
    assert abs(out.mean() - 5.0) < 0.02
    assert H.NoSampling()(X) is None


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_device_generator_is_standard_normal(H, handle, dtype):
    z = handle.tape_fill(dtype, H.capi.HDP_STREAM_STEP, True, 99, 0, 4000, 250).astype(np.float64).ravel()
    assert abs(z.mean()) < 4 / math.sqrt(z.size) and abs(z.var() - 1) < 0.01
    assert abs(np.mean(z ** 3)) < 0.02 and abs(np.mean(z ** 4) - 3) < 0.05
    u = handle.tape_fill(dtype, H.capi.HDP_STREAM_TIME, False, 99, 0, 100000, 1).astype(np.float64).ravel()
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.005
    # counter-based: any sub-range reproduces the same numbers
    z2 = handle.tape_fill(dtype, H.capi.HDP_STREAM_STEP, True, 99, 1000, 10, 250)
    assert np.array_equal(z2.ravel(), z.reshape(4000, 250)[1000:1010].astype(dtype).ravel())


def test_device_philox_matches_cpu_philox(H, O, handle):
    """Same counter layout on both sides: uniforms (no transcendental) are bit-equal."""
    u = handle.tape_fill(np.float64, H.capi.HDP_STREAM_SAMPLER, False, 12345678901234, 5, 7, 9)
    uo = O.philox_tape(np.float64, 12345678901234, H.capi.HDP_STREAM_SAMPLER, 5, 7, 9, False)
    assert np.array_equal(u, uo)
    u32 = handle.tape_fill(np.float32, H.capi.HDP_STREAM_X0, False, 42, 0, 3, 10)
    assert np.array_equal(u32, O.philox_tape(np.float32, 42, H.capi.HDP_STREAM_X0, 0, 3, 10, False))
    z = handle.tape_fill(np.float64, H.capi.HDP_STREAM_STEP, True, 7, 0, 50, 10)
    zo = O.philox_tape(np.float64, 7, H.capi.HDP_STREAM_STEP, 0, 50, 10, True)
    assert np.max(np.abs(z - zo)) < 1e-13  # libm vs CUDA transcendentals


# ------------------------------ K1: sde_loop! -----------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("d,B,N,bc", [(10, 100, 10, (-1.0, 1.0)), (100, 64, 10, (-0.5, 0.5)), (7, 33, 4, None),
                                      (10, 50, 6, "vec")])
def test_sde_loop_injected_matches_oracle(H, O, handle, dtype, d, B, N, bc):
    rng = np.random.default_rng(17)
    dW = rng.standard_normal((d, B, N)).astype(dtype)
    if bc == "vec":
        bc = (np.linspace(-1.0, -0.5, d), np.linspace(0.5, 1.0, d))
    prob = H.ParabolicPDEProblem(0.05, 1.0, np.zeros(d, dtype=dtype), (0.0, 1.0), g=H.SumSq(), f=H.FZero(), neumann_bc=bc)
    y1 = np.zeros((d, B), dtype=dtype, order="F")
    y0 = np.empty_like(y1, order="F")
    H.sde_loop(prob, y0, y1, N, 0.1, dWall=dW, handle=handle)
    pb, _k = prob.to_c()
    o0, o1, its = O.sde_loop(pb, np.zeros((d, B), dtype=dtype, order="F"), N, 0.1, dW)
    tol = 1e-12 if dtype == np.float64 else 1e-5
    assert np.max(np.abs(y1 - o1)) <= tol and np.max(np.abs(y0 - o0)) <= tol
    if bc is None:
        assert np.array_equal(y1, o1) and np.array_equal(y0, o0)  # no reflection: same IEEE ops, bit-equal


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sde_loop_philox_equals_its_own_tape(H, O, handle, dtype):
    """RNG mode == tape mode with the tape generated on the fly (also pins the dWall layout)."""
    d, B, N = 100, 128, 10
    prob = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d, dtype=dtype), (0.0, 1.0), g=H.SumSq(), f=H.FZero(),
                                 neumann_bc=(-1.0, 1.0), x0_sample=H.UniformSampling(-0.25, 0.25))
    y1 = np.zeros((d, B), dtype=dtype, order="F")
    y0 = np.empty_like(y1, order="F")
    H.sde_loop(prob, y0, y1, N, 0.1, seed=5, handle=handle)
    dW = handle.tape_fill(dtype, H.capi.HDP_STREAM_SDE, True, 5, 0, N, d * B)  # [step][q]
    dW = np.asfortranarray(dW.reshape(N, B, d).transpose(2, 1, 0))
    x0raw = handle.tape_fill(dtype, H.capi.HDP_STREAM_X0, False, 5, 0, 1, d * B).reshape(B, d).T
    pb, _k = prob.to_c()
    sm, _k2 = prob.x0_sample._descriptor(d)
    o0, o1, its = O.sde_loop(pb, np.zeros((d, B), dtype=dtype, order="F"), N, 0.1, dW, sm, np.asfortranarray(x0raw))
    tol = 1e-12 if dtype == np.float64 else 1e-5
    assert np.max(np.abs(y1 - o1)) <= tol and np.max(np.abs(y0 - o0)) <= tol
    assert np.all(np.abs(y1) <= 1)


def test_sde_loop_fused_mc_epilogue(H, handle):
    d, B, N, K = 10, 64, 3, 4
    prob = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d), (0.0, 1.0), g=H.SumSq(), f=H.FZero(), neumann_bc=(-1.0, 1.0))
    y1 = np.zeros((d, B), order="F")
    y0 = np.empty_like(y1, order="F")
    z = np.empty((d, B, K), order="F")
    H.sde_loop(prob, y0, y1, N, 0.1, seed=9, mc_sample=H.NormalSampling(0.1, True), z=z, handle=handle)
    raw = handle.tape_fill(np.float64, H.capi.HDP_STREAM_MC, True, 9, 0, K, d * B).reshape(K, B, d).transpose(2, 1, 0)
    assert np.array_equal(z, raw * 0.1 + y1[:, :, None])


def test_sde_loop_start_outside(H, handle):
    prob = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(4), (0.0, 1.0), g=H.SumSq(), f=H.FZero(), neumann_bc=(-1.0, 1.0))
    y1 = np.full((4, 8), 3.0, order="F")
    with pytest.raises(H.StartOutsideDomain):
        H.sde_loop(prob, np.empty_like(y1, order="F"), y1, 2, 0.1, handle=handle)


# ------------------------------ K3: MLP ------------------------------------------
@pytest.mark.parametrize("name", _names())
def test_mlp_injected_tapes_match_oracle(H, O, handle, name):
    prob, alg, dt = zoo(H)[name]
    tapes = random_tapes(H, prob, alg, dt, seed=21)
    uo, info = oracle_solve(H, O, prob, alg, dt, tapes)
    assert info["rc"] == 0
    for target in (1, 64, 1000, 0):
        u, st = H._ml_picard_batch(prob, alg, tapes=tapes, target_units=target, handle=handle, dtype=dt)
        assert rel_err(float(u[0]), uo) <= TOL[dt], (name, target, float(u[0]), uo)
        assert st["bounces"] == info["bounces"], (name, target)  # reflection decisions exact
        assert st["path_steps"] == info["path_steps"] and st["sampler_calls"] == info["sampler_calls"]


@pytest.mark.parametrize("G", [1, 2, 4, 8, 16, 32])
def test_mlp_every_group_size(H, O, handle, G):
    """Results do not depend on how many lanes own a path."""
    for name in ("allen_cahn_neumann_d5", "nonlocal_ac_neumann_d5", "black_scholes_d20", "allen_cahn_neumann_f32_d7"):
        prob, alg, dt = zoo(H)[name]
        tapes = random_tapes(H, prob, alg, dt, seed=8)
        uo, info = oracle_solve(H, O, prob, alg, dt, tapes)
        for target in (1, 300):
            u, st = H._ml_picard_batch(prob, alg, tapes=tapes, target_units=target, handle=handle, dtype=dt, group_size=G)
            assert st["group_size"] == G
            assert rel_err(float(u[0]), uo) <= TOL[dt] and st["bounces"] == info["bounces"]


def test_mlp_philox_mode_equals_its_own_tape(H, O, handle):
    """In-register Philox normals: dump the same draws as tapes, feed the oracle."""
    for name in ("allen_cahn_neumann_d5", "nonlocal_ac_neumann_d5", "replicator_d2", "heat_neumann_f32_d3"):
        prob, alg, dt = zoo(H)[name]
        nl = H._integrate(alg.mc_sample)
        cn = H.mlp_counts(alg.M, alg.L, alg.K, nl)
        seed = 31337
        u, st = H._ml_picard_batch(prob, alg, n_solves=3, seed=seed, handle=handle, dtype=dt)
        for i in (0, 2):
            r = handle.tape_fill(dt, H.capi.HDP_STREAM_TIME, False, seed + i, 0, cn["time_uniforms"], 1)
            n = handle.tape_fill(dt, H.capi.HDP_STREAM_STEP, True, seed + i, 0, cn["path_steps"], prob.d)
            s = None
            if nl:
                s = handle.tape_fill(dt, H.capi.HDP_STREAM_SAMPLER, isinstance(alg.mc_sample, H.NormalSampling),
                                     seed + i, 0, cn["sampler_calls"], prob.d)
            uo, info = oracle_solve(H, O, prob, alg, dt, (r, n, s))
            assert rel_err(float(u[i]), uo) <= TOL[dt], (name, i)


def _philox_tapes(H, handle, prob, alg, dt, seed):
    nl = H._integrate(alg.mc_sample)
    cn = H.mlp_counts(alg.M, alg.L, alg.K, nl)
    r = handle.tape_fill(dt, H.capi.HDP_STREAM_TIME, False, seed, 0, cn["time_uniforms"], 1)
    n = handle.tape_fill(dt, H.capi.HDP_STREAM_STEP, True, seed, 0, cn["path_steps"], prob.d)
    s = None
    if nl:
        s = handle.tape_fill(dt, H.capi.HDP_STREAM_SAMPLER, isinstance(alg.mc_sample, H.NormalSampling), seed, 0,
                             cn["sampler_calls"], prob.d)
    return r, n, s


def _fast_loop_cases(H):
    """The specialised coordinate loops of the units kernel (em_loop_fast): no boundary, symmetric scalar box,
    general scalar box; odd d (a partial vec-group), more than 32 vec-groups per lane (chunked fix-up mask),
    steps far larger than the box (second and later wall hits, far excursions)."""
    c = {}
    c["sym_d5"] = (H.ParabolicPDEProblem(0.0, 1.0, np.zeros(5), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                         neumann_bc=(-0.5, 0.5)), H.MLP(M=3, L=3, K=2), np.float64)
    c["box_d6"] = (H.ParabolicPDEProblem(0.0, 1.0, np.full(6, 0.1), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                         neumann_bc=(-0.3, 0.7)), H.MLP(M=3, L=3, K=2), np.float64)
    c["box_d7_f32"] = (H.ParabolicPDEProblem(0.0, 1.0, np.full(7, 0.1, dtype=np.float32), (0.0, 0.5), g=H.ExpSumSq(0.25),
                                             f=H.FAllenCahn(1.0), neumann_bc=(-0.3, 0.7)), H.MLP(M=3, L=3, K=2), np.float32)
    c["nobc_d9"] = (H.ParabolicPDEProblem(0.05, 0.3, np.zeros(9), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0)),
                    H.MLP(M=3, L=3, K=2), np.float64)
    c["wild_sym_d4"] = (H.ParabolicPDEProblem(0.0, 40.0, np.zeros(4), (0.0, 1.0), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                              neumann_bc=(-0.5, 0.5)), H.MLP(M=3, L=3, K=2), np.float64)
    c["wild_far_d4"] = (H.ParabolicPDEProblem(0.0, 900.0, np.zeros(4), (0.0, 1.0), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                              neumann_bc=(-0.5, 0.5)), H.MLP(M=2, L=3, K=2), np.float64)
    c["nonlocal_sym_d5"] = (H.PIDEProblem(0.0, 0.6, np.zeros(5), (0.0, 0.5), H.ExpSumSq(0.25), H.FNLAllenCahnDiff(),
                                          neumann_bc=(-0.5, 0.5)),
                            H.MLP(M=3, L=3, K=3, mc_sample=H.UniformSampling(-0.5, 0.5)), np.float64)
    c["sym_d1"] = (H.ParabolicPDEProblem(0.0, 1.0, np.zeros(1), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                         neumann_bc=(-0.5, 0.5)), H.MLP(M=3, L=3, K=2), np.float64)
    c["box_d1_f32"] = (H.ParabolicPDEProblem(0.0, 1.0, np.full(1, 0.2, dtype=np.float32), (0.0, 0.5), g=H.ExpSumSq(0.25),
                                             f=H.FAllenCahn(1.0), neumann_bc=(-0.3, 0.7)), H.MLP(M=3, L=3, K=2), np.float32)
    c["long_d2101"] = (H.ParabolicPDEProblem(0.0, 1.0, np.zeros(2101), (0.0, 0.5), g=H.ExpSumSq(1e-3), f=H.FAllenCahn(1.0),
                                             neumann_bc=(-0.5, 0.5)), H.MLP(M=2, L=2, K=2), np.float64)
    return c


@pytest.mark.parametrize("name", ["sym_d5", "box_d6", "box_d7_f32", "nobc_d9", "wild_sym_d4", "wild_far_d4",
                                  "nonlocal_sym_d5", "sym_d1", "box_d1_f32", "long_d2101"])
def test_mlp_fast_loops_match_oracle_on_their_own_draws(H, O, handle, name):
    """RNG mode runs the specialised loops; the oracle is fed the very draws (hdp_tape_fill): values to tolerance,
    wall-hit counts exact, whatever the lane-group width and the tree cut."""
    prob, alg, dt = _fast_loop_cases(H)[name]
    seed = 4242
    tapes = _philox_tapes(H, handle, prob, alg, dt, seed)
    uo, info = oracle_solve(H, O, prob, alg, dt, tapes)
    assert info["rc"] == 0
    groups = (None, 32) if name == "long_d2101" else (None, 1, 2, 8, 32)
    for G in groups:
        for target in (1, 50, 0):
            u, st = H._ml_picard_batch(prob, alg, n_solves=1, seed=seed, target_units=target, handle=handle, dtype=dt,
                                       group_size=G or 0)
            assert rel_err(float(u[0]), uo) <= TOL[dt], (name, G, target, float(u[0]), uo)
            assert st["bounces"] == info["bounces"], (name, G, target)


def test_mlp_compat_mlt_partition(H, O, handle):
    """`_ml_picard_mlt` with Threads.nthreads()==NT (src/MLP.jl:218-404), quirk Q4."""
    prob, alg, dt = zoo(H)["allen_cahn_neumann_d5"]
    for nt in (1, 8):
        tapes = random_tapes(H, prob, alg, dt, seed=5, compat=nt)
        pb, keep = prob.to_c()
        sm, k2 = alg.mc_sample._descriptor(prob.d)
        rng = O.Rng.tape(dt, *tapes)
        uo, info = O.ml_picard_mlt(pb, sm, alg.M, alg.L, alg.K, prob.x, nt, [rng] * (nt + 1), parallel=False)
        for target in (1, 200):
            u, st = H._ml_picard_batch(prob, alg, tapes=tapes, compat_mlt_threads=nt, target_units=target, handle=handle)
            assert rel_err(float(u[0]), uo) <= 1e-12 and st["bounces"] == info["bounces"]


def test_mlp_errors_and_edges(H, O, handle):
    d = 3
    # Q5: a sampled point outside the box makes `_reflect` error
    prob = H.PIDEProblem(0.0, 0.1, np.zeros(d), (0.0, 0.5), H.ExpSumSq(1.0), H.FNLFisher(),
                         neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
    with pytest.raises(H.StartOutsideDomain):
        H._ml_picard_batch(prob, H.MLP(M=2, L=3, K=2, mc_sample=H.NormalSampling(5.0)), seed=1, handle=handle)
    # root point outside the box
    prob2 = H.ParabolicPDEProblem(0.0, 0.1, np.full(d, 2.0), (0.0, 0.5), g=H.SumSq(), f=H.FZero(), neumann_bc=(-0.5, 0.5))
    with pytest.raises(H.StartOutsideDomain):
        H._ml_picard_batch(prob2, H.MLP(), seed=1, handle=handle)
    # x0_sample must be NoSampling (src/MLP.jl:65-66)
    prob3 = H.ParabolicPDEProblem(0.0, 0.1, np.zeros(d), (0.0, 0.5), g=H.SumSq(), f=H.FZero(),
                                  x0_sample=H.NormalSampling(1.0))
    with pytest.raises(H.HdpError) as e:
        H.solve(prob3, H.MLP())
    assert e.value.code == H.capi.HDP_ERR_X0_SAMPLE_NOT_SUPPORTED
    # non-local f without a sampler
    prob4 = H.PIDEProblem(0.0, 0.1, np.zeros(d), (0.0, 0.5), H.ExpSumSq(1.0), H.FNLFisher())
    with pytest.raises(H.HdpError):
        H._ml_picard_batch(prob4, H.MLP(), handle=handle)
    # L = 0 returns zero(elxType) (src/MLP.jl:117-119); L = 1, M = 1, K = 1 minimal tree; d = 1
    prob5 = H.ParabolicPDEProblem(0.0, 0.1, np.full(1, 2.0), (0.0, 0.5), g=H.SumSq(), f=H.FLinear(0.1))
    u, _ = H._ml_picard_batch(prob5, H.MLP(L=0), handle=handle)
    assert u[0] == 0.0
    alg = H.MLP(M=1, L=1, K=1)
    tapes = random_tapes(H, prob5, alg, np.float64, seed=3)
    uo, info = oracle_solve(H, O, prob5, alg, np.float64, tapes)
    u, _ = H._ml_picard_batch(prob5, alg, tapes=tapes, handle=handle)
    assert rel_err(float(u[0]), uo) <= 1e-12
    # per-solve start points
    prob6, alg6, _ = zoo(H)["heat_d10"]
    x0 = np.asfortranarray(np.stack([np.full(10, 2.0), np.full(10, 1.0), np.zeros(10)], axis=1))
    u, _ = H._ml_picard_batch(prob6, alg6, n_solves=3, x0=x0, seed=2, handle=handle)
    assert abs(u[0] - 40.05) < 2 and abs(u[1] - 10.05) < 1 and abs(u[2] - 0.05) < 0.05


def test_solve_dropin_surface(H, handle):
    """solve(prob, MLP()) returns PIDESolution(x, [tspan...], nothing, [g(x), usol], nothing) (src/MLP.jl:99)."""
    H.set_default_handle(handle)
    try:
        d = 10
        prob = H.ParabolicPDEProblem(0.0, 0.1, np.full(d, 2.0), (0.0, 0.5), g=H.SumSq(), f=H.FZero())
        sol = H.solve(prob, H.MLP(), multithreading=False, seed=100)
        assert sol.ts == [0.0, 0.5] and sol.us[0] == 40.0 and sol.losses is None and sol.ufuns is None
        assert abs(sol.us[-1] - 40.05) / 40.05 < 0.1  # test/MLP.jl:45
        sol2 = H.solve(prob, H.MLP(), multithreading=True, seed=100)
        assert abs(sol2.us[-1] - 40.05) / 40.05 < 0.1
    finally:
        H.set_default_handle(None)


# ------------------------------ `_mlt_sde_loop!` stand-alone ----------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_mlt_sde_loop_single_step_matches_oracle(H, O, handle, dtype):
    """src/MLP.jl:406-421: x2 = x + (μ*dt + σ*sqrt(dt)*ξ) — the MLP association, which differs from the
    DeepSplitting step — followed by the Vector `_reflect`."""
    rng = np.random.default_rng(12)
    d, B = 10, 37
    bc = (np.linspace(-1.0, -0.5, d), np.linspace(0.5, 1.0, d))
    prob = H.ParabolicPDEProblem(H.LinearCoef(0.3), 0.8, np.zeros(d, dtype=dtype), (0.0, 1.0), g=H.SumSq(), f=H.FZero(),
                                 neumann_bc=bc)
    x = np.asfortranarray(rng.uniform(-0.5, 0.5, (d, B)).astype(dtype))
    xi = np.asfortranarray(rng.standard_normal((d, B)).astype(dtype))
    x2 = np.empty_like(x, order="F")
    H._mlt_sde_loop(x2, x, 0.1, 0.45, prob, normals=xi, handle=handle)
    pb, _k = prob.to_c()
    tol = 1e-12 if dtype == np.float64 else 2e-6
    for j in range(B):
        want = O.mlt_sde_loop(pb, x[:, j].copy(), 0.1, 0.45, xi[:, j].copy())
        assert np.max(np.abs(x2[:, j] - want)) <= tol
    # without bc: the same IEEE operations in the same order, bit-equal
    prob_nb = H.ParabolicPDEProblem(H.LinearCoef(0.3), 0.8, np.zeros(d, dtype=dtype), (0.0, 1.0), g=H.SumSq(), f=H.FZero())
    H._mlt_sde_loop(x2, x, 0.1, 0.45, prob_nb, normals=xi, handle=handle)
    pbn, _k = prob_nb.to_c()
    want = np.stack([O.mlt_sde_loop(pbn, x[:, j].copy(), 0.1, 0.45, xi[:, j].copy()) for j in range(B)], axis=1)
    assert np.array_equal(x2, want)
    # in-register normals: the same numbers hdp_tape_fill reports
    H._mlt_sde_loop(x2, x, 0.1, 0.45, prob_nb, seed=3, handle=handle)
    z = handle.tape_fill(dtype, H.capi.HDP_STREAM_SDE, True, 3, 0, 1, d * B).reshape(B, d).T
    want = np.stack([O.mlt_sde_loop(pbn, x[:, j].copy(), 0.1, 0.45, np.ascontiguousarray(z[:, j])) for j in range(B)], axis=1)
    assert np.array_equal(x2, want)
