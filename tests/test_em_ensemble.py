"""state-/time-dependent coefficients on the `sde_loop!` path
(QuadraticCoef of test/DeepSplitting.jl:355, NVRTC-compiled user μ/σ with t = ts[N+1-i]) and the plain
Euler–Maruyama ensemble that NNKolmogorov / NNStopping pre-simulate with (src/NNKolmogorov.jl:81-112,
src/NNStopping.jl:90-99).

CPU part: the oracle restatement against closed forms and an independent numpy evaluation of the same
association; the NVRTC build of the K1 kernels (needs no GPU).  GPU part (-m gpu): the CUDA kernels
through the C ABI against the oracle on identical normals — bit-equal where no reflection is involved
(the same IEEE operations in the same order), <=1e-12 / <=1e-5 otherwise."""
import numpy as np
import pytest

TOL = {np.float64: 1e-12, np.float32: 1e-5}


def _prob(H, mu, sigma, d, dtype=np.float64, tspan=(0.0, 1.0), bc=None):
    return H.ParabolicPDEProblem(mu, sigma, np.zeros(d, dtype=dtype), tspan, g=H.SumSq(), f=H.FZero(), neumann_bc=bc)


# ------------------------------ oracle pins (CPU) ---------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_em_ensemble_against_numpy_restatement(H, O, dtype):
    """u = (uprev + dt*mu) + sigma*(sqrt(dt)*xi), left to right, for each enumerated kind."""
    rng = np.random.default_rng(3)
    d, B, N = 5, 7, 6
    dt, t0 = dtype(0.05), dtype(0.0)
    x0 = np.asfortranarray(rng.uniform(0.5, 1.5, (d, B)).astype(dtype))
    xi = np.asfortranarray(rng.standard_normal((d, B, N)).astype(dtype))
    kinds = {"const": (H.ConstCoef(-2.0), H.ConstCoef(2.0), lambda x: dtype(-2.0) + 0 * x, lambda x: dtype(2.0) + 0 * x),
             "linear": (H.LinearCoef(0.125), H.LinearCoef(0.5), lambda x: dtype(0.125) * x, lambda x: dtype(0.5) * x),
             "quad": (H.ConstCoef(0.0), H.QuadraticCoef(0.4), lambda x: 0 * x, lambda x: dtype(0.4) * (x * x))}
    for name, (mu, sg, fmu, fsg) in kinds.items():
        pb, _k = _prob(H, mu, sg, d, dtype).to_c()
        xe, traj = O.em_ensemble(pb, x0, N, dt, t0, xi, saveat=True)
        u = x0.copy()
        sq = np.sqrt(dt).astype(dtype)
        want_traj = [u.copy()]
        for st in range(N):
            u = ((u + dt * fmu(u)).astype(dtype) + (fsg(u) * (sq * xi[:, :, st]).astype(dtype)).astype(dtype)).astype(dtype)
            want_traj.append(u.copy())
        assert np.array_equal(xe, u), name
        assert traj.shape == (d, N + 1, B)
        assert np.array_equal(traj, np.stack(want_traj, axis=1)), name  # Array(sim): d x (N+1) x trajectories


def test_oracle_em_ensemble_closed_forms(H, O):
    """test/NNKolmogorov.jl:13-14 (mu = -2, sigma = 2): X_T ~ N(x0 - 2T, 4T) exactly under EM;
    sigma = 0: the deterministic Euler polygon x0 (1 + c dt)^N for mu = c x."""
    rng = np.random.default_rng(11)
    d, B, N, T = 1, 200000, 10, 0.5
    pb, _k = _prob(H, -2.0, 2.0, d).to_c()
    x0 = np.zeros((d, B), order="F")
    xe, _ = O.em_ensemble(pb, x0, N, T / N, 0.0, rng.standard_normal((d, B, N)))
    assert abs(xe.mean() - (-2.0 * T)) < 5 * np.sqrt(4 * T / B)
    assert abs(xe.var() - 4 * T) < 5 * 4 * T * np.sqrt(2.0 / B)
    pb2, _k = _prob(H, H.LinearCoef(0.3), 0.0, 3).to_c()
    x0 = np.asfortranarray(np.full((3, 4), 2.0))
    xe, _ = O.em_ensemble(pb2, x0, 8, 0.125, 0.0, np.zeros((3, 4, 8)))
    assert np.allclose(xe, 2.0 * (1 + 0.3 * 0.125) ** 8, rtol=1e-14)


def test_oracle_sde_loop_passes_ts_to_user_coefficients(H, O):
    """src/DeepSplitting.jl:157-160: mu(y0, p, t), sigma(y0, p, t) with t = ts[N + 1 - i]."""
    rng = np.random.default_rng(5)
    d, B, N, dt = 4, 3, 5, 0.1
    ts = np.linspace(0.5, 0.1, N)  # the reversed grid the reference walks
    O.set_user_kinds(mu="p[0]*x*t", sigma="p[0] + 0.5*t + 0.01*i")
    prob = _prob(H, H.UserCoef("p[0]*x*t", (0.3,)), H.UserCoef("p[0] + 0.5*t + 0.01*i", (0.2,)), d)
    pb, _k = prob.to_c()
    y1 = np.asfortranarray(rng.uniform(-1, 1, (d, B)))
    dW = rng.standard_normal((d, B, N))
    o0, o1, _ = O.sde_loop(pb, y1, N, dt, dW, ts=ts)
    u = y1.copy()
    sq = np.sqrt(dt)
    for st in range(N):
        prev = u
        sg = 0.2 + 0.5 * ts[st] + 0.01 * np.arange(d)[:, None]
        u = (prev + (0.3 * prev * ts[st]) * dt) + (sg * sq) * dW[:, :, st]
    assert np.array_equal(o1, u) and np.array_equal(o0, prev)


def test_user_coefficients_compile_for_k1_without_a_gpu(H):
    prob = _prob(H, H.UserCoef("p[0]*x*(1.0+t)", (0.02,)), H.UserCoef("p[0]*x*x + 0.01*i", (0.4,)), 8)
    assert prob.check_user_coefs(np.float64) is not None
    assert prob.check_user_coefs(np.float32) is not None
    assert _prob(H, 0.0, H.QuadraticCoef(0.4), 8).check_user_coefs() == ""  # enumerated kinds need no compilation
    bad = _prob(H, H.UserCoef("p[0]*y"), 1.0, 8)
    with pytest.raises(H.HdpError) as e:
        bad.check_user_coefs()
    assert e.value.code == H.capi.HDP_ERR_BAD_ARG and "\"y\"" in str(e.value)


# ------------------------------ GPU parity ----------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("d,B,N", [(10, 129, 8), (7, 33, 5), (1, 1000, 10), (100, 64, 3)])
@pytest.mark.parametrize("kind", ["const", "linear", "quad"])
def test_em_ensemble_injected_matches_oracle(H, O, handle, dtype, d, B, N, kind):
    rng = np.random.default_rng(d * 100 + B)
    mu, sg = {"const": (-2.0, 2.0), "linear": (H.LinearCoef(0.125), H.LinearCoef(0.5)),
              "quad": (0.05, H.QuadraticCoef(0.4))}[kind]
    T = 0.5
    prob = _prob(H, mu, sg, d, dtype, (0.0, T))
    x0 = np.asfortranarray(rng.uniform(0.5, 1.5, (d, B)).astype(dtype))
    xi = np.asfortranarray(rng.standard_normal((d, B, N)).astype(dtype))
    pb, _k = prob.to_c()
    dt = T / N
    want_x, want_traj = O.em_ensemble(pb, x0, N, dtype(dt), dtype(0.0), xi, saveat=True)
    got_x = H.em_ensemble(prob, x0, dt, normals=xi, handle=handle)
    got_traj = H.em_ensemble(prob, x0, dt, normals=xi, saveat=True, handle=handle)
    # no reflection, no transcendental: the same IEEE operations in the same order
    assert np.array_equal(got_x, want_x)
    assert got_traj.shape == (d, N + 1, B) and np.array_equal(got_traj, want_traj)
    assert np.array_equal(got_traj[:, 0, :], x0) and np.array_equal(got_traj[:, -1, :], got_x)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_em_ensemble_philox_equals_its_own_tape(H, O, handle, dtype):
    """In-register normals: draw(SDE, ordinal = step, i = element) — the numbers hdp_tape_fill reports."""
    d, B, N, T = 10, 256, 12, 1.0
    prob = _prob(H, H.LinearCoef(0.02), H.LinearCoef(0.2), d, dtype, (0.0, T))
    x0 = np.full(d, 100.0, dtype=dtype)  # NNStopping: every trajectory starts at prob.x
    got = H.em_ensemble(prob, x0, T / N, trajectories=B, saveat=True, seed=21, handle=handle)
    z = handle.tape_fill(dtype, H.capi.HDP_STREAM_SDE, True, 21, 0, N, d * B)  # [step][q]
    xi = np.asfortranarray(z.reshape(N, B, d).transpose(2, 1, 0))
    pb, _k = prob.to_c()
    _, want = O.em_ensemble(pb, np.asfortranarray(np.repeat(x0[:, None], B, axis=1)), N, dtype(T / N), dtype(0.0), xi,
                            saveat=True)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_em_ensemble_statistics_nnkolmogorov_shapes(H, handle):
    """test/NNKolmogorov.jl:13-19 (mu = -2, sigma = 2, T = 0.5... here dt = 0.01): X_T ~ N(x0 - 2T, 4T);
    :38-39 (geometric Brownian motion, mu = 0.125 x, sigma = 0.5 x): E X_T = x0 (1 + 0.125 dt)^N."""
    B, T, dt = 1 << 18, 0.5, 0.01
    x0 = np.asfortranarray(np.linspace(-1.0, 1.0, B)[None, :])
    xe = H.em_ensemble(_prob(H, -2.0, 2.0, 1, tspan=(0.0, T)), x0, dt, seed=1, handle=handle)
    r = (xe - x0).ravel()
    assert abs(r.mean() + 2.0 * T) < 5 * np.sqrt(4 * T / B)
    assert abs(r.var() - 4 * T) < 5 * 4 * T * np.sqrt(2.0 / B)
    N = int(round(T / dt))
    xg = H.em_ensemble(_prob(H, H.LinearCoef(0.125), H.LinearCoef(0.5), 1, tspan=(0.0, T)), np.ones((1, B), order="F"),
                       dt, seed=2, handle=handle).ravel()
    assert abs(xg.mean() - (1 + 0.125 * dt) ** N) < 5 * xg.std() / np.sqrt(B)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sde_loop_quadratic_sigma_matches_oracle(H, O, handle, dtype):
    """test/DeepSplitting.jl:352-355: mu = 0, sigma(x) = 0.4 x.^2, here also with reflection."""
    rng = np.random.default_rng(8)
    d, B, N, dt = 10, 65, 6, 0.1
    dW = rng.standard_normal((d, B, N)).astype(dtype)
    for bc in (None, (-1.0, 1.0)):
        prob = _prob(H, 0.0, H.QuadraticCoef(0.4), d, dtype, bc=bc)
        y1 = np.asfortranarray(rng.uniform(-0.9, 0.9, (d, B)).astype(dtype))
        pb, _k = prob.to_c()
        o0, o1, _ = O.sde_loop(pb, y1, N, dt, dW)
        y0 = np.empty_like(y1, order="F")
        H.sde_loop(prob, y0, y1, N, dt, dWall=dW, handle=handle)
        if bc is None:
            assert np.array_equal(y1, o1) and np.array_equal(y0, o0)
        else:
            assert np.max(np.abs(y1 - o1)) <= TOL[dtype] and np.max(np.abs(y0 - o0)) <= TOL[dtype]


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_user_coefficients_on_the_k1_paths_match_oracle(H, O, handle, dtype):
    """NVRTC-built sde_loop! / `_mlt_sde_loop!` / ensemble kernels with time- and index-dependent mu, sigma."""
    rng = np.random.default_rng(9)
    d, B, N, dt = 6, 40, 5, 0.1
    mu_e, sg_e = "p[0]*x*t", "p[0] + 0.5*t + 0.01*i"
    O.set_user_kinds(mu=mu_e, sigma=sg_e)
    prob = _prob(H, H.UserCoef(mu_e, (0.3,)), H.UserCoef(sg_e, (0.2,)), d, dtype, bc=(-2.0, 2.0))
    pb, _k = prob.to_c()
    ts = np.linspace(0.5, 0.1, N).astype(dtype)
    dW = rng.standard_normal((d, B, N)).astype(dtype)
    y1 = np.asfortranarray(rng.uniform(-1, 1, (d, B)).astype(dtype))
    o0, o1, _ = O.sde_loop(pb, y1, N, dtype(dt), dW, ts=ts)
    y0 = np.empty_like(y1, order="F")
    H.sde_loop(prob, y0, y1, N, dt, ts=ts, dWall=dW, handle=handle)
    tol = TOL[dtype]
    assert np.max(np.abs(y1 - o1)) <= tol and np.max(np.abs(y0 - o0)) <= tol
    # ts is part of the contract for user kinds
    with pytest.raises(H.HdpError):
        H.sde_loop(prob, y0, y1, N, dt, dWall=dW, handle=handle)
    # the stand-alone MLP step: coefficients at the start point and the END time (src/MLP.jl:417)
    x = np.asfortranarray(rng.uniform(-1, 1, (d, B)).astype(dtype))
    xi = np.asfortranarray(rng.standard_normal((d, B)).astype(dtype))
    x2 = np.empty_like(x, order="F")
    H._mlt_sde_loop(x2, x, 0.1, 0.45, prob, normals=xi, handle=handle)
    want = np.stack([O.mlt_sde_loop(pb, x[:, j].copy(), 0.1, 0.45, xi[:, j].copy()) for j in range(B)], axis=1)
    assert np.max(np.abs(x2 - want)) <= tol
    # the ensemble: t_k = t0 + k dt
    prob_e = _prob(H, H.UserCoef(mu_e, (0.3,)), H.UserCoef(sg_e, (0.2,)), d, dtype, (0.25, 0.75))
    pbe, _k = prob_e.to_c()
    x0 = np.asfortranarray(rng.uniform(0.5, 1.5, (d, B)).astype(dtype))
    xiN = np.asfortranarray(rng.standard_normal((d, B, N)).astype(dtype))
    want_x, want_traj = O.em_ensemble(pbe, x0, N, dtype(0.1), dtype(0.25), xiN, saveat=True)
    got = H.em_ensemble(prob_e, x0, 0.1, normals=xiN, saveat=True, handle=handle)
    assert np.max(np.abs(got - want_traj)) <= tol * max(1.0, float(np.max(np.abs(want_traj))))
    # enumerated kinds afterwards: the expressions are cleared from the handle
    plain = _prob(H, 0.05, 1.0, d, dtype, (0.0, 0.5))
    pbp, _k = plain.to_c()
    wx, _ = O.em_ensemble(pbp, x0, N, dtype(0.1), dtype(0.0), xiN)
    assert np.array_equal(H.em_ensemble(plain, x0, 0.1, normals=xiN, handle=handle), wx)


@pytest.mark.gpu
def test_em_ensemble_edges(H, handle):
    """Empty ensembles, zero steps, a vector start point without `trajectories`, wrong shapes."""
    prob = _prob(H, H.LinearCoef(0.1), 0.3, 4, tspan=(0.0, 0.5))
    assert H.em_ensemble(prob, np.zeros((4, 0), order="F"), 0.1, handle=handle).shape == (4, 0)
    x0 = np.asfortranarray(np.arange(12.0).reshape(4, 3))
    tr = H.em_ensemble(prob, x0, 1.0, saveat=True, handle=handle)  # dt > T - t0 rounds to zero steps
    assert tr.shape == (4, 1, 3) and np.array_equal(tr[:, 0, :], x0)
    with pytest.raises(ValueError):
        H.em_ensemble(prob, np.zeros(4), 0.1, handle=handle)
    with pytest.raises(ValueError):
        H.em_ensemble(prob, np.zeros((5, 2), order="F"), 0.1, handle=handle)
    with pytest.raises(ValueError):
        H.em_ensemble(prob, x0, 0.1, normals=np.zeros((4, 3, 4)), handle=handle)
    # a user kind without expressions on the handle is the caller's error, reported as such
    lib = handle.lib
    pb, _k = _prob(H, H.UserCoef("x"), 1.0, 4).to_c()
    H.capi.check(handle.h, lib.hdp_set_user_exprs(handle.h, None))
    x = np.zeros((4, 2), order="F")
    import ctypes as C
    rc = lib.hdp_em_ensemble_f64(handle.h, C.byref(pb), H.capi.np_ptr(x), 2, 3, C.c_double(0.1), C.c_double(0.0), None, 1,
                                 None, H.capi.HDP_MEM_HOST)
    assert rc == H.capi.HDP_ERR_BAD_ARG


@pytest.mark.gpu
def test_mlp_with_quadratic_sigma_matches_oracle(H, O, handle):
    """QuadraticCoef inside the MLP recursion (served as the user expression c*(x*x) through NVRTC) against the
    oracle's native kind, on injected tapes."""
    from helpers import oracle_solve, random_tapes, rel_err
    prob = H.ParabolicPDEProblem(H.LinearCoef(0.02), H.QuadraticCoef(0.4), np.full(6, 0.7), (0.0, 0.5), g=H.SumSq(),
                                 f=H.FAllenCahn(1.0), neumann_bc=(-2.0, 2.0))
    alg = H.MLP(M=3, L=3, K=2)
    tapes = random_tapes(H, prob, alg, np.float64, seed=4)
    want, _ = oracle_solve(H, O, prob, alg, np.float64, tapes)
    u, st = H._ml_picard_batch(prob, alg, n_solves=1, seed=0, handle=handle, tapes=tapes)
    assert rel_err(float(u[0]), float(want)) <= 1e-12
