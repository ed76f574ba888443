"""GPU parity on the EXACT BASELINE configurations (run with -m gpu on a B200): the kernel instantiations
bench.py times, compared directly with the CPU oracle — not through smaller stand-ins.

C1  MLP heat d=10, L=4, x0=0 and x0=2                     injected tapes
C2  MLP Allen–Cahn d=100, Neumann ±1/2, L=4, fp64         512 solves in the throughput regime (one lane per
                                                          point: mlp_units_kernel<double,1,3,EM_SYM>), the
                                                          oracle fed the very Philox draws of each solve
C3  reflected sde_loop!, B=8192, d=100, N=10, fp32        injected dWall, host buffers and device pointers
C5  MLP Allen–Cahn d=1000, L=5, fp64                      64 solves (8 lanes per point), one checked solve
Tolerances are north_star's: <=1e-12 relative (fp64), <=1e-5 (fp32), wall-hit counts exact."""
import concurrent.futures as cf

import numpy as np
import pytest

from helpers import oracle_solve, random_tapes, rel_err

pytestmark = pytest.mark.gpu


def _c2(H, d=100):
    return (H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                  neumann_bc=(np.full(d, -0.5), np.full(d, 0.5))), H.MLP(M=4, L=4, K=10))


def _tapes(H, handle, prob, alg, seed):
    cn = H.mlp_counts(alg.M, alg.L, alg.K, False)
    r = handle.tape_fill(np.float64, H.capi.HDP_STREAM_TIME, False, seed, 0, cn["time_uniforms"], 1)
    n = handle.tape_fill(np.float64, H.capi.HDP_STREAM_STEP, True, seed, 0, cn["path_steps"], prob.d)
    return r, n, None


@pytest.mark.parametrize("x0", [0.0, 2.0])
def test_c1_heat_d10_on_tapes(H, O, handle, x0):
    d = 10  # test/MLP.jl:19-47 (x0 = 2) and BASELINE configs[0] (x0 = 0)
    prob = H.ParabolicPDEProblem(0.0, 0.1, np.full(d, x0), (0.0, 0.5), g=H.SumSq(), f=H.FZero())
    alg = H.MLP()
    tapes = random_tapes(H, prob, alg, np.float64, seed=100)
    uo, info = oracle_solve(H, O, prob, alg, np.float64, tapes)
    assert info["rc"] == 0 and abs(uo - (d * x0 * x0 + d * 0.1 ** 2 * 0.5)) / max(uo, 0.05) < 0.1  # exact: sum x^2 + d sigma^2 T
    for target in (1, 0, 5000):
        u, st = H._ml_picard_batch(prob, alg, tapes=tapes, target_units=target, handle=handle)
        assert rel_err(float(u[0]), uo) <= 1e-12 and st["path_steps"] == 4916


def test_c2_benchmarked_kernel_matches_oracle_on_its_own_draws(H, O, handle):
    """The launch bench.py times, at 512 solves: every solve against `_ml_picard` of the oracle on the Philox
    draws that solve consumed (hdp_tape_fill), values to 1e-12, and the wall-hit total of the batch exact."""
    prob, alg = _c2(H)
    n, seed = 512, 20261019
    u, st = H._ml_picard_batch(prob, alg, n_solves=n, seed=seed, handle=handle, std_err=True)
    assert st["group_size"] == 1  # one lane per point: the benchmarked instantiation
    assert st["path_steps"] == 4916

    def one(i):
        uo, info = oracle_solve(H, O, prob, alg, np.float64, _tapes(H, handle, prob, alg, seed + i))
        return uo, info["bounces"], info["rc"]

    with cf.ThreadPoolExecutor(8) as ex:
        res = list(ex.map(one, range(n)))
    uo = np.array([r[0] for r in res])
    assert all(r[2] == 0 for r in res)
    rel = np.abs(u - uo) / np.abs(uo)
    assert rel.max() <= 1e-12, rel.max()
    assert st["bounces"] == sum(r[1] for r in res)  # reflection decisions: exact over 2.5 M path-steps
    assert 20 < st["bounces"] / (n * 4916) < 32     # ~26 wall hits per path-step (step sizes are uniform in (0, T-s))
    # the error bar the library reports is the spread of the batch (independent seeds)
    sd = float(np.std(u, ddof=1))
    se = float(np.sqrt(np.mean(st["std_err"] ** 2)))
    assert 0.7 * sd <= se <= 1.4 * sd, (sd, se)
    assert abs(u.mean() - 0.2037) < 0.01


def test_c2_single_solve_every_width_and_cut(H, O, handle):
    prob, alg = _c2(H)
    seed = 77
    uo, info = oracle_solve(H, O, prob, alg, np.float64, _tapes(H, handle, prob, alg, seed))
    for G in (1, 4, 32):
        for target in (1, 600, 0):
            u, st = H._ml_picard_batch(prob, alg, seed=seed, handle=handle, group_size=G, target_units=target)
            assert rel_err(float(u[0]), uo) <= 1e-12 and st["bounces"] == info["bounces"], (G, target)


def test_c5_d1000_l5_matches_oracle(H, O, handle):
    """BASELINE configs[4]: the kernel of the 64-solve batch (8 lanes per point, three vec-groups per pass)."""
    prob, _ = _c2(H, 1000)
    alg = H.MLP(M=4, L=5, K=10)
    seed = 4242
    u, st = H._ml_picard_batch(prob, alg, n_solves=64, seed=seed, handle=handle)
    assert st["group_size"] == 8 and st["path_steps"] == 41652
    i = 63
    uo, info = oracle_solve(H, O, prob, alg, np.float64, _tapes(H, handle, prob, alg, seed + i))
    assert info["rc"] == 0
    assert rel_err(float(u[i]), uo) <= 1e-12, (float(u[i]), uo)
    # the same solve alone, forced onto the batch's kernel: wall hits exact (~400 per path-step)
    u1, st1 = H._ml_picard_batch(prob, alg, n_solves=1, seed=seed + i, handle=handle, group_size=8)
    assert rel_err(float(u1[0]), uo) <= 1e-12 and st1["bounces"] == info["bounces"]
    assert 150 < info["bounces"] / 41652 < 400


def test_c3_exact_shape_host_and_device_pointers(H, O, handle):
    """BASELINE configs[2]: B=8192, d=100, N=10, fp32, scalar box (-1, 1), injected dWall; through host buffers and
    through device pointers (synchronous and HDP_ASYNC + hdp_sync) — the three must agree bit for bit."""
    import torch
    d, B, N = 100, 8192, 10
    rng = np.random.default_rng(2026)
    dW = np.asfortranarray(rng.standard_normal((d, B, N)).astype(np.float32))
    prob = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d, dtype=np.float32), (0.0, 1.0), g=H.SumSq(), f=H.FZero(),
                                 neumann_bc=(-1.0, 1.0))
    y1 = np.zeros((d, B), dtype=np.float32, order="F")
    y0 = np.empty_like(y1, order="F")
    H.sde_loop(prob, y0, y1, N, 0.1, dWall=dW, handle=handle)
    pb, _k = prob.to_c()
    o0, o1, its = O.sde_loop(pb, np.zeros((d, B), dtype=np.float32, order="F"), N, 0.1, dW)
    assert np.max(np.abs(y1 - o1)) <= 1e-5 and np.max(np.abs(y0 - o0)) <= 1e-5
    assert np.all(np.abs(y1) <= 1) and its.max() >= 2  # some step needed a second pass of the reference's loop
    # decisions: a coordinate that the oracle reflected differently would be off by the overshoot, not by rounding
    assert np.mean(np.abs(y1 - o1) > 1e-6) == 0.0
    dev = torch.device("cuda", handle.info()["device"])
    t_dw = torch.from_numpy(dW.ravel(order="F").copy()).to(dev)
    for async_ in (False, True):
        t_y1 = torch.zeros(d * B, dtype=torch.float32, device=dev)
        t_y0 = torch.empty_like(t_y1)
        torch.cuda.synchronize(dev)
        H.sde_loop_device(prob, t_y0.data_ptr(), t_y1.data_ptr(), B, N, 0.1, dtype=np.float32, dWall_ptr=t_dw.data_ptr(),
                          async_=async_, handle=handle)
        handle.sync()
        assert np.array_equal(t_y1.cpu().numpy().reshape(B, d).T, y1)
        assert np.array_equal(t_y0.cpu().numpy().reshape(B, d).T, y0)
    # RNG mode on device pointers == RNG mode on host buffers (same seed, same draw addresses)
    h1 = np.zeros((d, B), dtype=np.float32, order="F")
    h0 = np.empty_like(h1, order="F")
    H.sde_loop(prob, h0, h1, N, 0.1, seed=11, handle=handle)
    t_y1 = torch.zeros(d * B, dtype=torch.float32, device=dev)
    t_y0 = torch.empty_like(t_y1)
    torch.cuda.synchronize(dev)
    H.sde_loop_device(prob, t_y0.data_ptr(), t_y1.data_ptr(), B, N, 0.1, dtype=np.float32, seed=11, async_=True, handle=handle)
    handle.sync()
    assert np.array_equal(t_y1.cpu().numpy().reshape(B, d).T, h1) and np.array_equal(t_y0.cpu().numpy().reshape(B, d).T, h0)
    # a start point outside the box surfaces at hdp_sync in asynchronous mode (src/reflect.jl:59)
    t_y1.fill_(3.0)
    torch.cuda.synchronize(dev)
    H.sde_loop_device(prob, t_y0.data_ptr(), t_y1.data_ptr(), B, N, 0.1, dtype=np.float32, seed=11, async_=True, handle=handle)
    with pytest.raises(H.StartOutsideDomain):
        handle.sync()


def test_mlp_device_pointers_and_async(H, handle):
    """hdp_mlp_solve_* with HDP_MEM_DEVICE (| HDP_ASYNC): x0, u and the error bars stay on the device."""
    import torch
    prob, alg = _c2(H)
    n, seed = 64, 5
    u, st = H._ml_picard_batch(prob, alg, n_solves=n, seed=seed, handle=handle, std_err=True)
    dev = torch.device("cuda", handle.info()["device"])
    x0 = torch.zeros(prob.d, dtype=torch.float64, device=dev)
    for async_ in (False, True):
        du = torch.full((n,), -1.0, dtype=torch.float64, device=dev)
        dse = torch.full((n,), -1.0, dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        H._ml_picard_device(prob, alg, x0.data_ptr(), du.data_ptr(), n_solves=n, seed=seed, std_err_ptr=dse.data_ptr(),
                            async_=async_, handle=handle)
        handle.sync()
        assert np.array_equal(du.cpu().numpy(), u) and np.array_equal(dse.cpu().numpy(), st["std_err"])
    # per-solve start points on the device; one of them outside the box -> reported by hdp_sync
    xs = torch.zeros(prob.d * n, dtype=torch.float64, device=dev)
    xs[5 * prob.d + 3] = 0.75
    torch.cuda.synchronize(dev)
    H._ml_picard_device(prob, alg, xs.data_ptr(), du.data_ptr(), n_solves=n, n_x0=n, seed=seed, async_=True, handle=handle)
    with pytest.raises(H.StartOutsideDomain):
        handle.sync()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_em_ensemble_device_pointers(H, handle, dtype):
    import ctypes as C
    import torch
    d, B, N = 12, 500, 7
    prob = H.ParabolicPDEProblem(H.LinearCoef(0.02), H.LinearCoef(0.2), np.zeros(d, dtype=dtype), (0.0, 0.7), g=H.SumSq(), f=H.FZero())
    x0 = np.asfortranarray(np.random.default_rng(1).uniform(50, 150, (d, B)).astype(dtype))
    want = H.em_ensemble(prob, x0, 0.1, seed=9, saveat=True, handle=handle)
    dev = torch.device("cuda", handle.info()["device"])
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    x = torch.from_numpy(x0.ravel(order="F").copy()).to(dev)
    traj = torch.zeros(d * (N + 1) * B, dtype=tdt, device=dev)
    pb, _k = prob.to_c()
    real = C.c_double if dtype == np.float64 else C.c_float
    fn = handle.lib.hdp_em_ensemble_f64 if dtype == np.float64 else handle.lib.hdp_em_ensemble_f32
    torch.cuda.synchronize(dev)
    H.capi.check(handle.h, fn(handle.h, C.byref(pb), C.c_void_p(x.data_ptr()), B, N, real(0.1), real(0.0),
                              C.c_void_p(traj.data_ptr()), 9, None, H.capi.HDP_MEM_DEVICE | H.capi.HDP_ASYNC))
    handle.sync()
    got = traj.cpu().numpy().reshape(B, N + 1, d).transpose(2, 1, 0)
    assert np.array_equal(got, want)
    assert np.array_equal(x.cpu().numpy().reshape(B, d).T, want[:, -1, :])


@pytest.mark.parametrize("dtype,d,B", [(np.float32, 100, 8192), (np.float64, 7, 33)])
def test_sde_loop_epochs_equal_separate_calls(H, handle, dtype, d, B):
    """hdp_sde_loop_epochs_*: E epochs of DeepSplitting paths in one launch (src/DeepSplitting.jl:181-195); slice e must
    be bit for bit the single call with seed + e — the C3 shape, and a shape whose slices are not 16-byte aligned."""
    import torch
    N, E, K = 10, 5, 2
    prob = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d, dtype=dtype), (0.0, 1.0), g=H.SumSq(), f=H.FZero(),
                                 neumann_bc=(-1.0, 1.0), x0_sample=H.UniformSampling(-0.25, 0.25))
    mc = H.NormalSampling(0.1, True)
    y1 = np.zeros((d, B, E), dtype=dtype, order="F")
    y0 = np.empty_like(y1, order="F")
    z = np.empty((d, B, K, E), dtype=dtype, order="F")
    H.sde_loop_epochs(prob, y0, y1, N, 0.1, seed=40, mc_sample=mc, z=z, handle=handle)
    for e in range(E):
        a1 = np.zeros((d, B), dtype=dtype, order="F")
        a0 = np.empty_like(a1, order="F")
        az = np.empty((d, B, K), dtype=dtype, order="F")
        H.sde_loop(prob, a0, a1, N, 0.1, seed=40 + e, mc_sample=mc, z=az, handle=handle)
        assert np.array_equal(y1[:, :, e], a1) and np.array_equal(y0[:, :, e], a0) and np.array_equal(z[:, :, :, e], az)
    assert not np.array_equal(y1[:, :, 0], y1[:, :, 1])
    # device-resident, asynchronous
    dev = torch.device("cuda", handle.info()["device"])
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    t1 = torch.zeros(d * B * E, dtype=tdt, device=dev)
    t0 = torch.empty_like(t1)
    tz = torch.empty(d * B * K * E, dtype=tdt, device=dev)
    torch.cuda.synchronize(dev)
    H.sde_loop_epochs(prob, None, None, N, 0.1, seed=40, mc_sample=mc, z=K, device_ptrs=(t0.data_ptr(), t1.data_ptr(), tz.data_ptr()),
                      B=B, n_epochs=E, dtype=dtype, async_=True, handle=handle)
    handle.sync()
    assert np.array_equal(t1.cpu().numpy(), y1.ravel(order="F")) and np.array_equal(t0.cpu().numpy(), y0.ravel(order="F"))
    assert np.array_equal(tz.cpu().numpy(), z.ravel(order="F"))


def test_wide_root_multi_block_combine_and_nonlocal_fine_cut(H, O, handle):
    """Code paths of the launch chain that only large single solves reach, against the oracle on the solve's own draws:
    (a) a root with more than 4096 units — its partial sums go through mlp_prereduce_kernel (64 blocks) before the
    combine, and the error bar with them; (b) a non-local problem cut so finely that level-1 sites become explicit
    with whole-call children, and its geometry passes run both as one lane group per (site, point) (short passes) and
    per site (passes of >= 32768 sites).  Every cut must give the oracle's value and wall-hit count."""
    # (a) Allen-Cahn d=6, M=8, L=5 (1,184,968 path-steps): 32768 level-0 sites, 32768 g-samples and 4096 level-1 sites at
    # the root — 4096 + 4096 + 4096 units of it at the finest cut (8 path-steps per unit)
    d = 6
    prob = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                 neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
    alg = H.MLP(M=8, L=5, K=3)
    seed = 9
    uo, info = oracle_solve(H, O, prob, alg, np.float64, _tapes(H, handle, prob, alg, seed))
    assert info["rc"] == 0
    ses = []
    for target in (1, 3000, 10 ** 7):
        u, st = H._ml_picard_batch(prob, alg, seed=seed, handle=handle, target_units=target, std_err=True)
        assert rel_err(float(u[0]), uo) <= 1e-12 and st["bounces"] == info["bounces"], (target, float(u[0]), uo)
        ses.append(float(st["std_err"][0]))
    assert st["n_units"] > 8192                      # the finest cut: the root alone holds more than 4096 units
    assert max(ses) - min(ses) <= 1e-9 * max(ses)    # the same error bar through one unit, a warp, and 64 blocks

    # (b) non-local Fisher-KPP d=4, M=3, L=4, K=8 (174,594 path-steps, 709,872 sampler calls per solve), 16 solves per
    # call so that the deepest geometry pass of the fine cut (2874 sites x 16) takes the per-site form
    d = 4
    prob = H.PIDEProblem(0.0, 0.3, np.zeros(d), (0.0, 0.5), H.ExpSumSq(1.0), H.FNLFisher(),
                         neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
    alg = H.MLP(M=3, L=4, K=8, mc_sample=H.UniformSampling(np.full(d, -0.5), np.full(d, 0.5)))
    cn = H.mlp_counts(alg.M, alg.L, alg.K, True)
    seed, n = 5, 16
    want = {}
    for i in (0, 11):
        r = handle.tape_fill(np.float64, H.capi.HDP_STREAM_TIME, False, seed + i, 0, cn["time_uniforms"], 1)
        nn = handle.tape_fill(np.float64, H.capi.HDP_STREAM_STEP, True, seed + i, 0, cn["path_steps"], d)
        ss = handle.tape_fill(np.float64, H.capi.HDP_STREAM_SAMPLER, False, seed + i, 0, cn["sampler_calls"], d)
        want[i] = oracle_solve(H, O, prob, alg, np.float64, (r, nn, ss))
        assert want[i][1]["rc"] == 0
    work = cn["path_steps"] + 0.5 * cn["sampler_calls"]
    sites = []
    for target in (0, int(n * work / 2000.0), int(n * work / 24.0)):   # thresh = W[L], 2000, 22.6 (W[1] = 18 < 22.6 < Ws[1] / 6)
        u, st = H._ml_picard_batch(prob, alg, n_solves=n, seed=seed, handle=handle, target_units=target)
        for i, (uo, info) in want.items():
            assert rel_err(float(u[i]), uo) <= 1e-12, (target, i, float(u[i]), uo)
        sites.append(st["n_sites"])
        u1, st1 = H._ml_picard_batch(prob, alg, n_solves=1, seed=seed + 11, handle=handle,
                                     target_units=max(1, target // n))
        assert rel_err(float(u1[0]), want[11][0]) <= 1e-12 and st1["bounces"] == want[11][1]["bounces"], target
    assert sites[-1] >= 2000 and sites[-1] * n >= 32768   # the finest cut made the level-1 sites explicit
