"""The product's own K1 device code (highdimpde.jl_b200/csrc/k_sde.cuh: `sde_loop!`, `_mlt_sde_loop!` and the EM
ensemble), compiled for the host by tests/hostemu (one emulated thread, per-element loads), against the oracle —
the step loops of every mode, the three associations and the draw addressing are checked on every CPU run, not
only on a GPU box.  Test infrastructure; the emulation is never part of libhdpde_b200.so."""
import ctypes as C

import numpy as np
import pytest

TOL = {np.float64: 1e-12, np.float32: 1e-5}


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _prob(H, mu, sigma, d, dtype, bc=None, tspan=(0.0, 1.0)):
    return H.ParabolicPDEProblem(mu, sigma, np.zeros(d, dtype=dtype), tspan, g=H.SumSq(), f=H.FZero(), neumann_bc=bc)


def emu_sde_loop(H, emu, prob, y1, nsteps, dt, *, ts=None, dW=None, seed=0, x0_sample=None, mlp_assoc=0, force_generic=0):
    dtp = y1.dtype
    real = C.c_double if dtp == np.float64 else C.c_float
    fn = emu.hdp_emu_sde_loop_f64 if dtp == np.float64 else emu.hdp_emu_sde_loop_f32
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, real, C.c_void_p, C.c_void_p, C.c_uint64,
                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ulonglong)]
    fn.restype = C.c_int
    pb, _k = prob.to_c()
    d, B = y1.shape
    x0s, _k2 = (x0_sample or H.NoSampling())._descriptor(d)
    none, _k3 = H.NoSampling()._descriptor(d)
    y1 = np.array(y1, order="F")
    y0 = np.empty_like(y1, order="F")
    tsa = None if ts is None else np.ascontiguousarray(ts, dtype=dtp)
    dw = None if dW is None else np.asfortranarray(dW, dtype=dtp)
    nb = C.c_ulonglong(0)
    rc = fn(C.byref(pb), _p(y0), _p(y1), B, nsteps, real(dt), _p(tsa), C.byref(x0s), seed, _p(dw), C.byref(none), None, 0,
            mlp_assoc, force_generic, C.byref(nb))
    return rc, y0, y1, nb.value


def emu_em_ensemble(H, emu, prob, x0, nsteps, dt, t0, *, xi=None, seed=0, saveat=False):
    dtp = x0.dtype
    real = C.c_double if dtp == np.float64 else C.c_float
    fn = emu.hdp_emu_em_ensemble_f64 if dtp == np.float64 else emu.hdp_emu_em_ensemble_f32
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, real, real, C.c_void_p, C.c_uint64, C.c_void_p]
    fn.restype = C.c_int
    pb, _k = prob.to_c()
    d, B = x0.shape
    x = np.array(x0, order="F")
    traj = np.empty((d, nsteps + 1, B), dtype=dtp, order="F") if saveat else None
    z = None if xi is None else np.asfortranarray(xi, dtype=dtp)
    rc = fn(C.byref(pb), _p(x), B, nsteps, real(dt), real(t0), _p(traj), seed, _p(z))
    assert rc == 0
    return x, traj


CASES = {
    "const_sym": (0.05, 1.0, (-1.0, 1.0)),                       # the BASELINE shape: fast symmetric-box loop
    "const_nobc": (0.05, 1.0, None),                             # fast loop without a boundary
    "const_box": (0.05, 1.0, (-0.75, 1.0)),                      # general scalar box -> generic loop
    "const_vecbox": (0.0, 1.0, "vec"),                           # vector bounds -> generic loop
    "black_scholes": ("lin0.02", "lin0.2", None),                # test/DeepSplitting.jl:470-471
    "quad_box": (0.0, "quad0.4", (-1.0, 1.0)),                   # test/DeepSplitting.jl:355
}


def _coef(H, v):
    if isinstance(v, str):
        return H.LinearCoef(float(v[3:])) if v.startswith("lin") else H.QuadraticCoef(float(v[4:]))
    return v


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("name", list(CASES))
def test_emulated_sde_loop_matches_oracle(H, O, emu, dtype, name):
    mu, sg, bc = CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    d, B, N, dt = 7, 19, 6, 0.1
    if bc == "vec":
        bc = (np.linspace(-1.0, -0.5, d), np.linspace(0.5, 1.0, d))
    prob = _prob(H, _coef(H, mu), _coef(H, sg), d, dtype, bc)
    y1 = np.asfortranarray(rng.uniform(-0.45, 0.45, (d, B)).astype(dtype))
    dW = rng.standard_normal((d, B, N)).astype(dtype)
    o0, o1, _ = O.sde_loop(prob.to_c()[0], y1, N, dt, dW)
    for force in (0, 1):  # the fast loops and the generic loop give the same bits for constant coefficients
        rc, y0, yf, nb = emu_sde_loop(H, emu, prob, y1, N, dt, dW=dW, force_generic=force)
        assert rc == 0
        if bc is None:
            assert np.array_equal(yf, o1) and np.array_equal(y0, o0)  # same IEEE operations in the same order
        else:
            assert np.max(np.abs(yf - o1)) <= TOL[dtype] and np.max(np.abs(y0 - o0)) <= TOL[dtype]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_emulated_sde_loop_philox_addresses_match_oracle(H, O, emu, dtype):
    """RNG mode: draw(SDE, ordinal = step, i = element) and draw(X0, 0, i) — the oracle is fed the same stream."""
    d, B, N = 6, 9, 4
    prob = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d, dtype=dtype), (0.0, 1.0), g=H.SumSq(), f=H.FZero(),
                                 neumann_bc=(-1.0, 1.0), x0_sample=H.UniformSampling(-0.25, 0.25))
    rc, y0, y1, nb = emu_sde_loop(H, emu, prob, np.zeros((d, B), dtype=dtype, order="F"), N, 0.1, seed=5,
                                  x0_sample=prob.x0_sample)
    assert rc == 0
    dW = O.philox_tape(dtype, 5, H.capi.HDP_STREAM_SDE, 0, N, d * B, True)  # [step][q]
    dW = np.asfortranarray(dW.reshape(N, B, d).transpose(2, 1, 0))
    x0raw = O.philox_tape(dtype, 5, H.capi.HDP_STREAM_X0, 0, 1, d * B, False).reshape(B, d).T
    sm, _k = prob.x0_sample._descriptor(d)
    o0, o1, _ = O.sde_loop(prob.to_c()[0], np.zeros((d, B), dtype=dtype, order="F"), N, 0.1, dW, sm, np.asfortranarray(x0raw))
    tol = 1e-11 if dtype == np.float64 else 1e-5
    assert np.max(np.abs(y1 - o1)) <= tol and np.max(np.abs(y0 - o0)) <= tol


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_emulated_mlt_sde_loop_association(H, O, emu, dtype):
    """One step with the MLP association x + (mu dt + (sigma sqrt(dt)) xi), coefficients at the END time (src/MLP.jl:417)."""
    rng = np.random.default_rng(12)
    d, B = 5, 11
    prob = _prob(H, H.LinearCoef(0.3), 0.8, d, dtype)
    x = np.asfortranarray(rng.uniform(-0.5, 0.5, (d, B)).astype(dtype))
    xi = rng.standard_normal((d, B, 1)).astype(dtype)
    s, t = dtype(0.1), dtype(0.45)
    rc, y0, x2, _ = emu_sde_loop(H, emu, prob, x, 1, dtype(t - s), ts=np.array([t]), dW=xi, mlp_assoc=1)
    pb, _k = prob.to_c()
    want = np.stack([O.mlt_sde_loop(pb, x[:, j].copy(), s, t, xi[:, j, 0].copy()) for j in range(B)], axis=1)
    assert rc == 0 and np.array_equal(x2, want) and np.array_equal(y0, x)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_emulated_em_ensemble_matches_oracle(H, O, emu, dtype):
    rng = np.random.default_rng(3)
    d, B, N, T = 5, 13, 6, 0.75
    x0 = np.asfortranarray(rng.uniform(0.5, 1.5, (d, B)).astype(dtype))
    xi = np.asfortranarray(rng.standard_normal((d, B, N)).astype(dtype))
    for mu, sg in ((-2.0, 2.0), (H.LinearCoef(0.125), H.LinearCoef(0.5)), (0.05, H.QuadraticCoef(0.4)),
                   (H.QuadraticCoef(-0.1), H.LinearCoef(0.3))):
        prob = _prob(H, mu, sg, d, dtype, tspan=(0.0, T))
        want_x, want_traj = O.em_ensemble(prob.to_c()[0], x0, N, dtype(T / N), dtype(0.0), xi, saveat=True)
        x, traj = emu_em_ensemble(H, emu, prob, x0, N, T / N, 0.0, xi=xi, saveat=True)
        assert np.array_equal(traj, want_traj) and np.array_equal(x, want_x)
        x_only, none = emu_em_ensemble(H, emu, prob, x0, N, T / N, 0.0, xi=xi)
        assert none is None and np.array_equal(x_only, want_x)
    # in-register normals: the same addresses as sde_loop!
    prob = _prob(H, H.LinearCoef(0.02), H.LinearCoef(0.2), d, dtype, tspan=(0.0, T))
    x, traj = emu_em_ensemble(H, emu, prob, x0, N, T / N, 0.0, seed=21, saveat=True)
    z = O.philox_tape(dtype, 21, H.capi.HDP_STREAM_SDE, 0, N, d * B, True)
    z = np.asfortranarray(z.reshape(N, B, d).transpose(2, 1, 0))
    _, want = O.em_ensemble(prob.to_c()[0], x0, N, dtype(T / N), dtype(0.0), z, saveat=True)
    tol = 1e-11 if dtype == np.float64 else 1e-5
    assert np.max(np.abs(traj - want)) <= tol * float(np.max(np.abs(want)))


def test_emulated_sde_loop_start_outside_and_empty(H, emu):
    prob = _prob(H, 0.0, 1.0, 4, np.float64, (-1.0, 1.0))
    rc, *_ = emu_sde_loop(H, emu, prob, np.full((4, 3), 3.0, order="F"), 2, 0.1, dW=np.zeros((4, 3, 2)))
    assert rc == H.capi.HDP_ERR_START_OUTSIDE_DOMAIN  # @assert of src/reflect.jl:59
    rc, y0, y1, nb = emu_sde_loop(H, emu, prob, np.zeros((4, 0), order="F"), 2, 0.1, dW=np.zeros((4, 0, 2)))
    assert rc == 0 and y1.shape == (4, 0)
