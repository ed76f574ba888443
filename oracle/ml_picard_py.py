"""A second, independent restatement of the reference's MLP recursion — TEST INFRASTRUCTURE ONLY.

Pure Python over Float64 (small cases: loops, no vectorisation tricks), written from the reference source and not
from oracle/hdp_oracle_impl.h, and — unlike the C restatement and the CUDA path — it takes *closures* for
mu, sigma, g, f exactly as `solve(prob, MLP())` does.  tests/test_oracle_pins.py requires the C oracle to agree with
it bit for bit on shared random tapes, so a misreading of the recursion would have to be made twice, in two
languages, to go unnoticed.  (The reference itself is pure Julia and cannot run in this image.)

    _ml_picard       src/MLP.jl:102-208   (and :210-216 for x === nothing)
    _ml_picard_mlt   src/MLP.jl:218-260, _ml_picard_call :262-388, _get_loop_num :391-404
    _mlt_sde_loop!   src/MLP.jl:406-421
    _reflect         src/reflect.jl:8-55  (Vector method, the literal sequential billiard)
    samplers         src/MCSample.jl:28-33, 63-71, 86
"""
from __future__ import annotations

import math

import numpy as np


class Tape:
    """rand(tType) / randn!(x2) / rand!-randn!(x_mc) served from arrays, in consumption order."""

    def __init__(self, r, n, s=None):
        self.r, self.n, self.s = np.asarray(r, dtype=np.float64), np.asarray(n, dtype=np.float64), s
        self.ir = self.in_ = self.is_ = 0

    def rand(self):
        v = self.r[self.ir]
        self.ir += 1
        return float(v)

    def randn(self, d):
        v = self.n[self.in_]
        self.in_ += 1
        assert v.shape == (d,)
        return v.copy()

    def sampler_raw(self, d):
        v = np.asarray(self.s[self.is_], dtype=np.float64)
        self.is_ += 1
        assert v.shape == (d,)
        return v.copy()


class NoSampling:  # src/MCSample.jl:84-86
    integrate = False

    def __call__(self, tape, x3, x2):
        return None


class UniformSampling:  # src/MCSample.jl:22-33
    integrate = True

    def __init__(self, a, b):
        self.a, self.b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)

    def __call__(self, tape, x3, x2):
        u = tape.sampler_raw(x3.shape[0])            # rand!(x_mc)                       :30
        m = (self.b + self.a) / 2.0                  # m = (b + a) ./ 2                  :31
        x3[:] = (u - 0.5) * (self.b - self.a) + m    # (x_mc .- 0.5) .* (b - a) .+ m     :32
        return x3


class NormalSampling:  # src/MCSample.jl:55-71
    integrate = True

    def __init__(self, sigma, shifted=False):
        self.sigma, self.shifted = float(sigma), bool(shifted)

    def __call__(self, tape, x3, x2):
        x3[:] = tape.sampler_raw(x3.shape[0])        # randn!(x_mc)                      :64
        x3 *= self.sigma                             # x_mc .*= σ                        :65
        if self.shifted:
            x3 += x2                                 # x_mc .+= x                        :70
        return x3


def reflect_vector(a, b, s, e):
    """src/reflect.jl:8-55, line by line; returns (b, number of wall hits)."""
    a, b = np.array(a, dtype=np.float64), np.array(b, dtype=np.float64)
    d = a.shape[0]
    r = 2.0                                                           # :9
    n = np.zeros(d)                                                   # :10
    if not np.all((a >= s) & (a <= e)):                               # :12
        raise ValueError("a not in hypercube")

    def scan(a, b, r, n):                                             # :14-30 / :36-52
        for i in range(d):
            if b[i] < s[i]:
                rtemp = (a[i] - s[i]) / (a[i] - b[i])
                if rtemp < r:
                    r = rtemp
                    n[:] = 0.0
                    n[i] = -1.0
            elif b[i] > e[i]:
                rtemp = (e[i] - a[i]) / (b[i] - a[i])
                if rtemp < r:
                    r = rtemp
                    n[:] = 0.0
                    n[i] = 1.0
        return r

    r = scan(a, b, r, n)
    hits = 0
    while r < 1:                                                      # :31
        c = a + r * (b - a)                                           # :32
        a = c                                                         # :33
        dot = 0.0
        for i in range(d):                                            # dot(b - c, n): one non-zero term
            dot += (b[i] - c[i]) * n[i]
        b = b - 2.0 * n * dot                                         # :34
        hits += 1
        r = scan(a, b, 2.0, n)                                        # :35-52
    return b, hits


class Problem:
    """The fields `_ml_picard` reads from prob: μ(x,p,t), σ(x,p,t) (scalar or length-d), p, and the closures g, f."""

    def __init__(self, mu, sigma, g, f, tspan, p=None, neumann_bc=None):
        self.mu, self.sigma, self.g, self.f, self.tspan, self.p = mu, sigma, g, f, tspan, p
        self.neumann_bc = None if neumann_bc is None else tuple(np.asarray(v, dtype=np.float64) for v in neumann_bc)


def mlt_sde_loop(tape, x2, x, s, t, prob, stats):
    """src/MLP.jl:406-421"""
    dt = t - s                                                        # :415
    xi = tape.randn(x.shape[0])                                       # :416 randn!(x2)
    # :417  x2 .= x + (μ .* dt .+ σ .* sqrt(dt) .* x2)   with  (σ .* sqrt(dt)) .* x2  (left to right)
    x2[:] = x + (prob.mu(x, prob.p, t) * dt + (prob.sigma(x, prob.p, t) * math.sqrt(dt)) * xi)
    stats["steps"] += 1
    if prob.neumann_bc is not None:                                   # :418-420
        b, hits = reflect_vector(x, x2, *prob.neumann_bc)
        x2[:] = b
        stats["bounces"] += hits


def ml_picard(M, L, K, x, s, t, mc_sample, tape, prob, stats):
    """src/MLP.jl:102-208; f is called as f(x2, x3, b2, v3, nothing, nothing, p, t) with the OUTER end time t."""
    if x is None:                                                     # :210-216
        return None
    if L == 0:                                                        # :117-119
        return 0.0
    d = x.shape[0]
    x2 = np.empty(d)                                                  # :121
    x3 = np.empty(d) if mc_sample.integrate else None                 # :122
    p, g, f = prob.p, prob.g, prob.f
    a = 0.0
    for l in range(0, min(L - 1, 1) + 1):                             # :126
        b = 0.0
        num = M ** (L - l)
        for _k in range(num):                                         # :130
            r = s + (t - s) * tape.rand()                             # :132
            mlt_sde_loop(tape, x2, x, s, r, prob, stats)              # :133
            b2 = ml_picard(M, l, K, x2, r, t, mc_sample, tape, prob, stats)  # :134
            b3 = 0.0
            for _h in range(K):                                       # :137
                z = mc_sample(tape, x3, x2)                           # :139
                b3 += f(x2, z, b2, ml_picard(M, l, K, z, r, t, mc_sample, tape, prob, stats), None, None, p, t)  # :140-146
            b += b3 / K                                               # :148
        a += (t - s) * b / num                                        # :150
    for l in range(2, L):                                             # :153
        b = 0.0
        num = M ** (L - l)
        for _k in range(num):
            r = s + (t - s) * tape.rand()                             # :157
            mlt_sde_loop(tape, x2, x, s, r, prob, stats)              # :158
            b2 = ml_picard(M, l, K, x2, r, t, mc_sample, tape, prob, stats)      # :159
            b4 = ml_picard(M, l - 1, K, x2, r, t, mc_sample, tape, prob, stats)  # :160-173
            b3 = 0.0
            for _h in range(K):                                       # :176
                z = mc_sample(tape, x3, x2)                           # :177
                first = f(x2, z, b2, ml_picard(M, l, K, z, r, t, mc_sample, tape, prob, stats), None, None, p, t)
                second = f(x2, z, b4, ml_picard(M, l - 1, K, z, r, t, mc_sample, tape, prob, stats), None, None, p, t)
                b3 += first - second                                  # :178-191
            b += b3 / K                                               # :193
        a += (t - s) * b / num                                        # :195
    num = M ** L                                                      # :198
    a2 = 0.0
    for _k in range(num):
        mlt_sde_loop(tape, x2, x, s, t, prob, stats)                  # :202
        a2 += g(x2)                                                   # :203
    a2 /= num                                                         # :205
    return a + a2                                                     # :207


def solve(prob, M, L, K, x, mc_sample, tape):
    """The serial branch of solve(::MLP), src/MLP.jl:84-98: u(t0, x)."""
    stats = {"steps": 0, "bounces": 0}
    u = ml_picard(M, L, K, np.asarray(x, dtype=np.float64), prob.tspan[0], prob.tspan[1], mc_sample, tape, prob, stats)
    return u, stats


# ------------------------------ the threaded variant, src/MLP.jl:218-404 ------------------------------
def get_loop_num(M, num, thread_id, nthreads):
    """src/MLP.jl:391-404, literally: float division, `M % num`; `for k in 1:loop_num` runs floor(loop_num) times."""
    if num < nthreads:
        return 0 if thread_id > num else 1
    remainder = M % num
    if remainder > 0 and thread_id <= remainder:
        loop_num = num / nthreads + 1
    else:
        loop_num = num / nthreads
    return int(math.floor(loop_num))


def ml_picard_call(M, L, K, x, s, t, mc_sample, tape, nthreads, thread_id, prob, stats):
    """src/MLP.jl:262-388: the level loops of one task; the sums are still divided by `num`, not by the number of
    iterations actually run (quirk Q4 of SURVEY.md)."""
    d = x.shape[0]
    x2 = np.empty(d)
    x3 = np.empty(d) if mc_sample.integrate else None
    p, f = prob.p, prob.f
    a = 0.0
    for l in range(0, min(L - 1, 1) + 1):                             # :284
        b = 0.0
        num = M ** (L - l)
        for _k in range(get_loop_num(M, num, thread_id, nthreads)):   # :287-288
            r = s + (t - s) * tape.rand()
            mlt_sde_loop(tape, x2, x, s, r, prob, stats)
            b2 = ml_picard(M, l, K, x2, r, t, mc_sample, tape, prob, stats)
            b3 = 0.0
            for _h in range(K):
                z = mc_sample(tape, x3, x2)
                b3 += f(x2, z, b2, ml_picard(M, l, K, z, r, t, mc_sample, tape, prob, stats), None, None, p, t)
            b += b3 / K
        a += (t - s) * b / num                                        # :307
    for l in range(2, L):                                             # :310
        b = 0.0
        num = M ** (L - l)
        for _k in range(get_loop_num(M, num, thread_id, nthreads)):
            r = s + (t - s) * tape.rand()
            mlt_sde_loop(tape, x2, x, s, r, prob, stats)
            b2 = ml_picard(M, l, K, x2, r, t, mc_sample, tape, prob, stats)
            b4 = ml_picard(M, l - 1, K, x2, r, t, mc_sample, tape, prob, stats)
            b3 = 0.0
            for _h in range(K):
                z = mc_sample(tape, x3, x2)
                first = f(x2, z, b2, ml_picard(M, l, K, z, r, t, mc_sample, tape, prob, stats), None, None, p, t)
                second = f(x2, z, b4, ml_picard(M, l - 1, K, z, r, t, mc_sample, tape, prob, stats), None, None, p, t)
                b3 += first - second
            b += b3 / K
        a += (t - s) * b / num                                        # :384
    return a


def solve_mlt(prob, M, L, K, x, mc_sample, tape, nthreads):
    """`_ml_picard_mlt` (src/MLP.jl:218-260) with Threads.nthreads() == nthreads, the tasks run one after the other
    on ONE shared tape (tasks 1..NT, then the g-term of the main task) — the order the C oracle uses when it is
    given a single sequential tape.  (With task-local RNGs the reference's draws are not ordered across tasks.)"""
    stats = {"steps": 0, "bounces": 0}
    x = np.asarray(x, dtype=np.float64)
    s, t = prob.tspan
    parts = [ml_picard_call(M, L, K, x, s, t, mc_sample, tape, nthreads, tid, prob, stats) for tid in range(1, nthreads + 1)]
    num = M ** L                                                      # :246
    x2 = np.empty(x.shape[0])
    a2 = 0.0
    for _k in range(num):
        mlt_sde_loop(tape, x2, x, s, t, prob, stats)                  # :251
        a2 += prob.g(x2)
    a2 /= num                                                         # :254
    a = 0.0
    for v in parts:                                                   # :257 sum([fetch(t) for t in tasks])
        a += v
    return a + a2, stats                                              # :259
