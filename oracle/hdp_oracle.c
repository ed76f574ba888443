/*
 * hdp_oracle.c — CPU oracle for the HighDimPDE.jl Monte-Carlo hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (`highdimpde.jl_b200/`,
 * libhdpde_b200.so) may import, link or execute this file; only `tests/`,
 * `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
 * legs of `bench.py` use it, and only as the checker / the CPU baseline.
 *
 * What it is: a plain-C, operation-for-operation restatement of
 *   src/MLP.jl:102-421      (_ml_picard, _ml_picard_mlt, _ml_picard_call,
 *                            _get_loop_num, _mlt_sde_loop!)
 *   src/reflect.jl:8-82     (_reflect, Vector and AbstractArray methods)
 *   src/MCSample.jl:22-94   (UniformSampling, NormalSampling, NoSampling)
 *   src/DeepSplitting.jl:153-166 (sde_loop! closure)
 * of SciML/HighDimPDE.jl v2.6.0 (bodies in hdp_oracle_impl.h).
 *
 * PARITY STATUS: "parity unpinned" at the RNG boundary.  The reference is
 * pure Julia; no Julia toolchain exists in this image, so the reference
 * cannot be executed here and its tests hold no golden vectors (all are
 * seed-100 statistical checks).  The restatement is pinned against every
 * closed-form answer those tests use (tests/test_oracle_pins.py).
 *
 * Random numbers are consumed through callbacks in the reference's own
 * consumption order (rand(tType) → randn!(x2) → mc_sample! …), so a
 * sequential tape, or a counter-based generator addressed by consumption
 * ordinal, reproduces any run exactly.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/hdpde.h"
#include "hdp_oracle.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ================================================================ */
/* Philox4x32-R (Salmon et al., SC'11), generic in counter, key, rounds. */
/* The CUDA library's draws use a FIXED key and carry the 64-bit seed   */
/* in the counter: ctr = (w0, w1, seed_lo, seed_hi) with (w0, w1) the   */
/* packed address (stream, ordinal, group) — see draw_address() in      */
/* highdimpde.jl_b200/csrc/hdp_device.cuh.                              */
/* ================================================================ */
void hdp_oracle_philox4x32(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < rounds; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
int hdp_oracle_philox_rounds(void) { return HDP_PHILOX_ROUNDS; }
/* the 10-round member of the family: the one Random123 publishes known-answer vectors for */
void hdp_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    hdp_oracle_philox4x32(ctr, key, 10, out);
}

/* the library's draw: streams TIME/STEP/SAMPLER pack (ordinal < 2^40, group < 2^21),
 * streams SDE/X0/MC pack (ordinal < 2^24, group < 2^37) */
static void draw_words(uint64_t seed, int stream, uint64_t ord, uint64_t group, uint32_t w[4])
{
    static const uint32_t key[2] = {0xA4093822u, 0x299F31D0u};
    uint32_t ctr[4];
    if (stream <= HDP_STREAM_SAMPLER) {
        ctr[0] = (uint32_t)ord;
        ctr[1] = (uint32_t)(ord >> 32) | ((uint32_t)group << 8) | ((uint32_t)stream << 29);
    } else {
        ctr[0] = (uint32_t)group;
        ctr[1] = (uint32_t)(group >> 32) | ((uint32_t)ord << 5) | ((uint32_t)stream << 29);
    }
    ctr[2] = (uint32_t)seed;
    ctr[3] = (uint32_t)(seed >> 32);
    hdp_oracle_philox4x32(ctr, key, HDP_PHILOX_ROUNDS, w); /* the library's round count (include/hdpde.h) */
}

static inline uint64_t k52(uint32_t hi, uint32_t lo)
{ /* 52-bit integer k = (hi >> 12) : lo in [0, 2^52) */
    return ((uint64_t)(hi >> 12) << 32) | (uint64_t)lo;
}

/* value i of ordinal `ord` in `stream`; element type f64: one Philox call yields 2 draws
 * (group = i>>1).  Uniform u = k*2^-52 in [0,1).  Box–Muller: u1 = (k1 | 1)*2^-52 in (0,1),
 * angle 2*pi*k2/2^52.  Same definitions as hdp_device.cuh (which evaluates the
 * transcendentals with its own tables and polynomials: the two agree to ~1e-15, not bit
 * for bit). */
double hdp_oracle_draw_f64(uint64_t seed, int stream, uint64_t ord, uint64_t i, int want_normal)
{
    uint32_t w[4];
    draw_words(seed, stream, ord, i >> 1, w);
    if (!want_normal) {
        return (i & 1) ? (double)k52(w[2], w[3]) * 0x1.0p-52 : (double)k52(w[0], w[1]) * 0x1.0p-52;
    }
    double u1 = (double)(k52(w[0], w[1]) | 1u) * 0x1.0p-52;
    double turn = (double)k52(w[2], w[3]) * 0x1.0p-52;
    double r = sqrt(-2.0 * log(u1));
    double th = 2.0 * M_PI * turn;
    return (i & 1) ? r * sin(th) : r * cos(th);
}

/* element type f32: one Philox call yields 4 draws (group = i>>2); 23-bit uniforms
 * u = (w >> 9)*2^-23; normals pair (w0,w1)->(z0,z1), (w2,w3)->(z2,z3). */
float hdp_oracle_draw_f32(uint64_t seed, int stream, uint64_t ord, uint64_t i, int want_normal)
{
    uint32_t w[4];
    draw_words(seed, stream, ord, i >> 2, w);
    uint32_t e = (uint32_t)(i & 3);
    if (!want_normal) return (float)(w[e] >> 9) * 0x1.0p-23f;
    uint32_t wa = w[e & 2], wb = w[(e & 2) + 1];
    float u1 = ((float)(wa >> 9) + 0.5f) * 0x1.0p-23f;
    float u2 = (float)(wb >> 9) * 0x1.0p-23f;
    float r = sqrtf(-2.0f * logf(u1));
    float th = 2.0f * (float)M_PI * u2;
    return (e & 1) ? r * sinf(th) : r * cosf(th);
}

/* ---------------------------------------------------------------- */
/* rng sources                                                       */
/* ---------------------------------------------------------------- */
typedef struct tape_state {
    int is_f32;
    const void *r, *n, *s;
    uint64_t ri, ni, si;
    uint64_t rn, nn, sn; /* capacities (elements) */
    int overrun;
} tape_state;

static inline double tape_get(const void* p, int is_f32, uint64_t i)
{
    return is_f32 ? (double)((const float*)p)[i] : ((const double*)p)[i];
}
static double tape_uniform(void* c)
{
    tape_state* t = (tape_state*)c;
    if (t->ri >= t->rn) { t->overrun = 1; return 0.0; }
    return tape_get(t->r, t->is_f32, t->ri++);
}
static void tape_normals(void* c, double* out, int d)
{
    tape_state* t = (tape_state*)c;
    if (t->ni + (uint64_t)d > t->nn) { t->overrun = 1; memset(out, 0, sizeof(double) * (size_t)d); return; }
    for (int i = 0; i < d; ++i) out[i] = tape_get(t->n, t->is_f32, t->ni + (uint64_t)i);
    t->ni += (uint64_t)d;
}
static void tape_sampler(void* c, double* out, int d)
{
    tape_state* t = (tape_state*)c;
    if (t->si + (uint64_t)d > t->sn) { t->overrun = 1; memset(out, 0, sizeof(double) * (size_t)d); return; }
    for (int i = 0; i < d; ++i) out[i] = tape_get(t->s, t->is_f32, t->si + (uint64_t)i);
    t->si += (uint64_t)d;
}

typedef struct philox_state {
    int is_f32, sampler_normal;
    uint64_t seed, r_ord, n_ord, s_ord;
} philox_state;
static double philox_uniform(void* c)
{
    philox_state* p = (philox_state*)c;
    uint64_t o = p->r_ord++;
    return p->is_f32 ? (double)hdp_oracle_draw_f32(p->seed, HDP_STREAM_TIME, o, 0, 0)
                     : hdp_oracle_draw_f64(p->seed, HDP_STREAM_TIME, o, 0, 0);
}
static void philox_normals(void* c, double* out, int d)
{
    philox_state* p = (philox_state*)c;
    uint64_t o = p->n_ord++;
    for (int i = 0; i < d; ++i)
        out[i] = p->is_f32 ? (double)hdp_oracle_draw_f32(p->seed, HDP_STREAM_STEP, o, (uint64_t)i, 1)
                           : hdp_oracle_draw_f64(p->seed, HDP_STREAM_STEP, o, (uint64_t)i, 1);
}
static void philox_sampler(void* c, double* out, int d)
{
    philox_state* p = (philox_state*)c;
    uint64_t o = p->s_ord++;
    for (int i = 0; i < d; ++i)
        out[i] = p->is_f32
                     ? (double)hdp_oracle_draw_f32(p->seed, HDP_STREAM_SAMPLER, o, (uint64_t)i, p->sampler_normal)
                     : hdp_oracle_draw_f64(p->seed, HDP_STREAM_SAMPLER, o, (uint64_t)i, p->sampler_normal);
}

/* xoshiro256++ + 128-layer ziggurat (Doornik's ZIGNOR): a fast host
 * generator of the same class as Julia's default (Xoshiro256++ with a
 * ziggurat randn), so the CPU baseline is not handicapped by Philox +
 * Box–Muller. */
typedef struct xo_state {
    uint64_t s[4];
    int sampler_normal, is_f32;
} xo_state;
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t xo_next(xo_state* g)
{
    uint64_t* s = g->s;
    uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
    uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return result;
}
static uint64_t splitmix64(uint64_t* x)
{
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
#define ZIG_N 128
static double zig_x[ZIG_N + 1], zig_r[ZIG_N];
static pthread_once_t zig_once = PTHREAD_ONCE_INIT;
static void zig_init(void)
{
    const double R = 3.442619855899, V = 9.91256303526217e-3;
    double f = exp(-0.5 * R * R);
    zig_x[0] = V / f;
    zig_x[1] = R;
    zig_x[ZIG_N] = 0;
    for (int i = 2; i < ZIG_N; ++i) {
        zig_x[i] = sqrt(-2.0 * log(V / zig_x[i - 1] + f));
        f = exp(-0.5 * zig_x[i] * zig_x[i]);
    }
    for (int i = 0; i < ZIG_N; ++i) zig_r[i] = zig_x[i + 1] / zig_x[i];
}
static inline double xo_u01(xo_state* g) { return (double)(xo_next(g) >> 11) * 0x1.0p-53; }
static double xo_randn(xo_state* g)
{
    const double R = 3.442619855899;
    for (;;) {
        uint64_t bits = xo_next(g);
        int i = (int)(bits & (ZIG_N - 1));
        double u = 2.0 * ((double)(bits >> 11) * 0x1.0p-53) - 1.0;
        if (fabs(u) < zig_r[i]) return u * zig_x[i];
        if (i == 0) { /* tail */
            double x, y;
            do {
                x = log(1.0 - xo_u01(g)) / R;
                y = log(1.0 - xo_u01(g));
            } while (-2.0 * y < x * x);
            return u < 0 ? x - R : R - x;
        }
        double x = u * zig_x[i];
        double f0 = exp(-0.5 * (zig_x[i] * zig_x[i] - x * x));
        double f1 = exp(-0.5 * (zig_x[i + 1] * zig_x[i + 1] - x * x));
        if (f1 + xo_u01(g) * (f0 - f1) < 1.0) return x;
    }
}
static double xo_uniform(void* c)
{
    xo_state* g = (xo_state*)c;
    if (g->is_f32) return (double)((float)(xo_next(g) >> 40) * 0x1.0p-24f);
    return xo_u01(g);
}
static void xo_normals(void* c, double* out, int d)
{
    xo_state* g = (xo_state*)c;
    for (int i = 0; i < d; ++i) out[i] = xo_randn(g);
}
static void xo_sampler(void* c, double* out, int d)
{
    xo_state* g = (xo_state*)c;
    if (g->sampler_normal) for (int i = 0; i < d; ++i) out[i] = xo_randn(g);
    else for (int i = 0; i < d; ++i) out[i] = xo_uniform(c);
}

hdp_oracle_rng* hdp_oracle_rng_tape(int is_f32, const void* r, uint64_t rn, const void* n, uint64_t nn,
                                    const void* s, uint64_t sn)
{
    hdp_oracle_rng* g = (hdp_oracle_rng*)calloc(1, sizeof(*g));
    tape_state* t = (tape_state*)calloc(1, sizeof(*t));
    t->is_f32 = is_f32;
    t->r = r; t->n = n; t->s = s;
    t->rn = rn; t->nn = nn; t->sn = sn;
    g->ctx = t;
    g->uniform = tape_uniform;
    g->normals = tape_normals;
    g->sampler = tape_sampler;
    g->kind = 1;
    return g;
}
/* consumed counts / overrun flag of a tape rng: out[0..2] = r, n, s positions, out[3] = overrun */
void hdp_oracle_rng_tape_pos(hdp_oracle_rng* g, uint64_t out[4])
{
    tape_state* t = (tape_state*)g->ctx;
    out[0] = t->ri; out[1] = t->ni; out[2] = t->si; out[3] = (uint64_t)t->overrun;
}
hdp_oracle_rng* hdp_oracle_rng_philox(uint64_t seed, int is_f32, int sampler_normal)
{
    hdp_oracle_rng* g = (hdp_oracle_rng*)calloc(1, sizeof(*g));
    philox_state* p = (philox_state*)calloc(1, sizeof(*p));
    p->seed = seed; p->is_f32 = is_f32; p->sampler_normal = sampler_normal;
    g->ctx = p;
    g->uniform = philox_uniform;
    g->normals = philox_normals;
    g->sampler = philox_sampler;
    g->kind = 2;
    return g;
}
hdp_oracle_rng* hdp_oracle_rng_xoshiro(uint64_t seed, int is_f32, int sampler_normal)
{
    pthread_once(&zig_once, zig_init);
    hdp_oracle_rng* g = (hdp_oracle_rng*)calloc(1, sizeof(*g));
    xo_state* x = (xo_state*)calloc(1, sizeof(*x));
    uint64_t sm = seed;
    for (int i = 0; i < 4; ++i) x->s[i] = splitmix64(&sm);
    x->is_f32 = is_f32; x->sampler_normal = sampler_normal;
    g->ctx = x;
    g->uniform = xo_uniform;
    g->normals = xo_normals;
    g->sampler = xo_sampler;
    g->kind = 3;
    return g;
}
void hdp_oracle_rng_free(hdp_oracle_rng* g)
{
    if (!g) return;
    free(g->ctx);
    free(g);
}

/* fill n draws from an rng (testing the generators themselves) */
void hdp_oracle_rng_fill(hdp_oracle_rng* g, int what /*0 uniform, 1 normals, 2 sampler*/, double* out, int n)
{
    if (what == 0) for (int i = 0; i < n; ++i) out[i] = g->uniform(g->ctx);
    else if (what == 1) g->normals(g->ctx, out, n);
    else g->sampler(g->ctx, out, n);
}

/* ================================================================ */
/* work-count recurrences, for cross-checks     */
/* ================================================================ */
void hdp_oracle_mlp_counts(int M, int L, int K, int nonlocal, int64_t out[5])
{
    /* C: path-steps, R: time uniforms, S: sampler calls, F: f evals, G: g evals */
    int64_t C[32] = {0}, R[32] = {0}, S[32] = {0}, F[32] = {0}, G[32] = {0};
    int64_t Kp = nonlocal ? K : 0;
    for (int lv = 1; lv <= L && lv < 32; ++lv) {
        int64_t c = 0, r = 0, s = 0, f = 0, g = 0;
        for (int l = 0; l < lv; ++l) {
            int64_t num = 1;
            for (int i = 0; i < lv - l; ++i) num *= M;
            int64_t cc = C[l] + (l >= 2 ? C[l - 1] : 0);
            int64_t rr = R[l] + (l >= 2 ? R[l - 1] : 0);
            int64_t ss = S[l] + (l >= 2 ? S[l - 1] : 0);
            int64_t ff = F[l] + (l >= 2 ? F[l - 1] : 0);
            int64_t gg = G[l] + (l >= 2 ? G[l - 1] : 0);
            c += num * (1 + (1 + Kp) * cc);
            r += num * (1 + (1 + Kp) * rr);
            s += num * (Kp + (1 + Kp) * ss);
            f += num * ((int64_t)K * (l >= 2 ? 2 : 1) + (1 + Kp) * ff);
            g += num * ((1 + Kp) * gg);
        }
        int64_t numg = 1;
        for (int i = 0; i < lv; ++i) numg *= M;
        C[lv] = c + numg; R[lv] = r; S[lv] = s; F[lv] = f; G[lv] = g + numg;
    }
    out[0] = C[L]; out[1] = R[L]; out[2] = S[L]; out[3] = F[L]; out[4] = G[L];
}

hdp_oracle_user_fns hdp_oracle_user = {0, 0, 0, 0};
void hdp_oracle_set_user_fns(void* mu, void* sigma, void* g, void* f)
{
    hdp_oracle_user.mu = (hdp_oracle_user_coef)mu;
    hdp_oracle_user.sigma = (hdp_oracle_user_coef)sigma;
    hdp_oracle_user.g = (hdp_oracle_user_g)g;
    hdp_oracle_user.f = (hdp_oracle_user_f)f;
}

/* ================================================================ */
/* the two instantiations of the restatement                           */
/* ================================================================ */
#define REAL double
#define SUFFIX _f64
#include "hdp_oracle_impl.h"
#undef REAL
#undef SUFFIX

#define REAL float
#define SUFFIX _f32
#include "hdp_oracle_impl.h"
#undef REAL
#undef SUFFIX

/* ================================================================ */
/* CPU baseline helper (bench.py `cpu_baseline` / `--impl reference`):  */
/* n_solves independent serial `_ml_picard` solves spread over          */
/* `nthreads` host threads, each with its own xoshiro256++ stream.      */
/* This uses every host core (the reference's own threaded variant,     */
/* `_ml_picard_mlt`, is capped by its 4 heavy top-level samples).       */
/* ================================================================ */
#include <time.h>
typedef struct bench_job {
    const hdp_problem* pb;
    const hdp_sampler* sm;
    int M, L, K, n_solves, nthreads, tid;
    const double* x;
    uint64_t seed;
    double* u_out;
    int rc;
    int* next; /* shared: the next solve nobody has taken yet (dynamic schedule: no thread waits on a static share) */
} bench_job;

static void* bench_main(void* arg)
{
    bench_job* j = (bench_job*)arg;
    for (;;) {
        const int i = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n_solves) break;
        hdp_oracle_rng* g = hdp_oracle_rng_xoshiro(j->seed + (uint64_t)i, 0,
                                                   j->sm && j->sm->kind == HDP_SAMPLER_NORMAL);
        double u = 0;
        int rc = hdp_oracle_ml_picard_f64(j->pb, j->sm, j->M, j->L, j->K, j->x, g, &u, NULL);
        if (rc) j->rc = rc;
        j->u_out[i] = u;
        hdp_oracle_rng_free(g);
    }
    return NULL;
}

int hdp_oracle_bench_solves_f64(const hdp_problem* pb, const hdp_sampler* sm, int M, int L, int K,
                                const double* x, int n_solves, int nthreads, uint64_t seed,
                                double* u_out, double* seconds)
{
    if (nthreads < 1) nthreads = 1;
    bench_job* jobs = (bench_job*)calloc((size_t)nthreads, sizeof(bench_job));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    struct timespec t0, t1;
    int next = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = (bench_job){pb, sm, M, L, K, n_solves, nthreads, t, x, seed, u_out, 0, &next};
        pthread_create(&th[t], NULL, bench_main, &jobs[t]);
    }
    int rc = 0;
    for (int t = 0; t < nthreads; ++t) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    free(jobs);
    free(th);
    return rc;
}
