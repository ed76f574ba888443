/*
 * hdp_oracle_impl.h — type-generic body of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Included twice by hdp_oracle.c with REAL = double (suffix _f64) and
 * REAL = float (suffix _f32).  A literal, operation-for-operation C
 * restatement of the HighDimPDE.jl v2.6.0 Monte-Carlo hot path; every
 * function cites the reference lines it follows.  Compile with
 * -ffp-contract=off: Julia never fuses a*b+c.
 *
 * PARITY STATUS: the reference is pure Julia and no Julia toolchain exists
 * in this image, so this restatement cannot be run against the reference
 * itself.  It is pinned by (1) the closed-form answers of the reference's
 * own tests (test/MLP.jl:27,117,244-258, literature values :156,:224),
 * (2) the predicates of test/reflect.jl and test/MCSample.jl, and (3) a
 * cross-check between the two independent reflection algorithms.  At the
 * RNG boundary (Julia's Xoshiro/ziggurat bit-stream) parity is UNPINNED:
 * the reference's tests pin no bit-stream either.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* ---------------------------------------------------------------- */
/* user-function kinds (prob.μ, prob.σ, prob.g, prob.f)              */
/* ---------------------------------------------------------------- */

/* μ(x,p,t)[i] or σ(x,p,t)[i]; scalar-valued kinds broadcast. */
static inline REAL FN(coef)(int kind, const double* p, REAL xi, int i, REAL t, int d, int which)
{
    if (kind == HDP_COEF_LINEAR) return (REAL)p[0] * xi; /* test/MLP.jl:191-192: 0.02f0 * x */
    if (kind == HDP_COEF_QUADRATIC) return (REAL)p[0] * (xi * xi); /* test/DeepSplitting.jl:355: sigma * x .^ 2 */
    if (kind == HDP_COEF_USER) {                        /* user μ(x,p,t)[i] / σ(x,p,t)[i] */
        hdp_oracle_user_coef fn = which ? hdp_oracle_user.sigma : hdp_oracle_user.mu;
        return (REAL)fn((double)xi, i, (double)t, d, p);
    }
    return (REAL)p[0];                                  /* test/MLP.jl:24-25 */
}

static REAL FN(sumsq)(const REAL* x, int d)
{
    REAL s = 0;
    for (int i = 0; i < d; ++i) s += x[i] * x[i];
    return s;
}

/* prob.g(x) */
static REAL FN(g_eval)(const hdp_problem* pb, const REAL* x)
{
    const int d = pb->d;
    const double* p = pb->g_p;
    switch (pb->g_kind) {
    case HDP_G_SUMSQ: /* sum(x .^ 2) .+ 2.0f0 (test/MLP.jl:28,118) */
        return FN(sumsq)(x, d) + (REAL)p[0];
    case HDP_G_INV_SUMSQ: /* 1.0f0 ./ (2.0f0 .+ 4.0f-1 * sum(X .^ 2)) (test/MLP.jl:146) */
        return (REAL)1 / ((REAL)p[0] + (REAL)p[1] * FN(sumsq)(x, d));
    case HDP_G_EXP_SUMSQ: /* exp.(-0.25f0 * sum(X .^ 2)) (test/MLP.jl:174) */
        return (REAL)exp((double)(-(REAL)p[0] * FN(sumsq)(x, d)));
    case HDP_G_MIN: { /* minimum(X) (test/MLP.jl:199) */
        REAL m = x[0];
        for (int i = 1; i < d; ++i) m = x[i] < m ? x[i] : m;
        return m;
    }
    case HDP_G_GAUSS: { /* (2π)^(-d/2) * ss0^(-d*0.5) * exp(-0.5*sum(x .^ 2 / ss0)) (test/MLP.jl:265) */
        double ss0 = p[0];
        double s = 0;
        for (int i = 0; i < d; ++i) s += ((double)x[i] * (double)x[i]) / ss0;
        return (REAL)(pow(2.0 * M_PI, -d / 2.0) * pow(ss0, -d * 0.5) * exp(-0.5 * s));
    }
    case HDP_G_USER: {
        REAL m = x[0];
        for (int i = 1; i < d; ++i) m = x[i] < m ? x[i] : m;
        return (REAL)hdp_oracle_user.g((double)FN(sumsq)(x, d), (double)m, d, p);
    }
    }
    return 0;
}

static inline REAL FN(maxz)(REAL u) { return u > 0 ? u : (REAL)0; }
static inline REAL FN(ac)(REAL u) { return u - u * u * u; } /* a(u) = u - u^3 */

/* prob.f in the 8-argument form f(y, z, v_y, v_z, ∇v_y, ∇v_z, p, t)
 * (src/MLP.jl:58-62 wraps the 5-argument ParabolicPDEProblem form).
 * y, z enter only through sum(y.^2), sum(z.^2) in the supported kinds. */
static REAL FN(f_eval)(const hdp_problem* pb, REAL ssq_y, REAL ssq_z, REAL vy, REAL vz)
{
    const double* p = pb->f_p;
    switch (pb->f_kind) {
    case HDP_F_ZERO: return (REAL)0 * vy;          /* 0.0e0 .* v_y (test/MLP.jl:35) */
    case HDP_F_LINEAR: return (REAL)p[0] * vy;     /* r * v_y (test/MLP.jl:122) */
    case HDP_F_ALLEN_CAHN: return (REAL)p[0] * FN(ac)(vy); /* ±a.(v_y) (test/MLP.jl:148,176) */
    case HDP_F_DEFAULT_RISK: { /* test/MLP.jl:209-215 */
        REAL delta = (REAL)p[0], R = (REAL)p[1], vh = (REAL)p[2], vl = (REAL)p[3];
        REAL gh = (REAL)p[4], gl = (REAL)p[5];
        REAL u = vy;
        REAL Q = (REAL)(u < vh) * gh + (REAL)(u >= vl) * gl +
                 (REAL)((u >= vh) && (u < vl)) * (((gh - gl) / (vh - vl)) * (u - vh) + gh);
        return -((REAL)1 - delta) * Q * vy - R * vy;
    }
    case HDP_F_FISHER_KPP: /* max(0.0, v_x) * (1 - max(0.0, v_x)) (docs/src/tutorials/mlp.md:20) */
        return FN(maxz)(vy) * ((REAL)1 - FN(maxz)(vy));
    case HDP_F_NL_ALLEN_CAHN_DIFF: /* a.(v_y) .- a.(v_z) (test/MLP.jl:304) */
        return FN(ac)(vy) - FN(ac)(vz);
    case HDP_F_NL_FISHER: /* max(0.0, v_x) * (1 - max(0.0, v_y)) (docs/src/tutorials/mlp.md:52) */
        return FN(maxz)(vy) * ((REAL)1 - FN(maxz)(vz));
    case HDP_F_NL_REPLICATOR: { /* test/MLP.jl:266-274 */
        double ss = p[0];
        int d = pb->d;
        double my = -0.5 * (double)ssq_y, mz = -0.5 * (double)ssq_z;
        double w = pow(2.0 * M_PI, d / 2.0) * pow(ss, (double)d) * exp(0.5 * (double)ssq_z / (ss * ss));
        double r = (double)FN(maxz)(vy) * (my - (double)FN(maxz)(vz) * mz * w);
        return (REAL)r;
    }
    case HDP_F_USER:
    case HDP_F_NL_USER: /* f(..., p, t) is called with the outer end time (src/MLP.jl:140-146) */
        return (REAL)hdp_oracle_user.f((double)vy, (double)vz, (double)ssq_y, (double)ssq_z, pb->t1, pb->d, p);
    }
    return 0;
}

/* ---------------------------------------------------------------- */
/* `_reflect(a::T,b::T,s::T,e::T) where T<:Vector`  src/reflect.jl:8-55 */
/* Sequential billiard.  n = zeros(size(a)) is Float64 whatever T is    */
/* (:10), so for Float32 inputs b becomes Float64 after the first       */
/* bounce (:34) and a after the second (:32-33).  We track that with a  */
/* double working copy plus the `stage` of promotion.                   */
/* bounces[i] (nullable) counts wall hits of coordinate i.              */
/* returns 0, or HDP_ERR_START_OUTSIDE_DOMAIN (:12 `error(...)`).       */
/* ---------------------------------------------------------------- */
int FN(hdp_oracle_reflect_vector)(const REAL* a_in, const REAL* b_in, const REAL* s, const REAL* e,
                                  int d, REAL* out, int32_t* bounces, int64_t* total_bounces)
{
    /* :12 all((a .>= s) .& (a .<= e)) ? nothing : error(...) */
    for (int i = 0; i < d; ++i)
        if (!(a_in[i] >= s[i] && a_in[i] <= e[i])) return HDP_ERR_START_OUTSIDE_DOMAIN;
    if (bounces) memset(bounces, 0, sizeof(int32_t) * (size_t)d);

    /* working copies; WREAL is the type the reference holds at each stage */
    double* a = (double*)malloc(sizeof(double) * (size_t)d * 2);
    double* b = a + d;
    for (int i = 0; i < d; ++i) { a[i] = a_in[i]; b[i] = b_in[i]; }
    /* stage 0: a,b are REAL.  stage 1: b is Float64, a REAL.  stage 2: both Float64.
     * For REAL=double all stages coincide. */
    int stage = 0;
    int64_t nb = 0;

    for (;;) {
        /* scan :14-30 / :37-53 — strict '<' keeps the first index on ties */
        double r = 2; /* r = 2 (Int) compared/promoted */
        int hit = -1, sign = 0;
        for (int i = 0; i < d; ++i) {
            if (b[i] < (double)s[i]) {
                double rtemp;
                if (stage == 0) rtemp = (double)(((REAL)a[i] - s[i]) / ((REAL)a[i] - (REAL)b[i]));
                else if (stage == 1) rtemp = (double)((REAL)a[i] - s[i]) / ((double)(REAL)a[i] - b[i]);
                else rtemp = (a[i] - (double)s[i]) / (a[i] - b[i]);
                if (rtemp < r) { r = rtemp; hit = i; sign = -1; }
            } else if (b[i] > (double)e[i]) {
                double rtemp;
                if (stage == 0) rtemp = (double)((e[i] - (REAL)a[i]) / ((REAL)b[i] - (REAL)a[i]));
                else if (stage == 1) rtemp = (double)(e[i] - (REAL)a[i]) / (b[i] - (double)(REAL)a[i]);
                else rtemp = ((double)e[i] - a[i]) / (b[i] - a[i]);
                if (rtemp < r) { r = rtemp; hit = i; sign = 1; }
            }
        }
        if (!(r < 1)) break; /* while r < 1  (:31) */
        /* c = a + r * (b - a)   (:32) */
        /* a = c                 (:33) */
        /* b = b - 2 * n * dot(b - c, n)  (:34): only coordinate `hit` changes */
        for (int j = 0; j < d; ++j) {
            double c;
            if (stage == 0) c = (double)((REAL)a[j] + (REAL)r * ((REAL)b[j] - (REAL)a[j]));
            else c = a[j] + r * (b[j] - a[j]); /* Float64 from stage 1 on (b or r is Float64) */
            if (j == hit) {
                /* dot(b-c,n) is Float64 because n is; 2*n*dot = 2*(b-c) exactly signed */
                double bc = b[j] - c;
                if (stage == 0) bc = (double)((REAL)b[j] - (REAL)c);
                b[j] = b[j] - 2.0 * bc;
            }
            a[j] = c;
        }
        (void)sign;
        if (bounces) bounces[hit]++;
        nb++;
        if (sizeof(REAL) == sizeof(double)) stage = 0; /* no promotion for Float64 */
        else stage = stage < 2 ? stage + 1 : 2;
        if (nb > 100000000) break; /* defensive: NaN-free inputs always terminate */
    }
    for (int i = 0; i < d; ++i) out[i] = (REAL)b[i]; /* x2 .= _reflect(...) converts back (src/MLP.jl:419) */
    if (total_bounces) *total_bounces += nb;
    free(a);
    return HDP_OK;
}

/* ---------------------------------------------------------------- */
/* `_reflect(a::T,b::T,s,e) where T<:AbstractArray` src/reflect.jl:58-82 */
/* Batched columnwise version; loops while ANY element of the whole     */
/* array is outside (:67).  s,e scalar (stride 0) or length-d.          */
/* iters_out (nullable): number of global while-iterations.             */
/* col_iters (nullable, length B): iterations in which column j still   */
/* had an element outside.  bounces (nullable d×B): per-coordinate wall */
/* hits (n == 1 while that element was outside).                        */
/* ---------------------------------------------------------------- */
int FN(hdp_oracle_reflect_array)(const REAL* a_in, const REAL* b_in, const REAL* s, const REAL* e,
                                 int bounds_are_scalar, int d, int B, REAL* out,
                                 int32_t* bounces, int32_t* col_iters, int32_t* iters_out)
{
    const size_t n_el = (size_t)d * (size_t)B;
    const int st = bounds_are_scalar ? 0 : 1;
    /* :59 @assert all((a .>= s) .& (a .<= e)) */
    for (size_t q = 0; q < n_el; ++q) {
        int i = (int)(q % (size_t)d);
        if (!(a_in[q] >= s[i * st] && a_in[q] <= e[i * st])) return HDP_ERR_START_OUTSIDE_DOMAIN;
    }
    REAL* a = (REAL*)malloc(sizeof(REAL) * n_el * 3);
    REAL* b = a + n_el;
    REAL* rtemp = b + n_el;
    unsigned char* out1 = (unsigned char*)malloc(n_el * 2);
    unsigned char* out2 = out1 + n_el;
    memcpy(a, a_in, sizeof(REAL) * n_el);
    memcpy(b, b_in, sizeof(REAL) * n_el);
    if (bounces) memset(bounces, 0, sizeof(int32_t) * n_el);
    if (col_iters) memset(col_iters, 0, sizeof(int32_t) * (size_t)B);
    int any = 0;
    for (size_t q = 0; q < n_el; ++q) { /* :61-63 */
        int i = (int)(q % (size_t)d);
        out1[q] = b[q] < s[i * st];
        out2[q] = b[q] > e[i * st];
        any |= out1[q] | out2[q];
    }
    int iters = 0;
    while (any) { /* :67 while any(out) */
        for (int j = 0; j < B; ++j) {
            REAL* aj = a + (size_t)j * d;
            REAL* bj = b + (size_t)j * d;
            REAL* rj = rtemp + (size_t)j * d;
            unsigned char* o1 = out1 + (size_t)j * d;
            unsigned char* o2 = out2 + (size_t)j * d;
            REAL rmin = 0;
            int col_out = 0;
            for (int i = 0; i < d; ++i) {
                /* :68-71 with Bool-multiply semantics reduced (false*x = ±0) */
                int o = o1[i] | o2[i];
                REAL div = o ? (bj[i] - aj[i]) : (REAL)1;
                REAL num = o1[i] ? (s[i * st] - aj[i]) : (o2[i] ? (e[i * st] - aj[i]) : (REAL)0);
                REAL rt = num / div + (o ? (REAL)0 : (REAL)1);
                rj[i] = rt;
                rmin = (i == 0 || rt < rmin) ? rt : rmin; /* :72 minimum(rtemp, dims=1) */
                col_out |= o;
            }
            if (col_out && col_iters) col_iters[j]++;
            for (int i = 0; i < d; ++i) {
                REAL nn = (rj[i] == rmin) ? (REAL)1 : (REAL)0;   /* :73 */
                REAL c = aj[i] + (bj[i] - aj[i]) * rmin;          /* :74 */
                if (bounces && nn != 0 && (o1[i] | o2[i])) bounces[(size_t)j * d + i]++;
                bj[i] = bj[i] - ((REAL)2 * nn) * (bj[i] - c);      /* :75 */
                aj[i] = c;                                          /* :76 */
            }
        }
        any = 0;
        for (size_t q = 0; q < n_el; ++q) { /* :77-79 */
            int i = (int)(q % (size_t)d);
            out1[q] = b[q] < s[i * st];
            out2[q] = b[q] > e[i * st];
            any |= out1[q] | out2[q];
        }
        iters++;
        if (iters > 10000000) break;
    }
    memcpy(out, b, sizeof(REAL) * n_el);
    if (iters_out) *iters_out = iters;
    free(a);
    free(out1);
    return HDP_OK;
}

/* ---------------------------------------------------------------- */
/* samplers — src/MCSample.jl                                        */
/* `raw` holds the rand!/randn! output; x_mc receives the result.    */
/* ---------------------------------------------------------------- */
/* (mc_sample::UniformSampling)(x_mc, kwargs...)  src/MCSample.jl:28-33 */
static void FN(uniform_apply)(const hdp_sampler* sm, const REAL* raw, REAL* x_mc, int d, size_t ncol)
{
    for (size_t j = 0; j < ncol; ++j)
        for (int i = 0; i < d; ++i) {
            REAL a = (REAL)(sm->bounds_are_scalar ? sm->a_scalar : sm->a[i]);
            REAL b = (REAL)(sm->bounds_are_scalar ? sm->b_scalar : sm->b[i]);
            REAL m = (b + a) / (REAL)2;                       /* :31 */
            size_t q = j * (size_t)d + (size_t)i;
            x_mc[q] = (raw[q] - (REAL)0.5) * (b - a) + m;     /* :32 */
        }
}
/* (mc_sample::NormalSampling)(x_mc) and (x_mc, x)  src/MCSample.jl:63-71.
 * xref: d×B reference points broadcast along the trailing axis (K), or NULL
 * for the 1-argument form. */
static void FN(normal_apply)(const hdp_sampler* sm, const REAL* raw, REAL* x_mc, const REAL* xref,
                             int d, size_t B, size_t K)
{
    const REAL sg = (REAL)sm->sigma;
    for (size_t k = 0; k < K; ++k)
        for (size_t q = 0; q < (size_t)d * B; ++q) {
            REAL v = raw[k * d * B + q] * sg;                 /* :65 x_mc .*= σ */
            if (xref && sm->shifted) v = v + xref[q];         /* :70 x_mc .+= x */
            x_mc[k * d * B + q] = v;
        }
}

/* public: apply a sampler to raw draws.  z: d×B×K, y1: d×B or NULL. */
int FN(hdp_oracle_mc_sample)(const hdp_sampler* sm, const REAL* raw, REAL* z, const REAL* y1,
                             int d, int B, int K)
{
    if (sm->kind == HDP_SAMPLER_UNIFORM) FN(uniform_apply)(sm, raw, z, d, (size_t)B * (size_t)K);
    else if (sm->kind == HDP_SAMPLER_NORMAL) FN(normal_apply)(sm, raw, z, y1, d, (size_t)B, (size_t)K);
    else return HDP_ERR_BAD_ARG; /* NoSampling()(x...) = nothing (:86) */
    return HDP_OK;
}

/* ---------------------------------------------------------------- */
/* MLP                                                               */
/* ---------------------------------------------------------------- */
typedef struct FN(mlp_ctx) {
    const hdp_problem* pb;
    const hdp_sampler* sm;
    int M, K;
    int integrate;           /* _integrate(mc_sample!)  src/MCSample.jl:88-94 */
    REAL t;                  /* outer end time, passed unchanged down the recursion */
    REAL* lo;                /* neumann_bc[1], neumann_bc[2] as length-d vectors */
    REAL* hi;
    hdp_oracle_rng* rng;
    int64_t bounces;
    int64_t n_steps, n_unif, n_samp, n_f, n_g;
    int err;
    REAL* tmp;               /* d scratch (reflect output / raw draws in REAL) */
    double* tmp_d;           /* d scratch for raw draws from the rng callbacks */
} FN(mlp_ctx);

/* `_mlt_sde_loop!(x2, x, s, t, prob, neumann_bc)`  src/MLP.jl:406-421 */
static void FN(mlt_sde_loop)(FN(mlp_ctx)* c, REAL* x2, const REAL* x, REAL s, REAL t)
{
    const hdp_problem* pb = c->pb;
    const int d = pb->d;
    REAL dt = t - s;                                           /* :415 */
    REAL sq = (REAL)sqrt((double)dt);                          /* sqrt(dt) in tType */
    if (sizeof(REAL) == sizeof(float)) sq = sqrtf((float)dt);
    c->rng->normals(c->rng->ctx, c->tmp_d, d);                 /* :416 randn!(x2) */
    for (int i = 0; i < d; ++i) {
        REAL xi = (REAL)c->tmp_d[i];
        REAL mu = FN(coef)(pb->mu_kind, pb->mu_p, x[i], i, t, d, 0); /* μ(x, p, t): start point, END time */
        REAL sg = FN(coef)(pb->sigma_kind, pb->sigma_p, x[i], i, t, d, 1);
        /* :417 x2 .= x + (μ .* dt .+ σ .* sqrt(dt) .* x2) */
        x2[i] = x[i] + ((mu * dt) + ((sg * sq) * xi));
    }
    c->n_steps++;
    if (pb->has_bc) {                                          /* :418-420 */
        int rc = FN(hdp_oracle_reflect_vector)(x, x2, c->lo, c->hi, d, c->tmp, NULL, &c->bounces);
        if (rc != HDP_OK) { c->err = rc; return; }
        memcpy(x2, c->tmp, sizeof(REAL) * (size_t)d);
    }
}

/* mc_sample!(x3, x2) inside the recursion  src/MLP.jl:139,177 */
static void FN(mlp_sample)(FN(mlp_ctx)* c, REAL* x3, const REAL* x2)
{
    const int d = c->pb->d;
    c->rng->sampler(c->rng->ctx, c->tmp_d, d); /* rand!/randn!  src/MCSample.jl:30,64 */
    for (int i = 0; i < d; ++i) c->tmp[i] = (REAL)c->tmp_d[i];
    if (c->sm->kind == HDP_SAMPLER_UNIFORM) FN(uniform_apply)(c->sm, c->tmp, x3, d, 1);
    else FN(normal_apply)(c->sm, c->tmp, x3, x2, d, 1, 1);
    c->n_samp++;
}

/* public single step for unit tests: normals = ξ (d values), x2_out = result */
int FN(hdp_oracle_mlt_sde_loop)(const hdp_problem* pb, const REAL* x, REAL s, REAL t, const REAL* normals, REAL* x2_out);

static int64_t FN(ipow)(int M, int e)
{
    int64_t r = 1;
    for (int i = 0; i < e; ++i) r *= M;
    return r;
}

static REAL FN(ml_picard)(FN(mlp_ctx)* c, int L, const REAL* x, REAL s);

/* body of one level-`l` sample (the HOT LOOP): src/MLP.jl:132-148 (l<=1)
 * and :157-194 (l>=2).  Returns b3 / K. */
static REAL FN(level_sample)(FN(mlp_ctx)* c, int l, const REAL* x, REAL s, REAL* x2, REAL* x3)
{
    const REAL t = c->t;
    const int K = c->K;
    REAL u = (REAL)c->rng->uniform(c->rng->ctx);
    c->n_unif++;
    REAL r = s + (t - s) * u;                                   /* :132 */
    FN(mlt_sde_loop)(c, x2, x, s, r);                           /* :133 */
    if (c->err) return 0;
    REAL b2 = FN(ml_picard)(c, l, x2, r);                       /* :134 */
    REAL b4 = 0;
    if (l >= 2) b4 = FN(ml_picard)(c, l - 1, x2, r);            /* :160-173 */
    REAL ssq2 = FN(sumsq)(x2, c->pb->d);
    REAL b3 = 0;
    for (int h = 0; h < K; ++h) {                               /* :137, :176 */
        REAL ssq3 = 0, v3 = 0, w3 = 0;
        if (c->integrate) {
            FN(mlp_sample)(c, x3, x2);                          /* :139, :177 */
            ssq3 = FN(sumsq)(x3, c->pb->d);
            v3 = FN(ml_picard)(c, l, x3, r);
            if (c->err) return 0;
        } /* else: x3 === nothing and _ml_picard(…, x::Nothing, …) = nothing (:210-216) */
        REAL term = FN(f_eval)(c->pb, ssq2, ssq3, b2, v3);      /* :140-146 */
        c->n_f++;
        if (l >= 2) {
            if (c->integrate) {
                w3 = FN(ml_picard)(c, l - 1, x3, r);
                if (c->err) return 0;
            }
            term = term - FN(f_eval)(c->pb, ssq2, ssq3, b4, w3); /* :178-191 */
            c->n_f++;
        }
        b3 += term;
    }
    return b3 / (REAL)K;                                        /* :148 b += b3 / K */
}

/* `_ml_picard(M, L, K, x, s, t, mc_sample!, g, f, verbose, prob, neumann_bc)`
 * src/MLP.jl:102-208 */
static REAL FN(ml_picard)(FN(mlp_ctx)* c, int L, const REAL* x, REAL s)
{
    if (L == 0) return 0;                                       /* :117-119 */
    if (c->err) return 0;
    const int d = c->pb->d;
    const int M = c->M;
    const REAL t = c->t;
    REAL* x2 = (REAL*)malloc(sizeof(REAL) * (size_t)d * 2);     /* :121-122 similar(x) */
    REAL* x3 = x2 + d;
    REAL a = 0;
    /* the two loops :126-151 (l = 0..min(L-1,1)) and :153-196 (l = 2..L-1)
     * visit l = 0..L-1 in increasing order; `level_sample` switches on l>=2 */
    for (int l = 0; l <= L - 1; ++l) {
        REAL b = 0;
        int64_t num = FN(ipow)(M, L - l);
        for (int64_t k = 0; k < num; ++k) {
            b += FN(level_sample)(c, l, x, s, x2, x3);
            if (c->err) { free(x2); return 0; }
        }
        a += (t - s) * b / (REAL)num;                           /* :150, :195 */
    }
    int64_t num = FN(ipow)(M, L);                               /* :198 */
    REAL a2 = 0;
    for (int64_t k = 0; k < num; ++k) {
        FN(mlt_sde_loop)(c, x2, x, s, t);                       /* :202 */
        if (c->err) { free(x2); return 0; }
        a2 += FN(g_eval)(c->pb, x2);                            /* :203 */
        c->n_g++;
    }
    a2 /= (REAL)num;                                            /* :205 */
    free(x2);
    return a + a2;                                              /* :207 */
}

/* `_get_loop_num(M, num, thread_id, NUM_THREADS)`  src/MLP.jl:391-404.
 * loop_num is a Float64 in the reference (`num / NUM_THREADS`); `1:loop_num`
 * iterates floor(loop_num) times. */
int64_t FN(hdp_oracle_get_loop_num)(int M, int64_t num, int thread_id, int nthreads)
{
    if (num < nthreads) return thread_id > num ? 0 : 1;
    int64_t remainder = (int64_t)M % num;
    double ln = (double)num / (double)nthreads;
    if (remainder > 0 && thread_id <= remainder) ln += 1.0;
    return (int64_t)floor(ln);
}

/* `_ml_picard_call(..., NUM_THREADS, thread_id, ...)`  src/MLP.jl:262-388 */
static REAL FN(ml_picard_call)(FN(mlp_ctx)* c, int L, const REAL* x, REAL s, int nthreads, int thread_id)
{
    const int d = c->pb->d;
    const int M = c->M;
    const REAL t = c->t;
    REAL* x2 = (REAL*)malloc(sizeof(REAL) * (size_t)d * 2);
    REAL* x3 = x2 + d;
    REAL a = 0;
    for (int l = 0; l <= L - 1; ++l) {
        REAL b = 0;
        int64_t num = FN(ipow)(M, L - l);
        int64_t loop_num = FN(hdp_oracle_get_loop_num)(M, num, thread_id, nthreads); /* :283, :311 */
        for (int64_t k = 0; k < loop_num; ++k) {
            b += FN(level_sample)(c, l, x, s, x2, x3);
            if (c->err) { free(x2); return 0; }
        }
        a += (t - s) * b / (REAL)num;                           /* :307, :384 (global num) */
    }
    free(x2);
    return a;
}

typedef struct FN(mlt_task) {
    FN(mlp_ctx) c;
    int L, nthreads, thread_id;
    const REAL* x;
    REAL s;
    REAL result;
} FN(mlt_task);

static void* FN(mlt_thread_main)(void* arg)
{
    FN(mlt_task)* tk = (FN(mlt_task)*)arg;
    tk->result = FN(ml_picard_call)(&tk->c, tk->L, tk->x, tk->s, tk->nthreads, tk->thread_id);
    return NULL;
}

static int FN(ctx_init)(FN(mlp_ctx)* c, const hdp_problem* pb, const hdp_sampler* sm, int M, int K,
                        hdp_oracle_rng* rng)
{
    memset(c, 0, sizeof(*c));
    c->pb = pb;
    c->sm = sm;
    c->M = M;
    c->K = K;
    c->integrate = sm && sm->kind != HDP_SAMPLER_NONE;
    c->t = (REAL)pb->t1;
    c->rng = rng;
    const int d = pb->d;
    c->tmp = (REAL*)malloc(sizeof(REAL) * (size_t)d * 3);
    c->lo = c->tmp + d;
    c->hi = c->lo + d;
    c->tmp_d = (double*)malloc(sizeof(double) * (size_t)d);
    for (int i = 0; i < d; ++i) {
        c->lo[i] = (REAL)(pb->has_bc ? (pb->bc_is_scalar ? pb->lo_scalar : pb->lo[i]) : 0.0);
        c->hi[i] = (REAL)(pb->has_bc ? (pb->bc_is_scalar ? pb->hi_scalar : pb->hi[i]) : 0.0);
    }
    if (((pb->f_kind >= HDP_F_NL_FIRST && pb->f_kind < HDP_F_USER) || pb->f_kind == HDP_F_NL_USER) && !c->integrate)
        return HDP_ERR_BAD_ARG;
    return HDP_OK;
}
static void FN(ctx_free)(FN(mlp_ctx)* c)
{
    free(c->tmp);
    free(c->tmp_d);
}

typedef struct FN(one_step_rng) { const REAL* xi; } FN(one_step_rng);
static void FN(one_step_normals)(void* c, double* out, int d)
{
    const REAL* xi = ((FN(one_step_rng)*)c)->xi;
    for (int i = 0; i < d; ++i) out[i] = (double)xi[i];
}
int FN(hdp_oracle_mlt_sde_loop)(const hdp_problem* pb, const REAL* x, REAL s, REAL t, const REAL* normals, REAL* x2_out)
{
    FN(one_step_rng) st = {normals};
    hdp_oracle_rng rng = {&st, NULL, FN(one_step_normals), NULL, 0};
    FN(mlp_ctx) c;
    int rc = FN(ctx_init)(&c, pb, NULL, 1, 1, &rng);
    if (rc == HDP_OK) {
        FN(mlt_sde_loop)(&c, x2_out, x, s, t);
        rc = c.err;
    }
    FN(ctx_free)(&c);
    return rc;
}

/* solve(prob, alg::MLP; multithreading=false) → `_ml_picard` (src/MLP.jl:84-98).
 * Returns usol; *counts (nullable) receives consumption counters. */
int FN(hdp_oracle_ml_picard)(const hdp_problem* pb, const hdp_sampler* sm, int M, int L, int K,
                             const REAL* x, hdp_oracle_rng* rng, REAL* u_out,
                             int64_t* counts /* [6]: steps, unif, samp, f, g, bounces */)
{
    FN(mlp_ctx) c;
    int rc = FN(ctx_init)(&c, pb, sm, M, K, rng);
    if (rc != HDP_OK) { FN(ctx_free)(&c); return rc; }
    REAL u = FN(ml_picard)(&c, L, x, (REAL)pb->t0);
    if (counts) {
        counts[0] = c.n_steps; counts[1] = c.n_unif; counts[2] = c.n_samp;
        counts[3] = c.n_f; counts[4] = c.n_g; counts[5] = c.bounces;
    }
    rc = c.err;
    FN(ctx_free)(&c);
    *u_out = u;
    return rc;
}

/* solve(prob, alg::MLP; multithreading=true) → `_ml_picard_mlt` (src/MLP.jl:218-260)
 * with Threads.nthreads() == nthreads.  rngs[0] is the main task's stream (the
 * g-term, :246-254); rngs[tid] (tid = 1..nthreads) the spawned tasks'.
 * parallel != 0 runs the tasks on pthreads (timing); parallel == 0 runs them
 * in order tid = 1..NT and then the main task's g-term, so a single shared
 * sequential tape is consumed in a defined order (parity). */
int FN(hdp_oracle_ml_picard_mlt)(const hdp_problem* pb, const hdp_sampler* sm, int M, int L, int K,
                                 const REAL* x, int nthreads, hdp_oracle_rng** rngs, int parallel,
                                 REAL* u_out, int64_t* counts)
{
    const int d = pb->d;
    FN(mlt_task)* tasks = (FN(mlt_task)*)calloc((size_t)nthreads, sizeof(FN(mlt_task)));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    int rc = HDP_OK;
    for (int i = 0; i < nthreads; ++i) {
        int r2 = FN(ctx_init)(&tasks[i].c, pb, sm, M, K, rngs[i + 1]);
        if (r2 != HDP_OK) rc = r2;
        tasks[i].L = L; tasks[i].nthreads = nthreads; tasks[i].thread_id = i + 1;
        tasks[i].x = x; tasks[i].s = (REAL)pb->t0;
    }
    FN(mlp_ctx) mc;
    int r3 = FN(ctx_init)(&mc, pb, sm, M, K, rngs[0]);
    if (r3 != HDP_OK) rc = r3;
    REAL a2 = 0;
    if (rc == HDP_OK) {
        if (parallel)
            for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, FN(mlt_thread_main), &tasks[i]);
        else
            for (int i = 0; i < nthreads; ++i) FN(mlt_thread_main)(&tasks[i]);
        /* main task: the g-term, :246-254 */
        int64_t num = FN(ipow)(M, L);
        REAL* x2 = (REAL*)malloc(sizeof(REAL) * (size_t)d);
        for (int64_t k = 0; k < num; ++k) {
            FN(mlt_sde_loop)(&mc, x2, x, (REAL)pb->t0, mc.t);
            if (mc.err) break;
            a2 += FN(g_eval)(pb, x2);
            mc.n_g++;
        }
        a2 /= (REAL)num;
        free(x2);
        if (parallel)
            for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    }
    /* a = sum([fetch(t) for t in tasks]); return a + a2  (:257-259) */
    REAL a = 0;
    int64_t cn[6] = {mc.n_steps, mc.n_unif, mc.n_samp, mc.n_f, mc.n_g, mc.bounces};
    if (mc.err) rc = mc.err;
    for (int i = 0; i < nthreads; ++i) {
        a += tasks[i].result;
        cn[0] += tasks[i].c.n_steps; cn[1] += tasks[i].c.n_unif; cn[2] += tasks[i].c.n_samp;
        cn[3] += tasks[i].c.n_f; cn[4] += tasks[i].c.n_g; cn[5] += tasks[i].c.bounces;
        if (tasks[i].c.err) rc = tasks[i].c.err;
        FN(ctx_free)(&tasks[i].c);
    }
    FN(ctx_free)(&mc);
    if (counts) memcpy(counts, cn, sizeof(cn));
    *u_out = a + a2;
    free(tasks);
    free(th);
    return rc;
}

/* ---------------------------------------------------------------- */
/* `sde_loop!(y0, y1, dWall)` closure  src/DeepSplitting.jl:153-166    */
/* dWall: d×B×nsteps standard normals (after randn!, :154).            */
/* x0_raw: raw draws for x0_sample!(y1) (:155) or NULL (NoSampling).   */
/* ts[i]: time passed to μ,σ at step i (= Julia's ts[N+1-i], :157).    */
/* y1 holds x0_batch on entry (:182).                                  */
/* ---------------------------------------------------------------- */
int FN(hdp_oracle_sde_loop)(const hdp_problem* pb, REAL* y0, REAL* y1, int B, int nsteps, REAL dt,
                            const REAL* ts, const hdp_sampler* x0s, const REAL* x0_raw,
                            const REAL* dWall, int32_t* iters_per_step)
{
    const int d = pb->d;
    const size_t n_el = (size_t)d * (size_t)B;
    if (x0s && x0s->kind != HDP_SAMPLER_NONE) {
        if (!x0_raw) return HDP_ERR_BAD_ARG;
        /* x0_sample!(y1): 1-argument form (:155) — never shifted */
        if (x0s->kind == HDP_SAMPLER_UNIFORM) FN(uniform_apply)(x0s, x0_raw, y1, d, (size_t)B);
        else FN(normal_apply)(x0s, x0_raw, y1, NULL, d, (size_t)B, 1);
    }
    REAL sq = sizeof(REAL) == sizeof(float) ? (REAL)sqrtf((float)dt) : (REAL)sqrt((double)dt);
    REAL* lo = NULL;
    REAL* hi = NULL;
    REAL* tmp = NULL;
    if (pb->has_bc) {
        lo = (REAL*)malloc(sizeof(REAL) * (size_t)d * 2);
        hi = lo + d;
        tmp = (REAL*)malloc(sizeof(REAL) * n_el);
        for (int i = 0; i < d; ++i) {
            lo[i] = (REAL)(pb->bc_is_scalar ? pb->lo_scalar : pb->lo[i]);
            hi[i] = (REAL)(pb->bc_is_scalar ? pb->hi_scalar : pb->hi[i]);
        }
    }
    int rc = HDP_OK;
    for (int st = 0; st < nsteps; ++st) {                       /* :156 */
        const REAL* dW = dWall + (size_t)st * n_el;             /* :158 */
        memcpy(y0, y1, sizeof(REAL) * n_el);                    /* :159 y0 .= y1 */
        const REAL t = ts ? ts[st] : (REAL)0;                   /* :157 t = ts[N + 1 - i] */
        for (size_t q = 0; q < n_el; ++q) {
            REAL mu = FN(coef)(pb->mu_kind, pb->mu_p, y0[q], (int)(q % (size_t)d), t, d, 0);
            REAL sg = FN(coef)(pb->sigma_kind, pb->sigma_p, y0[q], (int)(q % (size_t)d), t, d, 1);
            /* :160 y1 .= y0 .+ μ .* dt .+ σ .* sqrt(dt) .* dW   →  (y0 + μ*dt) + ((σ*sqrt(dt))*dW) */
            y1[q] = (y0[q] + mu * dt) + ((sg * sq) * dW[q]);
        }
        if (pb->has_bc) {                                       /* :161-163 */
            int32_t it = 0;
            rc = FN(hdp_oracle_reflect_array)(y0, y1, lo, hi, 0, d, B, tmp, NULL, NULL, &it);
            if (rc != HDP_OK) break;
            memcpy(y1, tmp, sizeof(REAL) * n_el);
            if (iters_per_step) iters_per_step[st] = it;
        }
    }
    free(lo);
    free(tmp);
    return rc;
}

/* ---------------------------------------------------------------- */
/* Plain Euler–Maruyama ensemble: the pre-simulation stage of          */
/* NNKolmogorov (src/NNKolmogorov.jl:81-112) and NNStopping            */
/* (src/NNStopping.jl:90-99):                                          */
/*   solve(EnsembleProblem(SDEProblem(μ, σ, x0, tspan)), EM();         */
/*         dt, trajectories, adaptive = false [, saveat = dt])         */
/* The step is StochasticDiffEq.jl's `EM` (Project.toml compat "7";    */
/* third-party, not vendored under /root/reference, so its published   */
/* algorithm is restated): ΔW = sqrt(dt)·ξ (DiffEqNoiseProcess         */
/* WHITE_NOISE_DIST), u = uprev + dt·f(uprev,p,t) + g(uprev,p,t)·ΔW    */
/* for diagonal noise, evaluated left to right; t advances by dt.      */
/* x: d×B start → final states; traj (nullable): d×(nsteps+1)×B as     */
/* Array(sim); xi: d×B×nsteps standard normals.                        */
/* ---------------------------------------------------------------- */
int FN(hdp_oracle_em_ensemble)(const hdp_problem* pb, REAL* x, int B, int nsteps, REAL dt, REAL t0,
                               REAL* traj, const REAL* xi)
{
    const int d = pb->d;
    const size_t n_el = (size_t)d * (size_t)B;
    REAL sq = sizeof(REAL) == sizeof(float) ? (REAL)sqrtf((float)dt) : (REAL)sqrt((double)dt);
    for (int b = 0; b < B; ++b) {
        REAL* u = x + (size_t)b * d;
        REAL* tr = traj ? traj + (size_t)b * d * ((size_t)nsteps + 1) : NULL;
        if (tr) memcpy(tr, u, sizeof(REAL) * (size_t)d);
        REAL t = t0;
        for (int st = 0; st < nsteps; ++st) {
            const REAL* z = xi + (size_t)st * n_el + (size_t)b * d;
            for (int i = 0; i < d; ++i) {
                REAL dW = sq * z[i];
                REAL mu = FN(coef)(pb->mu_kind, pb->mu_p, u[i], i, t, d, 0);
                REAL sg = FN(coef)(pb->sigma_kind, pb->sigma_p, u[i], i, t, d, 1);
                u[i] = (u[i] + dt * mu) + sg * dW;
            }
            t = t + dt;
            if (tr) memcpy(tr + (size_t)(st + 1) * d, u, sizeof(REAL) * (size_t)d);
        }
    }
    return HDP_OK;
}

#undef FN
#undef CAT
#undef CAT_
