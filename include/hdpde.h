/*
 * hdpde.h — C ABI of libhdpde_b200.so
 *
 * B200-native (sm_100a) replacement for the Monte-Carlo hot path of
 * SciML/HighDimPDE.jl.  Julia reaches these entry points with `ccall`
 * (see INTEGRATION.md); Python reaches them with ctypes (the package
 * `highdimpde.jl_b200/`).  No torch types, no C++ types, no exceptions
 * cross this boundary: plain pointers, sizes, POD structs and int status.
 *
 * There is NO CPU fallback behind these calls: without a usable CUDA
 * device `hdp_create` returns HDP_ERR_NO_DEVICE and nothing else works.
 *
 * Every entry point cites the reference interface it replaces
 * (paths relative to the HighDimPDE.jl v2.6.0 tree).
 *
 * Array conventions: column-major exactly as Julia passes `pointer(A)`:
 * a d×B matrix stores the d coordinates of one path contiguously, a
 * d×B×N array stores step slices contiguously.
 *
 * Closures cannot cross a C ABI onto a GPU.  The user functions
 * mu, sigma, g, f of `PIDEProblem` / `ParabolicPDEProblem`
 * (src/HighDimPDE.jl:56-59) are therefore described by an enumerated
 * *kind* plus a parameter block.  The kinds cover every function used by
 * test/MLP.jl, test/DeepSplitting.jl and docs/src/tutorials/mlp.md.
 */
#ifndef HDPDE_H
#define HDPDE_H

#ifndef __CUDACC_RTC__
#include <stdint.h>
#include <stddef.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define HDP_VERSION 100 /* 0.1.0 */

/* ------------------------------------------------------------------ */
/* status codes (0 = OK, negative = error)                             */
/* ------------------------------------------------------------------ */
enum {
    HDP_OK = 0,
    HDP_ERR_BAD_ARG = -1,
    /* Vector `_reflect` errors when the start point is outside the box
     * (src/reflect.jl:12, :59) */
    HDP_ERR_START_OUTSIDE_DOMAIN = -2,
    /* solve(::MLP) errors unless prob.x0_sample isa NoSampling
     * (src/MLP.jl:65-66) */
    HDP_ERR_X0_SAMPLE_NOT_SUPPORTED = -3,
    HDP_ERR_CUDA = -4,
    HDP_ERR_NCCL = -5,
    HDP_ERR_NO_DEVICE = -6,
    HDP_ERR_ALLOC = -7,
    HDP_ERR_UNSUPPORTED = -8,
    HDP_ERR_NVRTC = -9
};

/* ------------------------------------------------------------------ */
/* memory flags: where the caller's array arguments live               */
/* ------------------------------------------------------------------ */
enum {
    HDP_MEM_HOST = 0,   /* plain host pointers (Julia Array)            */
    HDP_MEM_DEVICE = 1, /* device pointers (Julia CuArray / CuPtr)      */
    HDP_ASYNC = 2       /* with HDP_MEM_DEVICE: enqueue only, do not wait;
                           a start-outside-domain error is then reported by the
                           next hdp_sync (overlap with Flux)     */
};

/* ------------------------------------------------------------------ */
/* function kinds                                                      */
/* ------------------------------------------------------------------ */
/* mu(x,p,t), sigma(x,p,t): prob.μ / prob.σ, used at src/MLP.jl:417 and
 * src/DeepSplitting.jl:160.  Element-wise (diagonal) coefficients only,
 * as the reference's broadcast expression requires. */
enum {
    HDP_COEF_CONST = 0,  /* c             p[0]=c   (test/MLP.jl:24-25)    */
    HDP_COEF_LINEAR = 1, /* c * x_i       p[0]=c   (test/MLP.jl:191-192)  */
    HDP_COEF_USER = 2,   /* hdp_user_exprs.mu / .sigma: a CUDA-C expression over
                            x (the coordinate), i (its index), t (the time the
                            reference passes: the step's end time in MLP, ts[N+1-i]
                            in sde_loop!, t_k in an EM ensemble), d, p[0..1]    */
    HDP_COEF_QUADRATIC = 3 /* c * x_i^2   p[0]=c   (test/DeepSplitting.jl:355: sigma * x .^ 2) */
};

/* g(x): terminal/initial condition prob.g, used at src/MLP.jl:203 */
enum {
    HDP_G_SUMSQ = 0,        /* sum(x^2) + p0                 (test/MLP.jl:28,118)      */
    HDP_G_INV_SUMSQ = 1,    /* 1/(p0 + p1*sum(x^2))          (test/MLP.jl:146)         */
    HDP_G_EXP_SUMSQ = 2,    /* exp(-p0*sum(x^2))             (test/MLP.jl:174; docs/src/tutorials/mlp.md:17) */
    HDP_G_MIN = 3,          /* minimum(x)                    (test/MLP.jl:199)         */
    HDP_G_GAUSS = 4,        /* (2pi)^(-d/2) p0^(-d/2) exp(-0.5*sum(x^2)/p0)  (test/MLP.jl:265) */
    HDP_G_USER = 64         /* hdp_user_exprs.g: CUDA-C expression over ssq (= sum(x^2)),
                               mn (= minimum(x)), d, p[0..3]                          */
};

/* f: nonlinearity prob.f.  Local kinds take (u_y); non-local kinds take
 * (y, z, u_y, u_z) (8-arg PIDEProblem form, src/MLP.jl:140-146).
 * f is always called with the outer end time and `nothing` gradients. */
enum {
    HDP_F_ZERO = 0,           /* 0*u                               (test/MLP.jl:35)      */
    HDP_F_LINEAR = 1,         /* p0*u                              (test/MLP.jl:122)     */
    HDP_F_ALLEN_CAHN = 2,     /* p0*(u - u^3)                      (test/MLP.jl:148,176) */
    HDP_F_DEFAULT_RISK = 3,   /* -(1-p0)*Q(u)*u - p1*u, Q piecewise linear with
                                 vh=p2 vl=p3 gh=p4 gl=p5            (test/MLP.jl:209-215) */
    HDP_F_FISHER_KPP = 4,     /* max(0,u)*(1-max(0,u))             (docs/src/tutorials/mlp.md:20) */
    HDP_F_NL_FIRST = 16,
    HDP_F_NL_ALLEN_CAHN_DIFF = 16, /* a(u_y) - a(u_z), a(u)=u-u^3  (test/MLP.jl:304)     */
    HDP_F_NL_FISHER = 17,     /* max(0,u_y)*(1-max(0,u_z))         (docs/src/tutorials/mlp.md:52) */
    HDP_F_NL_REPLICATOR = 18, /* max(0,u_y)*(m(y) - max(0,u_z)*m(z)*(2pi)^(d/2)*p0^d*exp(0.5*sum(z^2)/p0^2)),
                                 m(x) = -0.5*sum(x^2), p0 = sigma_sampling  (test/MLP.jl:266-274) */
    HDP_F_USER = 64,          /* local:     hdp_user_exprs.f over vy, ssq_y, t, d, p[0..7]          */
    HDP_F_NL_USER = 65        /* non-local: hdp_user_exprs.f over vy, vz, ssq_y, ssq_z, t, d, p[0..7] */
};

/* samplers: src/MCSample.jl:22-94 */
enum {
    HDP_SAMPLER_NONE = 0,    /* NoSampling      (src/MCSample.jl:84-86) */
    HDP_SAMPLER_UNIFORM = 1, /* UniformSampling (src/MCSample.jl:22-33) */
    HDP_SAMPLER_NORMAL = 2   /* NormalSampling  (src/MCSample.jl:55-71) */
};

/* normal generator used when no tape is injected */
enum {
    HDP_NORMAL_DEFAULT = 0,  /* BM64 for f64 entry points, BM32 for f32    */
    HDP_NORMAL_BM64 = 1,     /* Box-Muller, 53-bit uniforms, fp64 log/sincospi */
    HDP_NORMAL_BM32 = 2      /* Box-Muller, 32-bit uniforms, fp32 MUFU transcendentals */
};

/* Rounds of the counter-based generator.  The reference draws from Julia's task-local
 * Xoshiro256++ (`rand`, `randn!`: src/MLP.jl:132,416, src/DeepSplitting.jl:154), whose bit
 * stream no reference test pins; the library's own stream is Philox4x32 with 7 rounds — the
 * fewest that Salmon, Moraes, Dror, Shaw (SC'11, "Parallel random numbers: as easy as 1, 2, 3",
 * table 2) report as passing BigCrush.  The oracle's Philox twin reads the same constant. */
#ifndef HDP_PHILOX_ROUNDS
#define HDP_PHILOX_ROUNDS 7
#endif

/* random streams (one Philox stream id per consumer).  Every draw is Philox4x32-R (R above) with a fixed
 * key at counter (w0, w1, seed_lo, seed_hi); (w0, w1) packs (stream, ordinal, group):
 * TIME/STEP/SAMPLER: ordinal < 2^40 (the reference's consumption ordinal), group < 2^21
 * (vec-group of a point); SDE/X0/MC: ordinal < 2^24 (step or sample index), group < 2^37
 * (vec-group of the whole d×B array).  Larger requests return HDP_ERR_BAD_ARG. */
enum {
    HDP_STREAM_TIME = 0,     /* uniforms r   (src/MLP.jl:132)          */
    HDP_STREAM_STEP = 1,     /* normals of `_mlt_sde_loop!` (:416)     */
    HDP_STREAM_SAMPLER = 2,  /* mc_sample! draws inside MLP            */
    HDP_STREAM_SDE = 3,      /* dWall of sde_loop! (DeepSplitting.jl:154) */
    HDP_STREAM_X0 = 4,       /* x0_sample!(y1) (DeepSplitting.jl:155)  */
    HDP_STREAM_MC = 5        /* mc_sample!(z, y1) (DeepSplitting.jl:188) */
};

/* ------------------------------------------------------------------ */
/* POD descriptors                                                     */
/* ------------------------------------------------------------------ */
typedef struct hdp_sampler {
    int32_t kind;              /* HDP_SAMPLER_*                                   */
    int32_t shifted;           /* NormalSampling.shifted (src/MCSample.jl:57)     */
    int32_t bounds_are_scalar; /* UniformSampling a,b scalar (1) or length-d (0)  */
    int32_t _pad;
    double a_scalar, b_scalar; /* UniformSampling bounds when scalar              */
    double sigma;              /* NormalSampling.σ                                */
    const double *a, *b;       /* length-d HOST arrays when !bounds_are_scalar    */
} hdp_sampler;

/* PIDEProblem / ParabolicPDEProblem (src/HighDimPDE.jl:53-303) reduced to
 * what the hot path reads: μ, σ, g, f, tspan, neumann_bc. */
typedef struct hdp_problem {
    int32_t d;                 /* size(x, 1)                                       */
    int32_t mu_kind;
    int32_t sigma_kind;
    int32_t g_kind;
    int32_t f_kind;
    int32_t has_bc;            /* !isnothing(prob.neumann_bc)                      */
    int32_t bc_is_scalar;      /* neumann_bc = (lo, hi) scalars vs length-d vectors */
    int32_t _pad;
    double t0, t1;             /* prob.tspan                                       */
    double mu_p[2];
    double sigma_p[2];
    double g_p[4];
    double f_p[8];
    double lo_scalar, hi_scalar;
    const double *lo, *hi;     /* length-d HOST arrays when has_bc && !bc_is_scalar */
} hdp_problem;

/* User-supplied function bodies for the *_USER kinds: CUDA-C
 * expressions compiled with NVRTC into the K3 kernels on first use and cached in the
 * handle.  Arbitrary Julia closures still cannot cross the ABI, but any g, f, μ, σ that
 * can be written as an expression of the quantities the hot path carries can.
 * NULL members are unused.  hdp_mlp_solve_* takes them through hdp_mlp_cfg.user (all four
 * kinds); hdp_sde_loop_*, hdp_mlt_sde_loop_* and hdp_em_ensemble_* take mu / sigma from
 * hdp_set_user_exprs (their signatures mirror reference closures that have no such slot). */
typedef struct hdp_user_exprs {
    const char* mu;
    const char* sigma;
    const char* g;
    const char* f;
} hdp_user_exprs;

#define HDP_MAX_LEVELS 6 /* largest L of MLP(M, L, K) the kernels are built for */

/* MLP(M, L, K, mc_sample) (src/MLP.jl:24-32) plus the knobs a GPU needs. */
typedef struct hdp_mlp_cfg {
    int32_t M, L, K;
    int32_t n_solves;          /* independent solves in this call; solve i uses seed+i */
    int32_t n_x0;              /* number of start points in x0: 1 (shared) or n_solves */
    int32_t normal_mode;       /* HDP_NORMAL_*                                     */
    int32_t compat_mlt_threads;/* 0: serial `_ml_picard` semantics (src/MLP.jl:102-208);
                                  NT>0: `_ml_picard_mlt` with Threads.nthreads()==NT,
                                  including the `_get_loop_num` partition (src/MLP.jl:218-404) */
    int32_t target_units;      /* tree-splitting target (0 = automatic)            */
    uint64_t seed;
    hdp_sampler sampler;       /* alg.mc_sample!                                   */
    /* Parity mode: injected random tapes in the reference's consumption
     * order (HOST pointers of the entry point's element type, n_solves==1).
     * r_tape[R], n_tape[C*d], s_tape[S*d] with R, C, S from hdp_mlp_counts.
     * All three NULL => counter-based Philox generated in registers. */
    const void *r_tape, *n_tape, *s_tape;
    const hdp_user_exprs* user; /* required iff the problem uses a *_USER kind, else NULL */
    /* --- launch tuning (0 = automatic; results never depend on either) --- */
    int32_t group_size;        /* lanes that own one path: 1, 2, 4, 8, 16 or 32                    */
    uint32_t mem_flags;        /* HDP_MEM_DEVICE: x0, u_out, std_err_out and the tapes are device
                                  pointers (CuArray); | HDP_ASYNC: enqueue only — stats then carry
                                  only the host-side fields and errors surface at hdp_sync        */
    double unit_oversub;       /* work units per resident lane group the tree is cut into          */
    /* --- Monte-Carlo error bar (the reference returns none) --- */
    void* std_err_out;         /* nullable, n_solves values of the entry point's element type: the
                                  standard error of each solve, sqrt(sum_l Var_l / n_l) over the
                                  root call's level sums and its g-term (their samples are i.i.d.
                                  within a level, levels are independent)                          */
} hdp_mlp_cfg;

typedef struct hdp_mlp_counts {
    uint64_t path_steps;       /* C(L): calls of `_mlt_sde_loop!` per solve        */
    uint64_t time_uniforms;    /* R(L): calls of rand(tType) per solve             */
    uint64_t sampler_calls;    /* S(L): calls of mc_sample! per solve              */
    uint64_t f_evals;
    uint64_t g_evals;
} hdp_mlp_counts;

typedef struct hdp_mlp_stats {
    hdp_mlp_counts per_solve;
    uint64_t bounces;          /* total wall hits over all solves (reflect)         */
    uint64_t n_units;          /* work units per solve after tree splitting         */
    uint64_t n_nodes;          /* explicit call nodes per solve                     */
    uint64_t n_sites;          /* explicit sites per solve                          */
    double device_ms;          /* CUDA-event time of the device work of this call   */
    int32_t group_size;        /* lanes that own one path                           */
    int32_t n_launches;        /* kernels launched by this call                     */
    double phase_ms[4];        /* device time of: expand (top-down geometry), units,
                                  all-reduce of the partial sums, combine (bottom-up) */
    /* root call of solve 0, filled when cfg->std_err_out is given: entry l < L is level l of
     * `_ml_picard`'s outer loops (src/MLP.jl:126-196; samples b3/K, weight (t-s)/M^(L-l)), entry L
     * the g-term (:198-205; samples g(x2), weight 1/M^L)                                         */
    uint64_t level_n[HDP_MAX_LEVELS + 1];     /* samples taken                                     */
    double level_sum[HDP_MAX_LEVELS + 1];     /* sum of the samples                                */
    double level_sumsq[HDP_MAX_LEVELS + 1];   /* sum of their squares                              */
    double std_error;                         /* of solve 0                                        */
    uint64_t exchange_bytes;   /* bytes each rank hands to the all-reduce (0 on one GPU)           */
    int32_t exchange_kind;     /* 0 none (one GPU), 1 ncclAllReduce, 2 the library's exchange kernel
                                  over peer windows (hdp_peer_window_*)                            */
    int32_t _pad;
} hdp_mlp_stats;

typedef struct hdp_devinfo {
    int32_t device;
    int32_t sm_count;
    int32_t cc_major, cc_minor;
    int32_t clock_khz;
    int32_t nccl_rank, nccl_nranks;
    int32_t _pad;
    uint64_t total_mem;
    char name[128];
} hdp_devinfo;

/* entry points: host only (the NVRTC build of the kernels sees the descriptors above) */
#ifndef __CUDACC_RTC__
typedef struct hdp_context* hdp_handle;

/* ------------------------------------------------------------------ */
/* handle                                                              */
/* ------------------------------------------------------------------ */
/* One handle per host task (Julia task / Python thread).  Owns the CUDA
 * stream, scratch buffers and (optionally) an NCCL communicator.
 * Calls on one handle are serialised by an internal lock.
 * device < 0 selects the current device. */
int hdp_create(hdp_handle* out, int device);
int hdp_destroy(hdp_handle h);
const char* hdp_last_error(hdp_handle h); /* h may be NULL: last create error */
int hdp_version(void);
int hdp_device_info(hdp_handle h, hdp_devinfo* out);
int hdp_sync(hdp_handle h);

/* Multi-GPU: one process (or handle) per GPU.  Rank 0 obtains a 128-byte
 * NCCL unique id and the host language distributes it (Distributed.jl /
 * MPI / torch.distributed); every rank then joins.  Replaces the
 * `Threads.@spawn` + `fetch` + `sum` of src/MLP.jl:235-259.  With a
 * communicator attached, hdp_mlp_solve_* shards its work units over the
 * ranks and all-reduces the per-unit partial sums (the only collective). */
int hdp_nccl_unique_id(uint8_t id_out[128]);
int hdp_comm_init_rank(hdp_handle h, const uint8_t id[128], int nranks, int rank);
int hdp_comm_destroy(hdp_handle h);
/* sum-allreduce of n doubles in place (host buffer) — plumbing check */
int hdp_comm_allreduce_f64(hdp_handle h, double* buf, size_t n);

/* Peer exchange window (ranks on ONE node, after hdp_comm_init_rank): the exchange of a sharded solve is a few
 * KB — one double per crown node — and a solve lasts 0.3-1.5 ms, so the 35-50 us an ncclAllReduce of that size
 * takes is all latency.  With windows attached the library's own exchange kernel does it instead: every rank
 * stores its doubles straight into a slot of every peer's window over NVLink (P2P stores through a CUDA-IPC
 * mapping), publishes a release flag per peer, waits for the peers' flags and sums the slots of its own window in
 * rank order — the same bits on every rank — and hands whole-call leaves back to their parent sites in the same
 * launch.  hdp_peer_window_export allocates this rank's window and returns its 64-byte IPC handle; the host
 * language gathers the handles of all ranks (rank order) and every rank attaches.  Collective: either all
 * ranks attach or all detach (hdp_peer_window_detach) — a rank that cannot map a peer (no P2P path) returns an
 * error and the caller must detach everywhere; solves then use the ncclAllReduce.  Exchanges larger than the
 * window slot (HDP_PEER_SLOT_DOUBLES per solve call) use NCCL as well.  A peer that does not arrive within
 * 20 s fails the call with HDP_ERR_NCCL instead of hanging the GPU. */
#define HDP_PEER_HANDLE_BYTES 64
#define HDP_PEER_MAX_RANKS 16
#define HDP_PEER_SLOT_DOUBLES 8192
int hdp_peer_window_export(hdp_handle h, uint8_t handle_out[HDP_PEER_HANDLE_BYTES]);
int hdp_peer_window_attach(hdp_handle h, const uint8_t* handles /* nranks x 64 bytes, rank order */, int nranks, int rank);
int hdp_peer_window_detach(hdp_handle h);

/* ------------------------------------------------------------------ */
/* K4: reflection operator                                             */
/* replaces `_reflect(a,b,s,e)` both methods (src/reflect.jl:8-55, 58-82),
 * called from src/MLP.jl:419 and src/DeepSplitting.jl:162.            */
/* a, b, out: d×B (B=1 for the Vector method).  lo/hi: scalars
 * (bounds_are_scalar) or length-d arrays in the same memory space as a.
 * bounce_count_out: nullable d×B int32, wall hits per coordinate.
 * Returns HDP_ERR_START_OUTSIDE_DOMAIN if any a is outside [lo,hi].    */
/* ------------------------------------------------------------------ */
int hdp_reflect_f64(hdp_handle h, const double* a, const double* b, double* out,
                    int d, int B, const double* lo, const double* hi,
                    int bounds_are_scalar, int32_t* bounce_count_out, unsigned mem_flags);
int hdp_reflect_f32(hdp_handle h, const float* a, const float* b, float* out,
                    int d, int B, const float* lo, const float* hi,
                    int bounds_are_scalar, int32_t* bounce_count_out, unsigned mem_flags);

/* ------------------------------------------------------------------ */
/* K2: Monte-Carlo samplers                                            */
/* replaces the call operators of UniformSampling / NormalSampling
 * (src/MCSample.jl:28-33, 63-71) as used at src/DeepSplitting.jl:155,188.
 * z: d×B×K output.  y1: d×B reference points (only read when the sampler
 * is a shifted NormalSampling; may be NULL otherwise).
 * inject: nullable d×B×K raw draws (uniforms in [0,1) or standard normals)
 * to use instead of the in-register Philox stream (parity mode).        */
/* ------------------------------------------------------------------ */
int hdp_mc_sample_f64(hdp_handle h, double* z, const double* y1, int d, int B, int K,
                      const hdp_sampler* sampler, uint64_t seed, const double* inject,
                      unsigned mem_flags);
int hdp_mc_sample_f32(hdp_handle h, float* z, const float* y1, int d, int B, int K,
                      const hdp_sampler* sampler, uint64_t seed, const float* inject,
                      unsigned mem_flags);

/* ------------------------------------------------------------------ */
/* K1: batched reflected Euler–Maruyama loop                           */
/* replaces the `sde_loop!(y0,y1,dWall)` closure of
 * src/DeepSplitting.jl:153-166 (+ the `y1 .= x0_batch` of :182 is the
 * caller's: y1 holds the start points on entry).
 * y0 (out): state before the last step; y1 (in/out): start → final state.
 * nsteps = size(dWall,3); ts[i] (i=0..nsteps-1) is the time passed to
 * μ,σ at step i, i.e. Julia's ts[N+1-i] (:157).  Always a HOST array (a
 * collected Julia range), whatever mem_flags says.  May be NULL for the
 * enumerated kinds (all time-independent); required for HDP_COEF_USER.
 * x0_sampler: prob.x0_sample, applied to y1 before the loop (:155);
 * NULL or kind NONE leaves y1 untouched.
 * dWall_inject: nullable d×B×nsteps standard normals (the reference's
 * dWall after randn!, :154).  NULL => Philox in registers.
 * mc / z: optional fused epilogue `mc_sample!(z, y1)` (:186-189);
 * z is d×B×K, both NULL to skip.                                       */
/* ------------------------------------------------------------------ */
int hdp_sde_loop_f64(hdp_handle h, const hdp_problem* prob, double* y0, double* y1,
                     int B, int nsteps, double dt, const double* ts,
                     const hdp_sampler* x0_sampler, uint64_t seed,
                     const double* dWall_inject,
                     const hdp_sampler* mc, double* z, int K,
                     int normal_mode, unsigned mem_flags);
int hdp_sde_loop_f32(hdp_handle h, const hdp_problem* prob, float* y0, float* y1,
                     int B, int nsteps, float dt, const float* ts,
                     const hdp_sampler* x0_sampler, uint64_t seed,
                     const float* dWall_inject,
                     const hdp_sampler* mc, float* z, int K,
                     int normal_mode, unsigned mem_flags);

/* Several epochs of paths per launch.  In DeepSplitting's training loop
 * (src/DeepSplitting.jl:181-195) every epoch calls `y1 .= x0_batch; sde_loop!(y0, y1, dWall)` and
 * then takes one gradient step; the paths do not depend on the network, so E epochs' worth can be
 * simulated by one launch and consumed one slice per epoch.  y0, y1: d×B×E (y1 holds the start
 * points of every epoch on entry), z: d×B×K×E.  Slice e is bit for bit what
 * hdp_sde_loop_*(…, seed + e, dWall_inject = NULL, …) returns for it.  1 <= n_epochs <= 65535. */
int hdp_sde_loop_epochs_f64(hdp_handle h, const hdp_problem* prob, double* y0, double* y1,
                            int B, int nsteps, double dt, const double* ts,
                            const hdp_sampler* x0_sampler, uint64_t seed, int n_epochs,
                            const hdp_sampler* mc, double* z, int K, unsigned mem_flags);
int hdp_sde_loop_epochs_f32(hdp_handle h, const hdp_problem* prob, float* y0, float* y1,
                            int B, int nsteps, float dt, const float* ts,
                            const hdp_sampler* x0_sampler, uint64_t seed, int n_epochs,
                            const hdp_sampler* mc, float* z, int K, unsigned mem_flags);

/* User expressions for the K1 entry points (hdp_sde_loop_*, hdp_mlt_sde_loop_*,
 * hdp_em_ensemble_*) when prob->mu_kind / sigma_kind is HDP_COEF_USER.  The strings are
 * copied into the handle (the caller's pointers are not retained); NULL clears them.
 * The kernels are built by NVRTC on first use and cached per handle and expression set.
 * With a USER kind hdp_sde_loop_* requires `ts` (the reference passes t = ts[N+1-i] to
 * μ and σ, src/DeepSplitting.jl:157-160). */
int hdp_set_user_exprs(hdp_handle h, const hdp_user_exprs* user);

/* ------------------------------------------------------------------ */
/* K1e: plain Euler–Maruyama ensemble                 */
/* replaces the pre-simulation stage of NNKolmogorov (src/NNKolmogorov.jl:81-112:
 * `solve(EnsembleProblem(SDEProblem(μ, σ, xi[:,1], tspan), prob_func), EM(), ...;
 * dt, trajectories, adaptive=false)`, only `sol.u[end]` kept) and of NNStopping
 * (src/NNStopping.jl:90-99: the same with `saveat = dt`, `Array(sim)` kept).
 * Step (StochasticDiffEq v7 `EM`, a dependency that is not vendored in the reference
 * tree): dW = sqrt(dt)·ξ;  u = uprev + dt·μ(uprev,p,t_k) + σ(uprev,p,t_k)·dW, evaluated
 * left to right, t_k = t0 + k·dt accumulated.  Diagonal noise, no reflection.
 * x (in/out): d×B, column b = start point of trajectory b (NNKolmogorov: xi[:, b];
 * NNStopping: prob.x repeated) → its final state.
 * traj: nullable d×(nsteps+1)×B, `Array(sim)`'s layout, traj[:, 1, b] = start point.
 * normals_inject: nullable d×B×nsteps standard normals ξ (parity mode); NULL =>
 * Philox in registers (stream HDP_STREAM_SDE, ordinal = step, like hdp_sde_loop_*). */
/* ------------------------------------------------------------------ */
int hdp_em_ensemble_f64(hdp_handle h, const hdp_problem* prob, double* x, int B, int nsteps, double dt,
                        double t0, double* traj, uint64_t seed, const double* normals_inject,
                        unsigned mem_flags);
int hdp_em_ensemble_f32(hdp_handle h, const hdp_problem* prob, float* x, int B, int nsteps, float dt,
                        float t0, float* traj, uint64_t seed, const float* normals_inject,
                        unsigned mem_flags);

/* ------------------------------------------------------------------ */
/* One Euler–Maruyama step of the MLP recursion, stand-alone              */
/* replaces `_mlt_sde_loop!(x2, x, s, t, prob, neumann_bc)` (src/MLP.jl:406-421):
 * x2 = x + (μ(x,p,t)*dt + σ(x,p,t)*sqrt(dt)*ξ), dt = t - s, then `_reflect(x, x2, bc...)`.
 * Note the association differs from the DeepSplitting step (hdp_sde_loop_*).
 * x, x2: d×B (B independent columns; the reference is B = 1).
 * normals_inject: nullable d×B standard normals (`randn!(x2)`, :416); NULL => Philox
 * (stream HDP_STREAM_SDE, ordinal 0, draw index = linear element index).    */
/* ------------------------------------------------------------------ */
int hdp_mlt_sde_loop_f64(hdp_handle h, const hdp_problem* prob, double* x2, const double* x, int B,
                         double s, double t, uint64_t seed, const double* normals_inject, unsigned mem_flags);
int hdp_mlt_sde_loop_f32(hdp_handle h, const hdp_problem* prob, float* x2, const float* x, int B,
                         float s, float t, uint64_t seed, const float* normals_inject, unsigned mem_flags);

/* ------------------------------------------------------------------ */
/* K3: multilevel Picard                                               */
/* replaces `_ml_picard` / `_ml_picard_mlt` / `_ml_picard_call`
 * (src/MLP.jl:102-388) as called from solve(::MLP) (src/MLP.jl:68-98),
 * including `_mlt_sde_loop!` (:406-421), Vector `_reflect` and the
 * sampler calls inside the recursion.
 * x0: d × cfg->n_x0 HOST array (column-major).  u_out: n_solves HOST.
 * stats: nullable.                                                     */
/* ------------------------------------------------------------------ */
/* consumption counts per solve (tape lengths): R uniforms, C*d normals, S*d sampler
 * draws.  compat_mlt_threads as in hdp_mlp_cfg (0 = serial). */
int hdp_mlp_counts_get(int M, int L, int K, int nonlocal, int compat_mlt_threads, hdp_mlp_counts* out);
int hdp_mlp_solve_f64(hdp_handle h, const hdp_problem* prob, const hdp_mlp_cfg* cfg,
                      const double* x0, double* u_out, hdp_mlp_stats* stats);
int hdp_mlp_solve_f32(hdp_handle h, const hdp_problem* prob, const hdp_mlp_cfg* cfg,
                      const float* x0, float* u_out, hdp_mlp_stats* stats);
/* Compiles the user expressions exactly as hdp_mlp_solve_* would (NVRTC, sm_100a) without
 * touching a device: validation on a machine without a GPU.  is_f32: element type;
 * log (nullable) receives the compiler log.  Returns HDP_OK or HDP_ERR_BAD_ARG.          */
int hdp_user_kinds_check(const hdp_problem* prob, const hdp_user_exprs* user, int is_f32,
                         char* log, size_t log_cap);
/* the same for the K1 entry points: compiles sde_loop! / EM-ensemble kernels with user mu / sigma */
int hdp_user_coefs_check(const hdp_problem* prob, const hdp_user_exprs* user, int is_f32,
                         char* log, size_t log_cap);

/* ------------------------------------------------------------------ */
/* Random tapes: the exact numbers the in-register generator produces,
 * materialised (HOST output) so a CPU oracle can be fed identical draws.
 * stream: HDP_STREAM_*; ordinal0: first ordinal; n_ord × width values,
 * value (o, i) = draw i of ordinal ordinal0+o.                         */
/* ------------------------------------------------------------------ */
int hdp_tape_fill_f64(hdp_handle h, double* out, int stream, int want_normal, int normal_mode,
                      uint64_t seed, uint64_t ordinal0, uint64_t n_ord, int width);
int hdp_tape_fill_f32(hdp_handle h, float* out, int stream, int want_normal, int normal_mode,
                      uint64_t seed, uint64_t ordinal0, uint64_t n_ord, int width);

/* ------------------------------------------------------------------ */
/* Measurement helpers (roofline denominators, measured live)          */
/* ------------------------------------------------------------------ */
/* Saturating DFMA (which=64) / FFMA (which=32) microbenchmark: returns the
 * measured vector-pipe peak in lane-instructions per second.           */
int hdp_measure_pipe_peak(hdp_handle h, int which, double* lane_instr_per_s);
/* CUDA-event time (ms) of the kernels launched by the last entry-point call on this
 * handle (K1/K2/K3/K4); synchronises the handle's stream. */
int hdp_last_device_ms(hdp_handle h, double* ms);
#endif /* __CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* HDPDE_H */
