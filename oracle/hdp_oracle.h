/*
 * hdp_oracle.h — public interface of the CPU oracle (TEST INFRASTRUCTURE ONLY;
 * see the header of hdp_oracle.c).  Loaded with ctypes by oracle/oracle.py.
 */
#ifndef HDP_ORACLE_H
#define HDP_ORACLE_H
#include <stdint.h>
#include "../include/hdpde.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Random source, consumed in the reference's order:
 *   uniform  ↔ rand(tType)              src/MLP.jl:132,157,290,315
 *   normals  ↔ randn!(x2)               src/MLP.jl:416
 *   sampler  ↔ rand!/randn!(x_mc)       src/MCSample.jl:30,64            */
typedef struct hdp_oracle_rng {
    void* ctx;
    double (*uniform)(void* ctx);
    void (*normals)(void* ctx, double* out, int d);
    void (*sampler)(void* ctx, double* out, int d);
    int kind; /* 1 tape, 2 philox-by-ordinal, 3 xoshiro+ziggurat */
} hdp_oracle_rng;

/* user-defined kinds (HDP_*_USER): plain C function pointers, set by the tests from a
 * gcc-compiled twin of the CUDA-C expressions handed to the library */
typedef double (*hdp_oracle_user_coef)(double x, int i, double t, int d, const double* p);
typedef double (*hdp_oracle_user_g)(double ssq, double mn, int d, const double* p);
typedef double (*hdp_oracle_user_f)(double vy, double vz, double ssq_y, double ssq_z, double t, int d, const double* p);
typedef struct hdp_oracle_user_fns {
    hdp_oracle_user_coef mu, sigma;
    hdp_oracle_user_g g;
    hdp_oracle_user_f f;
} hdp_oracle_user_fns;
extern hdp_oracle_user_fns hdp_oracle_user;
void hdp_oracle_set_user_fns(void* mu, void* sigma, void* g, void* f);

hdp_oracle_rng* hdp_oracle_rng_tape(int is_f32, const void* r, uint64_t rn, const void* n, uint64_t nn,
                                    const void* s, uint64_t sn);
void hdp_oracle_rng_tape_pos(hdp_oracle_rng* g, uint64_t out[4]);
hdp_oracle_rng* hdp_oracle_rng_philox(uint64_t seed, int is_f32, int sampler_normal);
hdp_oracle_rng* hdp_oracle_rng_xoshiro(uint64_t seed, int is_f32, int sampler_normal);
void hdp_oracle_rng_free(hdp_oracle_rng* g);
void hdp_oracle_rng_fill(hdp_oracle_rng* g, int what, double* out, int n);

void hdp_oracle_philox4x32(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]);
int hdp_oracle_philox_rounds(void);
void hdp_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double hdp_oracle_draw_f64(uint64_t seed, int stream, uint64_t ord, uint64_t i, int want_normal);
float hdp_oracle_draw_f32(uint64_t seed, int stream, uint64_t ord, uint64_t i, int want_normal);
void hdp_oracle_mlp_counts(int M, int L, int K, int nonlocal, int64_t out[5]);

#define HDP_ORACLE_DECL(REAL, SFX)                                                                     \
    int hdp_oracle_reflect_vector##SFX(const REAL* a, const REAL* b, const REAL* s, const REAL* e,     \
                                       int d, REAL* out, int32_t* bounces, int64_t* total_bounces);    \
    int hdp_oracle_reflect_array##SFX(const REAL* a, const REAL* b, const REAL* s, const REAL* e,      \
                                      int bounds_are_scalar, int d, int B, REAL* out,                  \
                                      int32_t* bounces, int32_t* col_iters, int32_t* iters_out);       \
    int hdp_oracle_mc_sample##SFX(const hdp_sampler* sm, const REAL* raw, REAL* z, const REAL* y1,     \
                                  int d, int B, int K);                                                \
    int64_t hdp_oracle_get_loop_num##SFX(int M, int64_t num, int thread_id, int nthreads);             \
    int hdp_oracle_ml_picard##SFX(const hdp_problem* pb, const hdp_sampler* sm, int M, int L, int K,   \
                                  const REAL* x, hdp_oracle_rng* rng, REAL* u_out, int64_t* counts);   \
    int hdp_oracle_ml_picard_mlt##SFX(const hdp_problem* pb, const hdp_sampler* sm, int M, int L,      \
                                      int K, const REAL* x, int nthreads, hdp_oracle_rng** rngs,       \
                                      int parallel, REAL* u_out, int64_t* counts);                     \
    int hdp_oracle_sde_loop##SFX(const hdp_problem* pb, REAL* y0, REAL* y1, int B, int nsteps,         \
                                 REAL dt, const REAL* ts, const hdp_sampler* x0s, const REAL* x0_raw,  \
                                 const REAL* dWall, int32_t* iters_per_step);                         \
    int hdp_oracle_em_ensemble##SFX(const hdp_problem* pb, REAL* x, int B, int nsteps, REAL dt,        \
                                    REAL t0, REAL* traj, const REAL* xi);

HDP_ORACLE_DECL(double, _f64)
HDP_ORACLE_DECL(float, _f32)

int hdp_oracle_bench_solves_f64(const hdp_problem* pb, const hdp_sampler* sm, int M, int L, int K,
                                const double* x, int n_solves, int nthreads, uint64_t seed,
                                double* u_out, double* seconds);

#ifdef __cplusplus
}
#endif
#endif
