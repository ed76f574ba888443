// hdp_convert.h — POD descriptor → device descriptor field conversion (no CUDA calls).
#pragma once
#include <math.h>
#include "hdp_device.cuh"

namespace hdp {

// Scalar fields only; lo/hi (and sampler a/b) device pointers are set by the caller.
template <typename T> void fill_dev_problem(const hdp_problem* p, DevProblem<T>* out)
{
    DevProblem<T>& q = *out;
    q.d = p->d;
    q.mu_kind = p->mu_kind;
    q.sigma_kind = p->sigma_kind;
    q.g_kind = p->g_kind;
    q.f_kind = p->f_kind;
    q.has_bc = p->has_bc;
    q.bc_is_scalar = p->bc_is_scalar;
    q.t0 = (T)p->t0;
    q.t1 = (T)p->t1;
    q.mu_c = (T)p->mu_p[0];
    q.sigma_c = (T)p->sigma_p[0];
    for (int i = 0; i < 2; ++i) { q.mu_p[i] = (T)p->mu_p[i]; q.sigma_p[i] = (T)p->sigma_p[i]; }
    for (int i = 0; i < 4; ++i) q.g_p[i] = (T)p->g_p[i];
    for (int i = 0; i < 8; ++i) q.f_p[i] = (T)p->f_p[i];
    q.f_pd[0] = q.f_pd[1] = 0;
    q.g_pd[0] = q.g_pd[1] = 1;
    if (p->f_kind == HDP_F_NL_REPLICATOR) {
        double ss = p->f_p[0];
        q.f_pd[0] = pow(2.0 * M_PI, p->d / 2.0) * pow(ss, (double)p->d);
        q.f_pd[1] = 0.5 / (ss * ss);
    }
    if (p->g_kind == HDP_G_GAUSS) {
        double ss0 = p->g_p[0];
        q.g_pd[0] = pow(2.0 * M_PI, -p->d / 2.0) * pow(ss0, -p->d * 0.5);
        q.g_pd[1] = ss0;
    }
    q.lo_s = (T)p->lo_scalar;
    q.hi_s = (T)p->hi_scalar;
    q.lo = q.hi = nullptr;
}

template <typename T> void fill_dev_sampler(const hdp_sampler* s, DevSampler<T>* out)
{
    DevSampler<T>& q = *out;
    q.kind = HDP_SAMPLER_NONE;
    q.shifted = 0;
    q.scalar = 1;
    q.a_s = q.b_s = q.sigma = (T)0;
    q.a = q.b = nullptr;
    if (!s || s->kind == HDP_SAMPLER_NONE) return;
    q.kind = s->kind;
    q.shifted = s->shifted;
    q.scalar = s->bounds_are_scalar;
    q.a_s = (T)s->a_scalar;
    q.b_s = (T)s->b_scalar;
    q.sigma = (T)s->sigma;
}

} // namespace hdp
