// hdp_host_util.h — host helpers: POD descriptor → device descriptor.
#pragma once
#include <math.h>
#include <vector>
#include "hdp_context.h"
#include "hdp_convert.h"

namespace hdp {

inline int validate_problem(hdp_context* h, const hdp_problem* p, bool allow_user = false)
{
    if (!p) return set_err(h, HDP_ERR_BAD_ARG, "problem is NULL");
    if (p->d <= 0) return set_err(h, HDP_ERR_BAD_ARG, "problem.d = %d", p->d);
    auto coef_ok = [&](int k) {
        return k == HDP_COEF_CONST || k == HDP_COEF_LINEAR || k == HDP_COEF_QUADRATIC || (allow_user && k == HDP_COEF_USER);
    };
    if (!coef_ok(p->mu_kind) || !coef_ok(p->sigma_kind))
        return set_err(h, HDP_ERR_BAD_ARG, "unknown mu/sigma kind%s", allow_user ? "" : " (this entry point takes no user-defined coefficients)");
    if (p->has_bc && !p->bc_is_scalar && (!p->lo || !p->hi))
        return set_err(h, HDP_ERR_BAD_ARG, "vector neumann_bc without lo/hi arrays");
    return HDP_OK;
}

// Uploads vector bounds (if any) into handle buffer `which` (2*d elements).
template <typename T>
int make_dev_problem(hdp_context* h, const hdp_problem* p, int which, DevProblem<T>* out, bool allow_user = false)
{
    int rc = validate_problem(h, p, allow_user);
    if (rc != HDP_OK) return rc;
    fill_dev_problem<T>(p, out);
    DevProblem<T>& q = *out;
    bool uniform_vec = false;
    if (p->has_bc && !p->bc_is_scalar) {
        // neumann_bc = [fill(lo, d), fill(hi, d)] (test/MLP.jl:86): identical to scalar bounds
        uniform_vec = true;
        for (int i = 1; i < p->d; ++i)
            if (p->lo[i] != p->lo[0] || p->hi[i] != p->hi[0]) { uniform_vec = false; break; }
        if (uniform_vec) {
            q.bc_is_scalar = 1;
            q.lo_s = (T)p->lo[0];
            q.hi_s = (T)p->hi[0];
        }
    }
    if (p->has_bc && !p->bc_is_scalar && !uniform_vec) {
        std::vector<T> tmp(2 * (size_t)p->d);
        for (int i = 0; i < p->d; ++i) { tmp[i] = (T)p->lo[i]; tmp[p->d + i] = (T)p->hi[i]; }
        T* d = (T*)ensure(h, which, tmp.size() * sizeof(T));
        if (!d) return HDP_ERR_ALLOC;
        cudaError_t e = cudaMemcpyAsync(d, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream); // tmp is pageable and goes out of scope
        if (e != cudaSuccess) return set_err(h, HDP_ERR_CUDA, "bounds upload: %s", cudaGetErrorString(e));
        q.lo = d;
        q.hi = d + p->d;
    }
    if (p->has_bc && p->bc_is_scalar && !(p->lo_scalar < p->hi_scalar))
        return set_err(h, HDP_ERR_BAD_ARG, "neumann_bc: lo >= hi");
    return HDP_OK;
}

template <typename T>
int make_dev_sampler(hdp_context* h, const hdp_sampler* s, int d, int which, DevSampler<T>* out)
{
    DevSampler<T>& q = *out;
    if (s && s->kind != HDP_SAMPLER_NONE && s->kind != HDP_SAMPLER_UNIFORM && s->kind != HDP_SAMPLER_NORMAL)
        return set_err(h, HDP_ERR_BAD_ARG, "unknown sampler kind %d", s->kind);
    fill_dev_sampler<T>(s, out);
    if (!s || s->kind == HDP_SAMPLER_NONE) return HDP_OK;
    if (s->kind == HDP_SAMPLER_UNIFORM && !s->bounds_are_scalar) {
        if (!s->a || !s->b) return set_err(h, HDP_ERR_BAD_ARG, "UniformSampling vector bounds missing");
        bool uni = true;
        for (int i = 1; i < d; ++i)
            if (s->a[i] != s->a[0] || s->b[i] != s->b[0]) { uni = false; break; }
        if (uni) { // UniformSampling(fill(a, d), fill(b, d)): identical to scalar bounds
            q.scalar = 1;
            q.a_s = (T)s->a[0];
            q.b_s = (T)s->b[0];
            return HDP_OK;
        }
        std::vector<T> tmp(2 * (size_t)d);
        for (int i = 0; i < d; ++i) { tmp[i] = (T)s->a[i]; tmp[d + i] = (T)s->b[i]; }
        T* dv = (T*)ensure(h, which, tmp.size() * sizeof(T));
        if (!dv) return HDP_ERR_ALLOC;
        cudaError_t e = cudaMemcpyAsync(dv, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) return set_err(h, HDP_ERR_CUDA, "sampler bounds upload: %s", cudaGetErrorString(e));
        q.a = dv;
        q.b = dv + d;
    }
    return HDP_OK;
}

} // namespace hdp
