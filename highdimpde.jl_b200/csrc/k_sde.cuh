// k_sde.cuh — K1: the batched Euler–Maruyama kernels (device code only, so that the NVRTC build
// for user-defined coefficients compiles the very same text: user_kinds.cu).
//   sde_loop_kernel    — `sde_loop!` of DeepSplitting with reflection (src/DeepSplitting.jl:153-166)
//                        and the stand-alone `_mlt_sde_loop!` (src/MLP.jl:406-421)
//   em_ensemble_kernel — plain Euler–Maruyama ensembles, the pre-simulation stage of NNKolmogorov /
//                        NNStopping (src/NNKolmogorov.jl:81-112, src/NNStopping.jl:90-99)
// One thread owns VEC consecutive elements of the column-major d×B state (one 128-bit load/store,
// one Philox call) and keeps them in registers across all steps.
#pragma once
#include "hdp_device.cuh"

namespace hdp {

template <typename T, bool ALIGNED>
__device__ __forceinline__ Vec<T> load_el(const T* p, size_t q0, size_t n)
{
    constexpr int VEC = Tr<T>::VEC;
    if constexpr (ALIGNED) return ldv<T>(p + q0);
    else {
        Vec<T> r;
#pragma unroll
        for (int e = 0; e < VEC; ++e) r.v[e] = (q0 + e < n) ? p[q0 + e] : (T)0;
        return r;
    }
}
template <typename T, bool ALIGNED>
__device__ __forceinline__ void store_el(T* p, size_t q0, size_t n, const Vec<T>& r)
{
    constexpr int VEC = Tr<T>::VEC;
    if constexpr (ALIGNED) stv<T>(p + q0, r);
    else {
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            if (q0 + e < n) p[q0 + e] = r.v[e];
    }
}

template <typename T>
__device__ __forceinline__ T sample_one(const DevSampler<T>& sm, T raw, int i, T xref)
{
    if (sm.kind == HDP_SAMPLER_UNIFORM) {
        T a = sm.scalar ? sm.a_s : sm.a[i];
        T b = sm.scalar ? sm.b_s : sm.b[i];
        return uniform_map<T>(raw, a, b);
    }
    return normal_map<T>(raw, sm.sigma, sm.shifted != 0, xref);
}


template <int N> struct KTag { static constexpr int value = N; };
// μ(x,p,t)[i] (which = 0) or σ(x,p,t)[i] (which = 1) of a kind known at compile time: no switch in the step loop
template <int KIND, typename T> __device__ __forceinline__ T coef_k(const T* p, T x, int i, T t, int d, int which)
{
    if constexpr (KIND == HDP_COEF_LINEAR) return mul_rn(p[0], x);
    else if constexpr (KIND == HDP_COEF_QUADRATIC) return mul_rn(p[0], mul_rn(x, x));
    else if constexpr (KIND == HDP_COEF_USER) return which ? user_sigma<T>(x, i, t, d, p) : user_mu<T>(x, i, t, d, p);
    else return p[0];
}
// calls fn(KTag<kind>) for a run-time kind (once per thread, outside the step loop).  The NVRTC
// build knows both kinds when it compiles (HDP_FIXED_MU_KIND / HDP_FIXED_SIGMA_KIND): one copy.
template <int WHICH, typename F> __device__ __forceinline__ void with_kind(int kind, F&& fn)
{
#if defined(HDP_FIXED_MU_KIND) && defined(HDP_FIXED_SIGMA_KIND)
    (void)kind;
    if constexpr (WHICH) fn(KTag<HDP_FIXED_SIGMA_KIND>{});
    else fn(KTag<HDP_FIXED_MU_KIND>{});
#else
    switch (kind) {
    case HDP_COEF_LINEAR: fn(KTag<HDP_COEF_LINEAR>{}); break;
    case HDP_COEF_QUADRATIC: fn(KTag<HDP_COEF_QUADRATIC>{}); break;
    case HDP_COEF_USER: fn(KTag<HDP_COEF_USER>{}); break;
    default: fn(KTag<HDP_COEF_CONST>{}); break;
    }
#endif
}

// ------------------------------ K1 -----------------------------------
// N reflected EM steps per element; normals from dWall (INJECT) or Philox in
// registers.  draw addressing: dWall[q + n_el*step] == draw(SDE, ordinal=step, i=q).
// Bounds, coefficient products and validity flags are resolved into registers before
// the step loop, so the loop body is: 1 Philox call, VEC x (Box–Muller, 3 flops, mirror).
// BC: K1_GENERIC — any coefficient kind, any bounds, either association;
//     K1_NOBC / K1_SYM — constant coefficients, DeepSplitting association, no boundary /
//     symmetric scalar box (the BASELINE shape): nothing but the arithmetic is left in the loop.
//     K1_NOBC0 / K1_SYM0 — the same with μ = 0 (every BASELINE configuration): `y0 + μ*dt` is y0, one addition
//     per coordinate-step less to issue, the same value.
enum { K1_GENERIC = 0, K1_NOBC = 1, K1_SYM = 2, K1_NOBC0 = 3, K1_SYM0 = 4 };
#ifndef K1_STEP_UNROLL
#define K1_STEP_UNROLL 2
#endif

template <typename T, bool ALIGNED, bool INJECT, int BCX>
__global__ void __launch_bounds__(128)
sde_loop_kernel(DevProblem<T> pb, T* __restrict__ y0, T* __restrict__ y1, size_t n_el, int nsteps, T dt,
                DevSampler<T> x0s, const T* __restrict__ x0_inject, u64 seed, const T* __restrict__ dW,
                DevSampler<T> mc, T* __restrict__ z, int K, int* flags, unsigned long long* bounce_ctr,
                int mlp_assoc, int zero, const T* __restrict__ ts)
{
    constexpr int VEC = Tr<T>::VEC;
    constexpr bool MU0 = BCX == K1_NOBC0 || BCX == K1_SYM0;
    constexpr int BC = MU0 ? BCX - 2 : BCX;
    if constexpr (sizeof(T) == 8) tables_init();
    // blockIdx.y: the epoch of a multi-epoch launch (hdp_sde_loop_epochs_*): epoch e works on slice e of every
    // array with seed + e, exactly as a call of its own would — DeepSplitting's paths do not depend on the
    // network, so several training epochs' worth are simulated per launch (src/DeepSplitting.jl:181-195)
    if (const size_t ep = blockIdx.y) {
        y0 += ep * n_el; y1 += ep * n_el;
        if (dW) dW += ep * n_el * (size_t)nsteps;
        if (x0_inject) x0_inject += ep * n_el;
        if (z) z += ep * n_el * (size_t)K;
        seed += ep;
    }
    const int d = pb.d;
    const size_t nv = (n_el + VEC - 1) / VEC;
    const T sq = sqrt_rn(dt);
    const bool has_bc = BC == K1_GENERIC ? pb.has_bc != 0 : BC == K1_SYM;
    const T mudt_c = mul_rn(pb.mu_c, dt), sgsq_c = mul_rn(pb.sigma_c, sq);
    bool bad = false;
    int nb_total = 0;
    const DrawK<T> dk = draw_k_init<T>(zero);
    [[maybe_unused]] const uint32_t far = sym_far_bits<T>(pb.hi_s);
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (size_t)gridDim.x * blockDim.x) {
        const size_t q0 = v * VEC;
        const int i0 = n_el <= 0xffffffffull ? (int)((uint32_t)q0 % (uint32_t)d) : (int)(q0 % (size_t)d);
        Vec<T> cur = load_el<T, ALIGNED>(y1, q0, n_el);
        // x0_sample!(y1), 1-argument sampler form (src/DeepSplitting.jl:155)
        if (x0s.kind != HDP_SAMPLER_NONE) {
            Vec<T> raw = x0_inject ? load_el<T, ALIGNED>(x0_inject, q0, n_el)
                                   : draw_vec<T>(seed, HDP_STREAM_X0, 0, (u64)v, x0s.kind == HDP_SAMPLER_NORMAL);
            int i = i0;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                if (x0s.kind == HDP_SAMPLER_UNIFORM) cur.v[e] = sample_one<T>(x0s, raw.v[e], i, (T)0);
                else cur.v[e] = normal_map<T>(raw.v[e], x0s.sigma, false, (T)0);
                if (++i == d) i = 0;
            }
        }
        Vec<T> lo, hi;
        if (has_bc) {
            int i = i0;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                lo.v[e] = pb.bc_is_scalar ? pb.lo_s : pb.lo[i];
                hi.v[e] = pb.bc_is_scalar ? pb.hi_s : pb.hi[i];
                // @assert all((a .>= s) .& (a .<= e))  (src/reflect.jl:59): a = y0 of the first step
                if (q0 + e < n_el && !(cur.v[e] >= lo.v[e] && cur.v[e] <= hi.v[e])) bad = true;
                if (++i == d) i = 0;
            }
        }
        Vec<T> prev = cur;
        int st = 0;
        if constexpr (BC != K1_GENERIC && !INJECT && K1_STEP_UNROLL > 1) {
            // the normals of consecutive steps do not depend on the state: several Philox / Box–Muller chains in flight
            constexpr int SU = K1_STEP_UNROLL;
            for (; st + SU - 1 < nsteps; st += SU) {
                Vec<T> xs[SU];
#pragma unroll
                for (int u = 0; u < SU; ++u) xs[u] = draw_vec<T>(seed, HDP_STREAM_SDE, (u64)(st + u), (u64)v, true, dk);
#pragma unroll
                for (int u = 0; u < SU; ++u) {
                    prev = cur; // y0 .= y1 (:159)
#pragma unroll
                    for (int e = 0; e < VEC; ++e)
                        cur.v[e] = add_rn(MU0 ? prev.v[e] : add_rn(prev.v[e], mudt_c), mul_rn(sgsq_c, xs[u].v[e])); // :160
                    if constexpr (BC == K1_SYM) reflect_vec_sym<T, false>(cur, pb.hi_s, far, nb_total); // :161-163
                }
            }
        }
        // the remaining steps (all of them in the generic mode: one copy of the loop per pair of
        // coefficient kinds, so that no kind is decided inside it)
        auto steps = [&](auto mu_tag, auto sg_tag) {
            [[maybe_unused]] constexpr int MU = decltype(mu_tag)::value;
            [[maybe_unused]] constexpr int SG = decltype(sg_tag)::value;
            for (; st < nsteps; ++st) {
                Vec<T> xi;
                if constexpr (INJECT) xi = load_el<T, ALIGNED>(dW + (size_t)st * n_el, q0, n_el);
                else xi = draw_vec<T>(seed, HDP_STREAM_SDE, (u64)st, (u64)v, true, dk);
                prev = cur; // y0 .= y1 (:159)
                if constexpr (BC == K1_GENERIC) {
                    // μ(y0, p, t), σ(y0, p, t) with t = ts[N + 1 - i] (:157), or the end time of the one
                    // `_mlt_sde_loop!` step (src/MLP.jl:417)
                    const T t = ts ? ts[st] : (T)0;
                    int i = i0;
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const T x = prev.v[e];
                        const T mudt = mul_rn(coef_k<MU, T>(pb.mu_p, x, i, t, d, 0), dt);
                        const T sgsq = mul_rn(coef_k<SG, T>(pb.sigma_p, x, i, t, d, 1), sq);
                        if (++i == d) i = 0;
                        if (mlp_assoc) cur.v[e] = add_rn(x, add_rn(mudt, mul_rn(sgsq, xi.v[e]))); // src/MLP.jl:417
                        else cur.v[e] = add_rn(add_rn(x, mudt), mul_rn(sgsq, xi.v[e])); // :160 (y0 + μ*dt) + (σ*sq)*dW
                    }
                    if (has_bc) reflect_vec<T, 1, false>(cur, lo, hi, 0u, nb_total); // :161-163
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e)
                        cur.v[e] = add_rn(MU0 ? prev.v[e] : add_rn(prev.v[e], mudt_c), mul_rn(sgsq_c, xi.v[e])); // :160
                    if constexpr (BC == K1_SYM) reflect_vec_sym<T, false>(cur, pb.hi_s, far, nb_total); // :161-163
                }
            }
        };
        if constexpr (BC == K1_GENERIC)
            with_kind<0>(pb.mu_kind, [&](auto m) { with_kind<1>(pb.sigma_kind, [&](auto sg) { steps(m, sg); }); });
        else steps(KTag<HDP_COEF_CONST>{}, KTag<HDP_COEF_CONST>{});
        store_el<T, ALIGNED>(y0, q0, n_el, prev);
        store_el<T, ALIGNED>(y1, q0, n_el, cur);
        // fused epilogue mc_sample!(z, y1) (src/DeepSplitting.jl:186-189)
        if (z != nullptr && mc.kind != HDP_SAMPLER_NONE) {
            const bool normal = mc.kind == HDP_SAMPLER_NORMAL;
            for (int k = 0; k < K; ++k) {
                Vec<T> raw = draw_vec<T>(seed, HDP_STREAM_MC, (u64)k, (u64)v, normal);
                int i = i0;
                Vec<T> o;
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    o.v[e] = sample_one<T>(mc, raw.v[e], i, cur.v[e]);
                    if (++i == d) i = 0;
                }
                store_el<T, ALIGNED>(z + (size_t)k * n_el, q0, n_el, o);
            }
        }
    }
    if (bad) atomicExch(&flags[0], 1);
    // one atomic per warp, not per thread: a single-address atomic serialises
    nb_total = __reduce_add_sync(0xffffffffu, nb_total);
    if (bounce_ctr && nb_total && (threadIdx.x & 31) == 0) atomicAdd(bounce_ctr, (unsigned long long)nb_total);
}


// ------------------------------ K1e ----------------------------------
// Plain Euler–Maruyama ensemble: B independent trajectories of dX = μ(X,p,t) dt + σ(X,p,t) dW
// (diagonal noise) on a fixed grid t_k = t0 + k dt — what `solve(EnsembleProblem(SDEProblem(μ, σ,
// x0, tspan)), EM(); dt, trajectories)` computes for NNKolmogorov (final states,
// src/NNKolmogorov.jl:81-112) and NNStopping (whole trajectories, `saveat = dt`,
// src/NNStopping.jl:90-99).  The step is StochasticDiffEq's (third-party, v7, not vendored in the
// reference tree): dW = sqrt(dt) ξ, u = uprev + dt μ(uprev,p,t) + σ(uprev,p,t) dW, i.e.
// (uprev + dt μ) + σ (sqrt(dt) ξ) — a third association, next to :160 and src/MLP.jl:417.
// x (in/out): d×B start → final states; traj (TRAJ): d×(nsteps+1)×B as `Array(sim)` lays it out.
// Normals: dW_inject[q + n_el*step] (INJECT) or draw(SDE, ordinal = step, i = q) like sde_loop!.
template <typename T, bool ALIGNED, bool INJECT, bool TRAJ>
__global__ void __launch_bounds__(128)
em_ensemble_kernel(DevProblem<T> pb, T* __restrict__ x, size_t n_el, int nsteps, T dt, T t0, T* __restrict__ traj,
                   u64 seed, const T* __restrict__ dW, int zero)
{
    constexpr int VEC = Tr<T>::VEC;
    if constexpr (sizeof(T) == 8) tables_init();
    const int d = pb.d;
    const size_t nv = (n_el + VEC - 1) / VEC;
    const T sq = sqrt_rn(dt);
    const DrawK<T> dk = draw_k_init<T>(zero);
    // one copy of the step loop per (μ kind, σ kind): nothing but the arithmetic is left inside
    with_kind<0>(pb.mu_kind, [&](auto mu_tag) {
        with_kind<1>(pb.sigma_kind, [&](auto sg_tag) {
            constexpr int MU = decltype(mu_tag)::value, SG = decltype(sg_tag)::value;
            for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (size_t)gridDim.x * blockDim.x) {
                const size_t q0 = v * VEC;
                const int i0 = n_el <= 0xffffffffull ? (int)((uint32_t)q0 % (uint32_t)d) : (int)(q0 % (size_t)d);
                Vec<T> cur = load_el<T, ALIGNED>(x, q0, n_el);
                // element q = i + d b of the state sits at q + d (k + nsteps b) of the trajectory array
                // (ALIGNED: d % VEC == 0, so the VEC elements share their column)
                T* tp[ALIGNED ? 1 : VEC] = {};
                if constexpr (TRAJ) {
#pragma unroll
                    for (int e = 0; e < (ALIGNED ? 1 : VEC); ++e) {
                        const size_t q = q0 + e < n_el ? q0 + e : q0;
                        tp[e] = traj + q + (size_t)d * (size_t)nsteps * (q / (size_t)d);
                    }
                }
                auto save = [&]() {
                    if constexpr (ALIGNED) {
                        stv<T>(tp[0], cur);
                        tp[0] += d;
                    } else {
#pragma unroll
                        for (int e = 0; e < VEC; ++e) {
                            if (q0 + e < n_el) *tp[e] = cur.v[e];
                            tp[e] += d;
                        }
                    }
                };
                if constexpr (TRAJ) save();
                T t = t0;
                for (int st = 0; st < nsteps; ++st) {
                    Vec<T> xi;
                    if constexpr (INJECT) xi = load_el<T, ALIGNED>(dW + (size_t)st * n_el, q0, n_el);
                    else xi = draw_vec<T>(seed, HDP_STREAM_SDE, (u64)st, (u64)v, true, dk);
                    int i = i0;
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const T xe = cur.v[e];
                        const T mu = coef_k<MU, T>(pb.mu_p, xe, i, t, d, 0);
                        const T sg = coef_k<SG, T>(pb.sigma_p, xe, i, t, d, 1);
                        cur.v[e] = add_rn(add_rn(xe, mul_rn(dt, mu)), mul_rn(sg, mul_rn(sq, xi.v[e])));
                        if (++i == d) i = 0;
                    }
                    t = add_rn(t, dt);
                    if constexpr (TRAJ) save();
                }
                store_el<T, ALIGNED>(x, q0, n_el, cur);
            }
        });
    });
}

} // namespace hdp
