// k_mlp.cuh — K3: the multilevel-Picard tree on the GPU.
//
// Replaces `_ml_picard` / `_ml_picard_mlt` / `_ml_picard_call`
// (src/MLP.jl:102-388), `_mlt_sde_loop!` (:406-421), the Vector `_reflect`
// (src/reflect.jl:8-55) and the sampler calls inside the recursion.
//
// Mapping.  A group of G lanes (G = 1..32, a power of two) owns one work unit;
// lane g holds vec-groups g, g+G, g+2G, … of every point (one vec-group = VEC
// coordinates = one Philox call = one 128-bit load/store).  The recursion is
// the compile-time family Call<LV>/eval_site<l>; the points x2, x3 of the
// active site of level l live in scratch slot 2l / 2l+1, laid out
// [slot][vec-group j][thread] so every access is a coalesced 128-bit
// transaction (SoA).  All draws are addressed by the reference's consumption
// ordinal, so results do not depend on G, on the cut, or on the GPU count.
#pragma once
#include "hdp_device.cuh"
#include "mlp_plan_pod.h"

namespace hdp {

// Launch tuning of the units kernel (measured on B200, BASELINE C1/C2/C4/C5): points with many
// vec-groups per lane run best with three or four of them in flight per pass — independent
// Philox / Box–Muller chains — at 4 blocks per SM (128 registers): four for rows of >= 48 vec-groups
// per lane (d = 100 with one lane per point, d = 1000 with eight), three for rows of 12..47 (wide
// lane groups in the latency regime; a fourth chain lost 5-20 % there); short rows (d = 10: five
// vec-groups per lane) with one chain at 6 blocks per SM (80 registers).  U is a template
// parameter of the kernel; the host picks it from the row length.
#ifndef HDP_MLP_LONG_MINBLOCKS
#define HDP_MLP_LONG_MINBLOCKS 4
#endif
template <int U> struct Tune {
    static constexpr int MB = U >= 2 ? HDP_MLP_LONG_MINBLOCKS : 6;
};
#ifndef HDP_MLP_BLOCK
#define HDP_MLP_BLOCK 128 // threads per block of the expand and units kernels
#endif
#ifndef HDP_EM_PRELOAD
#define HDP_EM_PRELOAD 1
#endif
#ifndef HDP_EM_PREFETCH
#define HDP_EM_PREFETCH 0 // long rows: 1 / 2 = prefetch the start points of the next pass into L1 / L2 (measured: profiles/r2a_*)
#endif
template <int N> struct IntTag { static constexpr int value = N; };
struct FullTag { static constexpr bool value = true; };
struct PartTag { static constexpr bool value = false; };

template <typename T> struct MlpParams {
    DevProblem<T> pb;
    DevSampler<T> sm;
    int M, K, L, Kp;
    int nvec, npl, dpad;
    u64 C[kMaxL + 1], R[kMaxL + 1], S[kMaxL + 1];
    u64 cs[kMaxL + 1], rs[kMaxL + 1], ss[kMaxL + 1];
    u64 Mpow[kMaxL + 2];
    const T *r_tape, *n_tape, *s_tape;
    u64 seed0;
    int n_solves, n_x0;
    const PlanUnit* units;
    const int32_t* exec;      // the units this rank evaluates, most expensive first
    int n_exec;
    const int8_t* unit_tag;
    const PlanNode* nodes;
    const PlanSite* sites;
    const int32_t* children;
    const int32_t* sites_by_depth;
    const int32_t* nodes_by_depth;
    int n_units, n_nodes, n_sites, n_points;
    T* points;       // [n_solves][n_points][dpad]
    T* feat;         // [n_solves][n_points]  sum(x^2)
    T* site_r;       // [n_solves][n_sites]   sampled time of each explicit site
    double* partial; // [n_solves][n_units]
    T* value;        // [n_solves][n_nodes]
    T* scratch;
    u64 total_threads;
    int rank, nranks;
    // one solve sharded over several ranks (mlp_plan.h: MlpSharding); all null on one GPU
    const int8_t* unit_mine;  // per unit: evaluated on this rank
    const int8_t* site_cls;   // per site: SITE_UPPER (replicated), SITE_MINE, SITE_OTHER (bottom sites)
    const int32_t* node_slot; // per node: slot of a crown node in xbuf, else -1
    double* xbuf;             // [n_solves][xstride] crown partial values, then the root statistics;
                              // followed by kXTail doubles for the whole call.  The only buffer all-reduced.
    int xstride, n_slots;
    // Monte-Carlo error bar of the root call (cfg->std_err_out); all null when not asked for
    double* rootsq;           // [n_solves][rootsq_stride]: (sum, sum of squares) per root unit, or per root level
    int rootsq_stride;        //   when the whole solve is one unit
    double* lvstat;           // [n_solves][3*(kMaxL+1)]: n, sum, sum of squares per root level (entry L: g-term)
    T* std_err;               // [n_solves]
    int zero;        // 0, opaque to the compiler (draw_k_init)
    // combine pass, nodes with more than kBigNode units (the root of a finely cut tree has 10^4..10^5): their
    // units are summed by stage_nb blocks each first (mlp_prereduce_kernel), stage[launch item][block][kCmbStage]
    const double* stage;
    int stage_nb;
    // ... or, for a list with per-node shares (the crown of a sharded solve: the root has 10^5 units and thousands of
    // sites, most crown nodes a handful): job j = (list index, share) sums that share of the node's units AND of
    // its sites into stage[solve][j]; stage_off[k] .. stage_off[k + 1] are the jobs of list entry k
    const int32_t* stage_off;
    int stage_jobs;
    unsigned long long* work_ctr;
    unsigned long long* bounce_ctr;
    int* flags;
};

template <typename T> struct XAcc {
    T* base;
    u64 vstride, voff;
    int unchecked; // 1: the point may lie outside the box (user x0, sampled x3) -> `_reflect` asserts it (src/reflect.jl:12)
    __device__ __forceinline__ T* at(int j) const { return base + ((u64)j * vstride + voff) * Tr<T>::VEC; }
};

template <typename T, int G> struct Ctx {
    const MlpParams<T>* P;
    u64 seed;
    u64 tid;
    int lane;
    unsigned gmask;
    int nb;
    int bad;
    int fast, fast_s; // em_step (EM_* mode, 0 = general form) / sample_step may take their fast specialisation
    int f_uses_z;     // f reads the sampled point itself (sum(z^2)), not only u(z)
    void* frame_mem;  // units kernel: this lane group's Frame<T> in shared memory
    __device__ __forceinline__ XAcc<T> slot(int k) const
    {
        XAcc<T> a;
        a.base = P->scratch + (u64)k * (u64)P->npl * P->total_threads * Tr<T>::VEC;
        a.vstride = P->total_threads;
        a.voff = tid;
        a.unchecked = k & 1; // odd slots hold sampled points x3, even slots reflected points x2
        return a;
    }
    __device__ __forceinline__ XAcc<T> point(int solve, int pt) const
    {
        XAcc<T> a;
        a.base = P->points + ((u64)solve * (u64)P->n_points + (u64)pt) * (u64)P->dpad;
        a.vstride = G;
        a.voff = (u64)lane;
        a.unchecked = 0; // explicit points are checked once, where they are created (init_item, expand_item)
        return a;
    }
};

template <typename T, int G> __device__ __forceinline__ T gsum(T v, unsigned mask)
{
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off);
    return v;
}
template <typename T, int G> __device__ __forceinline__ T gmin(T v, unsigned mask)
{
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
        T o = __shfl_xor_sync(mask, v, off);
        v = o < v ? o : v;
    }
    return v;
}
template <int G> __device__ __forceinline__ int gsumi(int v, unsigned mask)
{
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off);
    return v;
}

template <typename T> __device__ __forceinline__ T big() { return (T)3.0e38; }

// rand(tType)  (src/MLP.jl:132)
template <typename T> __device__ __forceinline__ T time_uniform(const MlpParams<T>& P, u64 seed, u64 ro)
{
    if (P.r_tape) return P.r_tape[ro];
    return draw_vec<T>(seed, HDP_STREAM_TIME, ro, 0, false).v[0];
}

// `_mlt_sde_loop!(x2, x, s, t, prob, neumann_bc)`  (src/MLP.jl:406-421):
// one Euler–Maruyama step s -> tend from the point `src`, then reflection.
// Writes x2 to `dst` when STORE; returns sum(x2^2) and min(x2).
//
// em_step_fast — constant coefficients, in-register RNG, g does not need min(x), and the
// boundary is absent (EM_NOBC), a symmetric scalar box (EM_SYM) or a general scalar box
// (EM_BOX): every BASELINE configuration.  The coordinate loop is 1 Philox call, the
// Box–Muller pair, 3 flops per coordinate, the mirror, one 128-bit load and (STORE) store.
// The start-point assertion of `_reflect` (src/reflect.jl:12) runs as a separate pass, only
// for points that can be outside the box (user x0, sampled x3).
enum { EM_NOBC = 1, EM_SYM = 2, EM_BOX = 3 };

// The coordinate loop itself; returns this lane's share of sum(x2^2).
// A vec-group in which a coordinate is still outside after its first mirror — a second wall hit
// in one step, ≈0.5 % of the coordinates of BASELINE C2 — is only flagged in a bit mask and
// parked in a small per-thread stash (local memory: write-back cached, unlike a global store),
// and finished after the loop.  The loop body therefore has no branch, no call and no
// convergence barrier.
template <typename T, int G, int MODE, int U = 1, bool MU0 = false>
__device__ __forceinline__ T em_loop_fast(const MlpParams<T>& P, const DrawK<T>& dk, u64 seed, int lane, const T* sp,
                                          uint32_t sstride, T* dp, uint32_t dstride, bool store, bool check_src, T s,
                                          T tend, u64 ord, int& nb, int& bad)
{
    constexpr int VEC = Tr<T>::VEC;
    constexpr bool BC = MODE != EM_NOBC;
#ifndef HDP_HOST_EMU
    // points and scratch rows are cudaMalloc'ed: keep the compiler on LDG/STG (it falls back to generic
    // LD/ST when it loses track of the address space across the two copies of this loop)
    __builtin_assume(__isGlobal(sp));
    __builtin_assume(__isGlobal(dp));
#endif
    const int d = P.pb.d, nvec = P.nvec;
    const T lo_s = MODE == EM_SYM ? -P.pb.hi_s : P.pb.lo_s, hi_s = P.pb.hi_s;
    const T dt = sub_rn(tend, s);  // :415
    const T sq = sqrt_rn(dt);
    const T mudt = mul_rn(P.pb.mu_c, dt), sgsq = mul_rn(P.pb.sigma_c, sq);
    const int nfull = d / VEC;
    [[maybe_unused]] const uint32_t far = sym_far_bits<T>(hi_s);
    if (BC && check_src) { // src/reflect.jl:12
        const T* q = sp;
        for (int p = lane; p < nvec; p += G, q += sstride) {
            const Vec<T> xv = ldv<T>(q);
#pragma unroll
            for (int e = 0; e < VEC; ++e)
                if (p * VEC + e < d && !(xv.v[e] >= lo_s && xv.v[e] <= hi_s)) bad = 1;
        }
    }
    T ssq = (T)0;
    unsigned fix = 0; // bit j: the j-th vec-group of the current chunk needs its reflection finished
    Vec<T> stash[32]; // … and waits here
    int it = 0, p0 = lane;
    T* dp0 = dp;
    Vec<T> xnext;
    if (U == 1 && HDP_EM_PRELOAD) {
        if (lane < nvec) xnext = ldv<T>(sp);
        else
#pragma unroll
            for (int e = 0; e < VEC; ++e) xnext.v[e] = (T)0;
    }
    // U vec-groups of this lane (p, p+G, …) per pass
    auto body = [&](int p, auto full_tag, auto u_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        constexpr int UU = decltype(u_tag)::value;
        Vec<T> xi[UU], o[UU];
#pragma unroll
        for (int u = 0; u < UU; ++u) xi[u] = draw_vec<T>(seed, HDP_STREAM_STEP, ord, (u64)(p + u * G), true, dk); // :416 randn!(x2)
        if (U == 1 && HDP_EM_PRELOAD) {
            // short rows, one vec-group per pass: the start point of the next pass is requested now,
            // a whole pass before it is used (its address does not depend on the draws)
            o[0] = xnext;
            sp += sstride;
            if (p + G < nvec) xnext = ldv<T>(sp);
        } else {
#pragma unroll
            for (int u = 0; u < UU; ++u) {
                o[u] = ldv<T>(sp);
                sp += sstride;
            }
#if HDP_EM_PREFETCH && !defined(HDP_HOST_EMU)
            if (UU > 1) {
#pragma unroll
                for (int u = 0; u < UU; ++u) {
                    if (p + (UU + u) * G >= nvec) continue; // only vec-groups of this row
                    if (HDP_EM_PREFETCH == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(sp + (size_t)u * sstride));
                    else asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + (size_t)u * sstride));
                }
            }
#endif
        }
        const unsigned bit0 = 1u << it; // this pass's first bit of `fix`
#pragma unroll
        for (int u = 0; u < UU; ++u) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                // :417  x + (μ dt + (σ sq) ξ); with μ = 0 the inner sum is its second term (0 + y = y), one rounding less to issue
                const T dx = mul_rn(sgsq, xi[u].v[e]);
                o[u].v[e] = add_rn(o[u].v[e], MU0 ? dx : add_rn(mudt, dx));
                if (!FULL && !((p + u * G) * VEC + e < d)) o[u].v[e] = MODE == EM_BOX ? lo_s : (T)0; // padding never reflects
            }
            bool again = false;
            if (MODE == EM_SYM) again = reflect_first_sym<T>(o[u], hi_s, nb) >= far; // :418-420
            if (MODE == EM_BOX) again = reflect_first<T>(o[u], lo_s, hi_s, nb);
            T q = ssq;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                if (!FULL && !((p + u * G) * VEC + e < d)) o[u].v[e] = (T)0; // padding coordinate
                q = xfma(o[u].v[e], o[u].v[e], q);
            }
            ssq = again ? ssq : q;
            if (BC && again) {
                fix |= bit0 << u;
                stash[it + u] = o[u];
            }
            // the stride of a scratch row is the same for every lane: one pointer bump per pass, below
            if (store) stv<T>(dp + (size_t)u * dstride, o[u]);
        }
        it += UU;
        dp += (size_t)UU * dstride; // (never dereferenced when !store)
    };
    // second and later wall hits of the flagged vec-groups.  The stash lives in local memory: the entry of the next
    // flagged vec-group is requested before this one is worked on (waiting for the load at the first compare of
    // reflect_coord was 7 % of the kernel's stall samples, r2u capture).
    auto finish = [&]() {
        if (!fix) return;
        auto lowest = [](unsigned m) {
#ifdef HDP_HOST_EMU
            return __builtin_ctz(m);
#else
            return __ffs((int)m) - 1;
#endif
        };
        if constexpr (U == 1) { // short rows (d = 10): one or two flagged vec-groups at most, the plain loop is faster (A/B: r2v)
            while (fix) {
                const int j = lowest(fix);
                fix &= fix - 1;
                Vec<T> o = stash[j];
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    if ((p0 + j * G) * VEC + e < d) o.v[e] = reflect_coord<T>(o.v[e], lo_s, hi_s, nb);
                    ssq = xfma(o.v[e], o.v[e], ssq);
                }
                if (store) stv<T>(dp0 + (size_t)j * dstride, o);
            }
            return;
        }
        int j = lowest(fix);
        Vec<T> o = stash[j];
        for (;;) {
            fix &= fix - 1;
            const int jn = fix ? lowest(fix) : j;
            const Vec<T> on = stash[jn];
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                if ((p0 + j * G) * VEC + e < d) o.v[e] = reflect_coord<T>(o.v[e], lo_s, hi_s, nb);
                ssq = xfma(o.v[e], o.v[e], ssq);
            }
            if (store) stv<T>(dp0 + (size_t)j * dstride, o);
            if (!fix) break;
            o = on;
            j = jn;
        }
    };
    const int nrag = (nvec % G) ? (nvec / G) * G : nvec; // vec-groups below this index form rows that every lane owns
    int p = lane;
    while (p < nvec) {
        // at most 32 vec-groups per chunk: one bit each in `fix`
        it = 0;
        dp0 = dp;
        p0 = p;
        // rows every lane of the group owns first: all lanes then run the same passes (a lane with one vec-group
        // more than its neighbours would drag the whole warp through its own tail passes)
        const int lim = nfull < nrag ? nfull : nrag;
        const int pe = p + 32 * G < lim ? p + 32 * G : lim;
        if (U > 1)
            for (; p + (U - 1) * G < pe; p += U * G) body(p, FullTag{}, IntTag<U>{});
        if (U > 2) // what is left of the chunk, still two at a time (d = 100: 50 vec-groups = 12 passes of four + one of two)
            for (; p + G < pe; p += 2 * G) body(p, FullTag{}, IntTag<2>{});
        for (; p < pe; p += G) body(p, FullTag{}, IntTag<1>{});
        if (p >= nrag && p < nfull && it < 32) { // the ragged last row: only some lanes own a vec-group there
            body(p, FullTag{}, IntTag<1>{});
            p += G;
        }
        if (p >= nfull && p < nvec && it < 32) { // the one partial vec-group (d % VEC != 0), if this lane owns it
            body(p, PartTag{}, IntTag<1>{});
            p += G;
        }
        if (BC) finish();
    }
    return ssq;
}

// the stand-alone step of the expand pass (x2 is always kept); rows of >= 12 vec-groups per lane
// take three per pass, like the units kernel
template <typename T, int G, int MODE>
__device__ __forceinline__ Feat<T> em_step_fast(Ctx<T, G>& c, const MlpParams<T>& P, const XAcc<T>& src, const XAcc<T>& dst,
                                                T s, T tend, u64 ord)
{
    constexpr int VEC = Tr<T>::VEC;
    const DrawK<T> dk = draw_k_init<T>(P.zero & (c.lane + 1));
    int nb = 0, bad = 0;
    const T* sp = src.base + src.voff * VEC;
    T* dp = dst.base + dst.voff * VEC;
    const uint32_t ss = (uint32_t)src.vstride * VEC, ds = (uint32_t)dst.vstride * VEC;
    const T ssq = P.npl >= 12 ? em_loop_fast<T, G, MODE, 3>(P, dk, c.seed, c.lane, sp, ss, dp, ds, true, src.unchecked != 0,
                                                           s, tend, ord, nb, bad)
                              : em_loop_fast<T, G, MODE, 1>(P, dk, c.seed, c.lane, sp, ss, dp, ds, true, src.unchecked != 0,
                                                           s, tend, ord, nb, bad);
    c.nb += nb;
    c.bad |= bad;
    Feat<T> f;
    f.ssq = gsum<T, G>(ssq, c.gmask);
    f.mn = big<T>();
    return f;
}

// the general form: injected tapes, vector bounds, state-dependent and user-defined coefficients, min(x)
template <typename T, int G>
__device__ HDP_NOINLINE Feat<T> em_step_generic(Ctx<T, G>& c, const XAcc<T>& src, const XAcc<T>* dst, T s, T tend, u64 ord)
{
    constexpr int VEC = Tr<T>::VEC;
    const MlpParams<T>& P = *c.P;
    const int d = P.pb.d, nvec = P.nvec;
    const int lane = c.lane;
    const unsigned gmask = c.gmask;
    const u64 seed = c.seed;
    const T* __restrict__ tape = P.n_tape;
    const bool has_bc = P.pb.has_bc != 0, bc_scalar = P.pb.bc_is_scalar != 0;
    const bool check_src = has_bc && src.unchecked != 0;
    const T* __restrict__ lo_v = P.pb.lo;
    const T* __restrict__ hi_v = P.pb.hi;
    const T lo_s = P.pb.lo_s, hi_s = P.pb.hi_s;
    const bool mu_lin = P.pb.mu_kind == HDP_COEF_LINEAR;
    const bool sg_lin = P.pb.sigma_kind == HDP_COEF_LINEAR;
    const bool mu_usr = P.pb.mu_kind == HDP_COEF_USER;
    const bool sg_usr = P.pb.sigma_kind == HDP_COEF_USER;
    const T mu_c = P.pb.mu_c, sg_c = P.pb.sigma_c;
    const bool want_min = P.pb.g_kind == HDP_G_MIN || P.pb.g_kind == HDP_G_USER;
    const T dt = sub_rn(tend, s);  // :415
    const T sq = sqrt_rn(dt);
    const T mudt_c = mul_rn(mu_c, dt), sgsq_c = mul_rn(sg_c, sq); // constant-coefficient products
    const T* sp = src.base + src.voff * VEC;
    const u64 sstride = src.vstride * VEC;
    T* dp = dst ? dst->base + dst->voff * VEC : nullptr;
    const u64 dstride = dst ? dst->vstride * VEC : 0;
    T ssq = (T)0, mn = big<T>();
    int nb = 0, bad = 0;
    for (int p = lane; p < nvec; p += G) {
        Vec<T> xi;
        if (tape) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const int i = p * VEC + e;
                xi.v[e] = i < d ? tape[ord * (u64)d + (u64)i] : (T)0;
            }
        } else xi = draw_vec<T>(seed, HDP_STREAM_STEP, ord, (u64)p, true); // :416 randn!(x2)
        const Vec<T> xv = ldv<T>(sp);
        sp += sstride;
        Vec<T> o, lo, hi;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const T x = xv.v[e];
            T mudt = mu_lin ? mul_rn(mul_rn(mu_c, x), dt) : mudt_c;
            T sgsq = sg_lin ? mul_rn(mul_rn(sg_c, x), sq) : sgsq_c;
            // μ(x,p,t), σ(x,p,t) are evaluated at the start point and the END time (:417)
            if (mu_usr) mudt = mul_rn(user_mu<T>(x, p * VEC + e, tend, d, P.pb.mu_p), dt);
            if (sg_usr) sgsq = mul_rn(user_sigma<T>(x, p * VEC + e, tend, d, P.pb.sigma_p), sq);
            o.v[e] = add_rn(x, add_rn(mudt, mul_rn(sgsq, xi.v[e]))); // :417  x + (μ*dt + (σ*sq)*ξ)
        }
        if (has_bc) { // :418-420
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const int i = p * VEC + e;
                const bool valid = i < d;
                lo.v[e] = bc_scalar ? lo_s : (valid ? lo_v[i] : (T)0);
                hi.v[e] = bc_scalar ? hi_s : (valid ? hi_v[i] : (T)0);
                if (!valid) { lo.v[e] = -big<T>(); hi.v[e] = big<T>(); } // padding never reflects
                if (valid && check_src && !(xv.v[e] >= lo.v[e] && xv.v[e] <= hi.v[e])) bad = 1; // src/reflect.jl:12
            }
            reflect_vec<T, G>(o, lo, hi, gmask, nb);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const bool valid = p * VEC + e < d;
            const T nx = valid ? o.v[e] : (T)0;
            o.v[e] = nx;
            ssq = xfma(nx, nx, ssq);
            if (want_min) mn = (valid && nx < mn) ? nx : mn;
        }
        if (dp) {
            stv<T>(dp, o);
            dp += dstride;
        }
    }
    c.nb += nb;
    c.bad |= bad;
    Feat<T> f;
    f.ssq = gsum<T, G>(ssq, gmask);
    f.mn = want_min ? gmin<T, G>(mn, gmask) : mn;
    return f;
}

// the stand-alone step (expand pass): dispatch on the mode of this launch
template <typename T, int G>
__device__ __forceinline__ Feat<T> em_step(Ctx<T, G>& c, const MlpParams<T>& P, const XAcc<T>& src, const XAcc<T>& dst, T s,
                                           T tend, u64 ord)
{
    switch (c.fast) {
    case EM_NOBC: return em_step_fast<T, G, EM_NOBC>(c, P, src, dst, s, tend, ord);
    case EM_SYM: return em_step_fast<T, G, EM_SYM>(c, P, src, dst, s, tend, ord);
    case EM_BOX: return em_step_fast<T, G, EM_BOX>(c, P, src, dst, s, tend, ord);
    }
    return em_step_generic<T, G>(c, src, &dst, s, tend, ord);
}

// mc_sample!(x3, x2)  (src/MLP.jl:139,177; src/MCSample.jl:28-33, 63-71)
template <typename T, int G, bool FAST>
__device__ HDP_NOINLINE Feat<T> sample_step_t(Ctx<T, G>& c, const XAcc<T>* dst, const XAcc<T>& x2, u64 ord)
{
    constexpr int VEC = Tr<T>::VEC;
    const MlpParams<T>& P = *c.P;
    const int d = P.pb.d, nvec = P.nvec;
    const int lane = c.lane;
    const u64 seed = c.seed;
    const T* __restrict__ tape = FAST ? nullptr : P.s_tape;
    const bool normal = P.sm.kind == HDP_SAMPLER_NORMAL;
    const bool shifted = normal && P.sm.shifted;
    const bool scalar = FAST ? true : (P.sm.scalar != 0);
    const T* __restrict__ a_v = FAST ? nullptr : P.sm.a;
    const T* __restrict__ b_v = FAST ? nullptr : P.sm.b;
    const T a_s = P.sm.a_s, b_s = P.sm.b_s, sigma = P.sm.sigma;
    const T* rp = x2.base + x2.voff * VEC;
    const u64 rstride = x2.vstride * VEC;
    T* dp = dst ? dst->base + dst->voff * VEC : nullptr;
    const u64 dstride = dst ? dst->vstride * VEC : 0;
    T ssq = (T)0;
    for (int p = lane; p < nvec; p += G) {
        Vec<T> raw;
        if (tape) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const int i = p * VEC + e;
                raw.v[e] = i < d ? tape[ord * (u64)d + (u64)i] : (T)0;
            }
        } else raw = draw_vec<T>(seed, HDP_STREAM_SAMPLER, ord, (u64)p, normal);
        Vec<T> ref;
        if (shifted) {
            ref = ldv<T>(rp);
            rp += rstride;
        }
        Vec<T> o;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const int i = p * VEC + e;
            const bool valid = i < d;
            T v;
            if (normal) v = normal_map<T>(raw.v[e], sigma, shifted, shifted ? ref.v[e] : (T)0);
            else v = uniform_map<T>(raw.v[e], scalar ? a_s : (valid ? a_v[i] : (T)0), scalar ? b_s : (valid ? b_v[i] : (T)0));
            v = valid ? v : (T)0;
            ssq = xfma(v, v, ssq);
            o.v[e] = v;
        }
        if (dp) {
            stv<T>(dp, o);
            dp += dstride;
        }
    }
    Feat<T> f;
    f.ssq = gsum<T, G>(ssq, c.gmask);
    f.mn = big<T>();
    return f;
}
template <typename T, int G>
__device__ __forceinline__ Feat<T> sample_step(Ctx<T, G>& c, const XAcc<T>* dst, const XAcc<T>& x2, u64 ord)
{
    return c.fast_s ? sample_step_t<T, G, true>(c, dst, x2, ord) : sample_step_t<T, G, false>(c, dst, x2, ord);
}

// ---------------------------------------------------------------------
// The same recursion without calls.  At any moment at most one call of each level is active
// (a call of level LV only ever starts calls of lower levels), so the frames of the active
// calls live in an array indexed by call level, and the walk is a state machine that runs
// bookkeeping until the next path-step or sampler call is due.  There is exactly one copy of
// the coordinate loop in the kernel, it is inlined in the kernel body (parameters come straight
// from the constant bank, no call ABI in the way of register allocation), and nothing is
// re-loaded per path-step.  Arithmetic and ordinals are those of Call<LV> / eval_site<l> above.
// ---------------------------------------------------------------------
template <typename T> struct Frame {
    T s, r, a, b, b3, b2, b4, f2, f3, term;
    u64 k, num, gnum, co, ro, so, cc, rr, qq;
    int lv, l, lend, h, phase, xslot, parent, pad_[sizeof(T) == 8 ? 1 : 3];
};
// one frame per thread, side by side in shared memory: an odd number of 8-byte words per frame keeps a
// warp's accesses to one field on 16 distinct bank pairs (two wavefronts for a 64-bit field: the minimum)
static_assert(sizeof(Frame<double>) % 8 == 0 && (sizeof(Frame<double>) / 8) % 2 == 1, "Frame<double> stride");
static_assert(sizeof(Frame<float>) % 8 == 0 && (sizeof(Frame<float>) / 8) % 2 == 1, "Frame<float> stride");
enum { CMB_ALL = 0, CMB_PARTIAL = 1, CMB_FINAL = 2 };
constexpr int kXTail = 8; // xbuf tail: [0..3] 16-bit chunks of the call hash, [4] wall hits, [5..7] spare
enum { PH_B2 = 0, PH_B4 = 1, PH_V = 2, PH_W = 3 };
enum { ST_LEVEL = 0, ST_SITE, ST_SITE_POST, ST_AFTER_B2, ST_SAMPLES, ST_SAMPLE_POST, ST_SITE_END, ST_G, ST_G_POST, ST_RETURN };
enum { ACT_NONE = 0, ACT_EM = 1, ACT_SAMPLE = 2 };

template <typename T, int G, int MODE, int U>
__device__ __forceinline__ T run_unit_iter(Ctx<T, G>& c, const MlpParams<T>& P, const PlanUnit& u, const PlanNode& nd,
                                           const XAcc<T>& xp, T s0, double* rsq)
{
    constexpr int VEC = Tr<T>::VEC;
    const DevProblem<T>& pb = P.pb;
    const T t1 = pb.t1;
    const bool shifted = P.sm.kind == HDP_SAMPLER_NORMAL && P.sm.shifted;
    const bool has_samples = P.Kp > 0;
    const int K = P.K;
    const int kind = u.kind;
    const u64 seed = c.seed;
    const int lane = c.lane;
    const unsigned gmask = c.gmask;
    const bool f_uses_z = pb.f_kind == HDP_F_NL_REPLICATOR || pb.f_kind == HDP_F_NL_USER;
    const bool mu0 = MODE != 0 && pb.mu_c == (T)0;
    const DrawK<T> dk = draw_k_init<T>(P.zero & (lane + 1)); // 0, but neither constant nor uniform to the compiler
    // The frame of the innermost active call lives in shared memory, one private copy per thread
    // (the lanes of a group keep identical copies: no cell is shared, so nothing depends on the
    // lanes running in lockstep); the frames of its callers are parked in local memory and only
    // touched when a call starts or returns.
    // rsq (nullable): where this unit reports (sum, sum of squares) of the samples it takes for the
    // root call of its solve — the error bar of cfg->std_err_out.  A slice unit reports one pair, a
    // unit that is the whole solve one pair per level and one for the g-term.
    Frame<T> parked[kMaxL + 1];
#ifdef HDP_HOST_EMU
    Frame<T> frame0;
    Frame<T>* const F = &frame0;
#else
    Frame<T>* const F = static_cast<Frame<T>*>(c.frame_mem);
#endif
    int cur, root;
    // sum of squares of the samples of the current level of the unit's root call (the error bar, cfg->std_err_out):
    // one per unit, so it need not travel with the frames
    double root_sq = 0.0;
    {
        const u64 n = (u64)(u.k1 - u.k0);
        root = kind == UNIT_WHOLE ? nd.lv : (kind == UNIT_G ? 1 : u.l + 1);
        F->lv = root;
        F->l = kind == UNIT_LEVEL ? u.l : 0;
        F->lend = kind == UNIT_WHOLE ? nd.lv : (kind == UNIT_G ? 0 : u.l + 1);
        F->k = 0;
        F->num = kind == UNIT_LEVEL ? n : P.Mpow[root];
        F->gnum = kind == UNIT_G ? n : P.Mpow[root];
        F->a = (T)0; F->b = (T)0;
        F->s = s0;
        F->xslot = -1;
        F->co = u.c; F->ro = u.r; F->so = u.s;
        F->parent = -1;
        cur = root;
    }
    int st = ST_LEVEL, nb = 0, bad = 0;
    T ret = (T)0, result = (T)0;
    Feat<T> feat;
    feat.ssq = (T)0; feat.mn = big<T>();
    for (;;) {
        // One pass runs the bookkeeping from the result of the last action to the next one.  The
        // handlers are ordered so that every transition except "return to the parent" lands on a
        // handler further down the same pass.
#ifdef HDP_FRAME_SHARED_EXPERIMENT
        if (G > 1) __syncwarp(gmask);
#endif
        int l = F->l;
        int act = ACT_NONE;
        // action parameters
        int a_src = 0, a_dst = 0;
        bool a_store = false;
        T a_s = (T)0, a_t = (T)0;
        u64 a_ord = 0;
        auto enter = [&](int lvc, int xslot) { // start the call U_lvc(slot, r) of the current site
            const T s = F->r;
            const u64 co = F->cc, ro = F->rr, so = F->qq;
            F->cc = co + P.C[lvc]; F->rr = ro + P.R[lvc]; F->qq = so + P.S[lvc];
            parked[cur] = *F;
            F->lv = lvc; F->l = 0; F->lend = lvc; F->k = 0;
            F->num = P.Mpow[lvc]; F->gnum = P.Mpow[lvc];
            F->a = (T)0; F->b = (T)0; F->s = s; F->xslot = xslot;
            F->co = co; F->ro = ro; F->so = so; F->parent = cur;
            cur = lvc; l = 0; st = ST_LEVEL;
        };
        if (st == ST_RETURN) {
            if (cur == root) { result = ret; break; }
            cur = F->parent;
            *F = parked[cur];
            l = F->l;
            if (F->phase == PH_B2) { F->b2 = ret; st = ST_AFTER_B2; }
            else if (F->phase == PH_B4) { F->b4 = ret; st = ST_SAMPLES; }
            else if (F->phase == PH_V) {
                const T term = f_eval<T>(pb, F->f2, F->f3, F->b2, ret); // :140-146
                if (l >= 2) {
                    F->term = term;
                    F->phase = PH_W;
                    enter(l - 1, 2 * l + 1);
                } else { F->b3 += term; F->h += 1; st = ST_SAMPLES; }
            } else {
                F->b3 += F->term - f_eval<T>(pb, F->f2, F->f3, F->b4, ret); // :178-191
                F->h += 1;
                st = ST_SAMPLES;
            }
        }
        if (st == ST_AFTER_B2) {
            if (l >= 2) { // :160-173
                F->phase = PH_B4;
                enter(l - 1, 2 * l);
            } else st = ST_SAMPLES;
        }
        if (st == ST_SITE_POST) {
            F->f2 = feat.ssq;
            F->cc = F->co + 1; F->rr = F->ro + 1; F->qq = F->so;
            F->b2 = (T)0; F->b4 = (T)0; F->b3 = (T)0; F->h = 0;
            if (l >= 1) { // :134
                F->phase = PH_B2;
                enter(l, 2 * l);
            } else st = ST_SAMPLES;
        }
        if (st == ST_SAMPLE_POST) {
            F->f3 = feat.ssq;
            F->qq += 1;
            if (l >= 1) {
                F->phase = PH_V;
                enter(l, 2 * l + 1);
            } else {
                F->b3 += f_eval<T>(pb, F->f2, F->f3, (T)0, (T)0); // :140-146 with U_0 = 0
                F->h += 1;
                st = ST_SAMPLES;
            }
        }
        if (st == ST_G_POST) {
            const T gv = g_eval<T>(pb, feat);
            F->b += gv; // :203
            if (rsq && cur == root) root_sq += (double)gv * (double)gv;
            F->k += 1;
            st = ST_G;
        }
        if (st == ST_SAMPLES) {
            const bool dead_samples = (l == 0) && !f_uses_z;
            if (has_samples && !dead_samples) {
                if (F->h < K) { // :137, :176
                    act = ACT_SAMPLE; a_dst = 2 * l + 1; a_store = l >= 1; a_src = 2 * l; a_ord = F->qq; // :139, :177
                    st = ST_SAMPLE_POST;
                } else st = ST_SITE_END;
            } else {
                // NoSampling: the K-loop re-evaluates f on identical arguments (:137-147, :210-216)
                T term = f_eval<T>(pb, F->f2, (T)0, F->b2, (T)0);
                if (l >= 2) term = term - f_eval<T>(pb, F->f2, (T)0, F->b4, (T)0);
                T b3 = (T)0;
                for (int h = 0; h < K; ++h) b3 += term;
                F->b3 = b3;
                st = ST_SITE_END;
            }
        }
        if (st == ST_SITE_END) {
            const T sv = F->b3 / (T)K;
            F->b += sv; // :148
            if (rsq && cur == root) root_sq += (double)sv * (double)sv;
            F->k += 1;
            F->co += P.cs[l]; F->ro += P.rs[l]; F->so += P.ss[l];
            st = ST_LEVEL;
        }
        if (st == ST_LEVEL) {
            for (;;) {
                if (l >= F->lend) { F->k = 0; F->b = (T)0; st = ST_G; break; }
                if (F->k < F->num) { st = ST_SITE; break; }
                if (cur == root && kind == UNIT_LEVEL) break; // the slice is done: result = b
                if (rsq && cur == root) { // a whole solve: one pair per level
                    if (lane == 0) { rsq[2 * l] = (double)F->b; rsq[2 * l + 1] = root_sq; }
                    root_sq = 0.0;
                }
                F->a += ((t1 - F->s) * F->b) / (T)F->num; // :150, :195
                l += 1;
                F->l = l; F->k = 0; F->b = (T)0;
                F->num = P.Mpow[F->lv - l];
            }
            if (st == ST_LEVEL) {
                result = F->b;
                if (rsq && lane == 0) { rsq[0] = (double)result; rsq[1] = root_sq; }
                break;
            }
        }
        if (st == ST_SITE) {
            const T uu = time_uniform<T>(P, seed, F->ro);
            const T r = add_rn(F->s, mul_rn(sub_rn(t1, F->s), uu)); // :132
            F->r = r;
            act = ACT_EM; a_src = F->xslot; a_dst = 2 * l; a_store = (l >= 1) || shifted; a_s = F->s; a_t = r; a_ord = F->co; // :133
            st = ST_SITE_POST;
        }
        if (st == ST_G) {
            if (F->k < F->gnum) { // :198-205
                act = ACT_EM; a_src = F->xslot; a_store = false; a_s = F->s; a_t = t1; a_ord = F->co + F->k; // :202
                st = ST_G_POST;
            } else {
                if (cur == root && kind == UNIT_G) {
                    result = F->b;
                    if (rsq && lane == 0) { rsq[0] = (double)result; rsq[1] = root_sq; }
                    break;
                }
                if (rsq && cur == root && lane == 0) { rsq[2 * P.L] = (double)F->b; rsq[2 * P.L + 1] = root_sq; }
                ret = F->a + F->b / (T)F->gnum; // :205, :207
                st = ST_RETURN;
            }
        }
        if (act == ACT_NONE) continue;
        const XAcc<T> xs = a_src < 0 ? xp : c.slot(a_src);
        if (act == ACT_EM) {
            if (MODE == 0) {
                const XAcc<T> xd = c.slot(a_dst);
                feat = em_step_generic<T, G>(c, xs, a_store ? &xd : nullptr, a_s, a_t, a_ord);
            } else {
                const XAcc<T> xd = c.slot(a_store ? a_dst : 0);
                // (μ = 0 — every BASELINE configuration — is a property of the launch: no lane ever takes the other copy)
                const T ssq = mu0 ? em_loop_fast<T, G, MODE, U, true>(P, dk, seed, lane, xs.base + xs.voff * VEC,
                                                                    (uint32_t)xs.vstride * VEC, xd.base + xd.voff * VEC,
                                                                    (uint32_t)xd.vstride * VEC, a_store, xs.unchecked != 0,
                                                                    a_s, a_t, a_ord, nb, bad)
                                  : em_loop_fast<T, G, MODE, U, false>(P, dk, seed, lane, xs.base + xs.voff * VEC,
                                                                     (uint32_t)xs.vstride * VEC, xd.base + xd.voff * VEC,
                                                                     (uint32_t)xd.vstride * VEC, a_store, xs.unchecked != 0,
                                                                     a_s, a_t, a_ord, nb, bad);
                feat.ssq = gsum<T, G>(ssq, gmask);
            }
        } else {
            const XAcc<T> xd = c.slot(a_dst);
            feat = sample_step<T, G>(c, a_store ? &xd : nullptr, xs, a_ord);
        }
    }
    c.nb += nb;
    c.bad |= bad;
    return result;
}

// which specialisation of the coordinate loop a launch takes: EM_* for constant coefficients, in-register
// RNG, g without min(x) and no / a scalar boundary, else 0 (the general form).  The host picks the units
// kernel compiled for exactly this mode (k_mlp_host.cu), the device uses it for everything else.
template <typename T> __host__ __device__ inline int em_mode_of(const DevProblem<T>& pb, bool normals_injected)
{
    if (pb.mu_kind == HDP_COEF_CONST && pb.sigma_kind == HDP_COEF_CONST && (!pb.has_bc || pb.bc_is_scalar) &&
        !normals_injected && pb.g_kind != HDP_G_MIN && pb.g_kind != HDP_G_USER)
        return !pb.has_bc ? EM_NOBC : (pb.lo_s == -pb.hi_s ? EM_SYM : EM_BOX);
    return 0;
}

template <typename T, int G>
__device__ __forceinline__ void ctx_init(Ctx<T, G>& c, const MlpParams<T>& P, u64 tid, int lane_w)
{
    c.P = &P;
    c.tid = tid;
    c.lane = lane_w % G;
    c.gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane_w - c.lane));
    c.nb = 0;
    c.bad = 0;
    c.seed = 0;
    c.frame_mem = nullptr;
    c.fast = em_mode_of<T>(P.pb, P.n_tape != nullptr);
    c.fast_s = P.s_tape == nullptr && (P.sm.kind != HDP_SAMPLER_UNIFORM || P.sm.scalar);
    c.f_uses_z = P.pb.f_kind == HDP_F_NL_REPLICATOR || P.pb.f_kind == HDP_F_NL_USER;
}

// ---------------------------------------------------------------------
// per-item bodies (shared by the kernels and by the test-only host emulation)
// ---------------------------------------------------------------------
// `@assert all((a .>= s) .& (a .<= e))` (src/reflect.jl:12) for an explicit point that later steps start from
template <typename T, int G>
__device__ __forceinline__ void assert_in_box(Ctx<T, G>& c, const MlpParams<T>& P, const XAcc<T>& x)
{
    constexpr int VEC = Tr<T>::VEC;
    const bool scalar = P.pb.bc_is_scalar != 0;
    for (int p = c.lane, j = 0; p < P.nvec; p += G, ++j) { // j-th vec-group of this lane
        const Vec<T> xv = ldv<T>(x.at(j));
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const int i = p * VEC + e;
            if (i < P.pb.d) {
                const T lo = scalar ? P.pb.lo_s : P.pb.lo[i], hi = scalar ? P.pb.hi_s : P.pb.hi[i];
                if (!(xv.v[e] >= lo && xv.v[e] <= hi)) c.bad = 1;
            }
        }
    }
}

// geometry of one explicit site: r, x2, x3_h
// (`list`: the explicit sites of one depth — or, when a solve is sharded, the replicated upper sites of one
// depth / the sites of one depth at and below this rank's ownership roots)
// One item per (site, solve, point): point 0 is x2, point 1 + h is x3_h.  A pass over a depth is a link in a
// chain of dependent launches and rarely fills the GPU, so its duration is the latency of one item: a site
// walked by a single lane group (x2, then K sampled points one after the other) held every launch at 25-30 us
// for a non-local problem.  A shifted sampler starts from x2: its items recompute x2 (the same bits, stored by
// all of them) instead of waiting for the first; only point 0 reports r, the feature and the wall hits of x2.
// (`split` = false: one item per (site, solve), the points one after the other — a pass over tens of thousands of
// sites fills the GPU either way, and then the shared per-site loads and index arithmetic are paid once, not 1 + K times.)
template <typename T, int G>
__device__ __forceinline__ void expand_item(Ctx<T, G>& c, const MlpParams<T>& P, const int32_t* list, u64 item,
                                            bool count_hits = true, bool split = true)
{
    const int npt = 1 + P.Kp;
    const int j = split ? (int)(item % (u64)npt) : -1; // -1: every point of the site
    if (split) item /= (u64)npt;
    const int solve = (int)(item % (u64)P.n_solves);
    const int sidx = list[item / (u64)P.n_solves];
    const PlanSite st = P.sites[sidx];
    const PlanNode nd = P.nodes[st.node];
    c.seed = P.seed0 + (u64)solve;
    const XAcc<T> x2 = c.point(solve, st.x2_point);
    if (j <= 0 || (P.sm.kind == HDP_SAMPLER_NORMAL && P.sm.shifted)) {
        const T s = nd.time_site < 0 ? P.pb.t0 : P.site_r[(u64)solve * P.n_sites + nd.time_site];
        const XAcc<T> xp = c.point(solve, nd.point);
        const T u = time_uniform<T>(P, c.seed, st.r);
        const T r = add_rn(s, mul_rn(sub_rn(P.pb.t1, s), u));
        const Feat<T> f2 = em_step<T, G>(c, P, xp, x2, s, r, st.c);
        if (j <= 0 && c.lane == 0) {
            P.site_r[(u64)solve * P.n_sites + sidx] = r;
            P.feat[(u64)solve * P.n_points + st.x2_point] = f2.ssq;
        }
        if (j > 0) c.nb = 0; // (point 0 counts the wall hits of the step)
    }
    if (j != 0 && P.Kp > 0) {
        const int l = st.l;
        const u64 sub = P.S[l] + (l >= 2 ? P.S[l - 1] : 0);
        const int h0 = j > 0 ? j - 1 : 0, h1 = j > 0 ? j : P.Kp;
        for (int h = h0; h < h1; ++h) {
            const u64 qq = st.s + sub + (u64)h * (1 + sub);
            const XAcc<T> x3 = c.point(solve, st.x3_point0 + h);
            const Feat<T> f3 = sample_step<T, G>(c, &x3, x2, qq);
            if (c.lane == 0) P.feat[(u64)solve * P.n_points + st.x3_point0 + h] = f3.ssq;
            if (P.pb.has_bc && l >= 1) assert_in_box<T, G>(c, P, x3); // U_l(x3) steps from x3 (a level-0 site never does)
        }
    }
    const int nb = gsumi<G>(c.nb, c.gmask);
    const int bad = gsumi<G>(c.bad, c.gmask);
    if (c.lane == 0) {
        if (nb && count_hits) atomicAdd(P.bounce_ctr, (unsigned long long)nb);
        if (bad) atomicExch(&P.flags[0], 1);
    }
}

// one work item w = (unit, solve): evaluate and store the partial sum
template <typename T, int G, int U = 1, int FM = -1>
__device__ __forceinline__ void unit_item(Ctx<T, G>& c, const MlpParams<T>& P, u64 w)
{
    const int solve = (int)(w % (u64)P.n_solves);
    const int uid = P.exec[w / (u64)P.n_solves];
    const PlanUnit un = P.units[uid];
    const PlanNode nd = P.nodes[un.node];
    c.seed = P.seed0 + (u64)solve;
    c.nb = 0;
    const T s = nd.time_site < 0 ? P.pb.t0 : P.site_r[(u64)solve * P.n_sites + nd.time_site];
    const XAcc<T> xp = c.point(solve, nd.point);
    // units of the root call report their samples' sum of squares when an error bar is wanted
    double* rsq = nullptr;
    if (P.rootsq && un.node == 0) rsq = P.rootsq + (u64)solve * P.rootsq_stride + (un.kind == UNIT_WHOLE ? 0 : 2 * uid);
    T v;
    // FM >= 0: a kernel compiled for one mode (the host launched it because em_mode_of says so) carries
    // only that copy of the walk; FM < 0 (NVRTC build, host emulation): decided here
    if constexpr (FM == EM_NOBC) v = run_unit_iter<T, G, EM_NOBC, U>(c, P, un, nd, xp, s, rsq);
    else if constexpr (FM == EM_SYM) v = run_unit_iter<T, G, EM_SYM, U>(c, P, un, nd, xp, s, rsq);
    else if constexpr (FM == EM_BOX) v = run_unit_iter<T, G, EM_BOX, U>(c, P, un, nd, xp, s, rsq);
    else if constexpr (FM == 0) v = run_unit_iter<T, G, 0, 1>(c, P, un, nd, xp, s, rsq);
    else {
        switch (c.fast) {
        case EM_NOBC: v = run_unit_iter<T, G, EM_NOBC, U>(c, P, un, nd, xp, s, rsq); break;
        case EM_SYM: v = run_unit_iter<T, G, EM_SYM, U>(c, P, un, nd, xp, s, rsq); break;
        case EM_BOX: v = run_unit_iter<T, G, EM_BOX, U>(c, P, un, nd, xp, s, rsq); break;
        default: v = run_unit_iter<T, G, 0, 1>(c, P, un, nd, xp, s, rsq); break;
        }
    }
    const int nb = gsumi<G>(c.nb, c.gmask);
    if (c.lane == 0) {
        P.partial[(u64)solve * P.n_units + uid] = (double)v;
        if (nb) atomicAdd(P.bounce_ctr, (unsigned long long)nb);
    }
}

// The f-term of one explicit site from its children's values: b3 / K of src/MLP.jl:139-194 (`val`, `part`, `feat`:
// the solve's node values, unit partial sums and point features; *l_out: the site's level)
template <typename T>
__device__ __forceinline__ T site_term(const MlpParams<T>& P, int si, const T* val, const double* part, const T* feat, int* l_out)
{
    const DevProblem<T>& pb = P.pb;
    const PlanSite st = P.sites[si];
    const int l = st.l;
    const int32_t* ch = P.children + st.child0;
    const T f2 = feat[st.x2_point];
    // child value: explicit node, whole-call leaf (straight from its unit's partial), or 0
    auto cv = [&](int32_t code) -> T {
        return code >= 0 ? val[code] : (code == -1 ? (T)0 : (T)part[-2 - code]);
    };
    const T b2 = cv(ch[0]);
    const T b4 = cv(ch[1]);
    T b3 = (T)0;
    if (P.Kp > 0) {
        // (unrolled: the loads of four sampled points are in flight together — a combine pass is a link in a
        // chain of dependent launches and lasts as long as one site's chain of dependent loads)
#pragma unroll 4
        for (int h = 0; h < P.K; ++h) {
            const T f3 = feat[st.x3_point0 + h];
            const T v = cv(ch[2 + 2 * h]);
            T term = f_eval<T>(pb, f2, f3, b2, v);
            if (l >= 2) {
                const T w = cv(ch[3 + 2 * h]);
                term = term - f_eval<T>(pb, f2, f3, b4, w);
            }
            b3 += term;
        }
    } else {
        T term = f_eval<T>(pb, f2, (T)0, b2, (T)0);
        if (l >= 2) term = term - f_eval<T>(pb, f2, (T)0, b4, (T)0);
        for (int h = 0; h < P.K; ++h) b3 += term;
    }
    *l_out = l;
    return b3 / (T)P.K;
}

// value of one explicit node from its units' partial sums and its explicit sites.
// `nl` lanes (a warp on the device, 1 in the host emulation) share the node: they stride over
// its units and over its explicit sites, then the per-level sums are butterfly-reduced — the
// root of a deeply cut tree has thousands of children, one thread would walk them serially.
//
// mode (one solve sharded over several ranks, MlpSharding in mlp_plan.h):
//   CMB_ALL      every unit and every site of the node -> value[node]            (one GPU; owned nodes)
//   CMB_PARTIAL  this rank's units and bottom sites of a crown node -> xbuf slot (summed by the all-reduce)
//   CMB_FINAL    the all-reduced slot + the node's upper sites (replicated)  -> value[node]
// A node's value is linear in its units and in the f-terms of its sites, so the three add up.
// The root node also gathers n, sum and sum of squares of its samples per level (the error bar).
// nl > 32 (device only): a whole block shares the node — the root of a finely cut tree has ten thousand
// children, and a single warp walking them through dependent loads was 0.4 ms of a 1.2 ms solve; `xwarp`
// is then the block's shared staging area (kCmbStage doubles per warp) for the fixed-order sum over its warps.
constexpr int kCmbStage = kMaxL + 1 + 3 * (kMaxL + 1);
constexpr int kBigNode = 4096;   // units of one node above which several blocks share the sum
constexpr unsigned long long kExpandSplitBelow = 32768; // (sites x solves) of a geometry pass below which its points are spread over lane groups
// STATS = false: a launch that cannot contain the root of a solve that wants an error bar — the 21 doubles of the
// level statistics then never exist (they cost the kernel 42 registers and a stack frame whoever the node is).
template <typename T, bool STATS = true>
__device__ __forceinline__ void combine_item(const MlpParams<T>& P, const int32_t* list, u64 item, int lane, int nl,
                                             unsigned mask = 0xffffffffu, int mode = CMB_ALL, double* xwarp = nullptr)
{
    const int solve = (int)(item % (u64)P.n_solves);
    const int nidx = list[item / (u64)P.n_solves];
    const PlanNode nd = P.nodes[nidx];
    const DevProblem<T>& pb = P.pb;
    const double* part = P.partial + (u64)solve * P.n_units;
    const T* val = P.value + (u64)solve * P.n_nodes;
    const T* feat = P.feat + (u64)solve * P.n_points;
    double* xb = P.xbuf ? P.xbuf + (u64)solve * P.xstride : nullptr;
    const int slot = P.node_slot ? P.node_slot[nidx] : -1;
    const bool leaf = nd.unit_end - nd.unit_begin == 1 && nd.site_end == nd.site_begin &&
                      (P.unit_tag[nd.unit_begin] >> 4) == UNIT_WHOLE;
    const bool stats = STATS && nidx == 0 && P.lvstat != nullptr;
    double* lv = stats ? P.lvstat + (u64)solve * (3 * (kMaxL + 1)) : nullptr;
    const double* rsq = stats ? P.rootsq + (u64)solve * P.rootsq_stride : nullptr;
    T result;
    double Ns[kMaxL + 1], Sm[kMaxL + 1], Sq[kMaxL + 1];
#pragma unroll
    for (int l = 0; l <= kMaxL; ++l) { Ns[l] = 0.0; Sm[l] = 0.0; Sq[l] = 0.0; }
    if (leaf) {
        // a whole-call leaf in the crown (below an upper site, or the root of a solve that is one unit): in a
        // sharded solve its owner contributes the value, the scatter pass hands the sum back through partial[]
        const bool have = mode == CMB_ALL || (mode == CMB_PARTIAL && P.unit_mine[nd.unit_begin]);
        result = (mode == CMB_FINAL || have) ? (T)part[nd.unit_begin] : (T)0;
        if (stats && have && lane == 0) { // the whole solve was one unit: it reported a pair per level
            for (int l = 0; l <= P.L; ++l) {
                Ns[l] = (double)P.Mpow[l < P.L ? P.L - l : P.L];
                Sm[l] = rsq[2 * l];
                Sq[l] = rsq[2 * l + 1];
            }
        }
    } else {
        const T s = nd.time_site < 0 ? pb.t0 : P.site_r[(u64)solve * P.n_sites + nd.time_site];
        T Bl[kMaxL];
#pragma unroll
        for (int l = 0; l < kMaxL; ++l) Bl[l] = (T)0;
        T Gs = (T)0;
        // (a list with per-node shares: a node without jobs is light and is walked here like any other)
        const int k = (int)(item / (u64)P.n_solves);
        const bool by_jobs = mode != CMB_FINAL && P.stage != nullptr && P.stage_off != nullptr && P.stage_off[k + 1] > P.stage_off[k];
        if (by_jobs || (mode != CMB_FINAL && P.stage != nullptr && P.stage_off == nullptr && nd.unit_end - nd.unit_begin > kBigNode)) {
            // the units (by_jobs: and the sites) were summed by mlp_prereduce_kernel: one record per block
            const int nb = by_jobs ? P.stage_off[k + 1] - P.stage_off[k] : P.stage_nb;
            const u64 rec0 = by_jobs ? (u64)solve * (u64)P.stage_jobs + (u64)P.stage_off[k] : item * (u64)P.stage_nb;
            for (int b = lane; b < nb; b += nl) {
                const double* q = P.stage + (rec0 + (u64)b) * kCmbStage;
#pragma unroll
                for (int l = 0; l < kMaxL; ++l) Bl[l] += (T)q[l];
                Gs += (T)q[kMaxL];
                if (stats) {
#pragma unroll
                    for (int l = 0; l <= kMaxL; ++l) {
                        Ns[l] += q[kMaxL + 1 + 3 * l]; Sm[l] += q[kMaxL + 2 + 3 * l]; Sq[l] += q[kMaxL + 3 + 3 * l];
                    }
                }
            }
        } else if (mode != CMB_FINAL) {
            // four units of this lane per pass, their loads in flight together (a node of a finely cut tree has
            // thousands of units, and a pass of the combine chain lasts as long as its longest lane: walking them one
            // load round-trip at a time was 160 us of a 7.3 ms solve); each lane still adds in ascending unit order
            constexpr int UB = 4;
            for (int u0 = nd.unit_begin + lane; u0 < nd.unit_end; u0 += UB * nl) {
                int tg[UB];
                T vv[UB];
#pragma unroll
                for (int q = 0; q < UB; ++q) {
                    const int u = u0 + q * nl;
                    const bool ok = u < nd.unit_end && !(mode == CMB_PARTIAL && !P.unit_mine[u]);
                    tg[q] = ok ? (int)P.unit_tag[u] : -1;
                    vv[q] = ok ? (T)part[u] : (T)0;
                }
#pragma unroll
                for (int q = 0; q < UB; ++q) {
                    if (tg[q] < 0) continue;
                    const int u = u0 + q * nl, tag = tg[q];
                    const T v = vv[q];
                    const int sl = (tag >> 4) == UNIT_G ? P.L : (tag & 15);
                    if ((tag >> 4) == UNIT_G) Gs += v;
                    else {
#pragma unroll
                        for (int l = 0; l < kMaxL; ++l)
                            if (l == (tag & 15)) Bl[l] += v;
                    }
                    if (stats) {
                        const PlanUnit un = P.units[u];
#pragma unroll
                        for (int l = 0; l <= kMaxL; ++l)
                            if (l == sl) { Ns[l] += (double)(un.k1 - un.k0); Sm[l] += rsq[2 * u]; Sq[l] += rsq[2 * u + 1]; }
                    }
                }
            }
        }
        for (int si = nd.site_begin + lane; si < nd.site_end && !by_jobs; si += nl) {
            if (mode != CMB_ALL) {
                const int cls = P.site_cls[si];
                if (mode == CMB_PARTIAL ? cls != SITE_MINE : cls != SITE_UPPER) continue;
            }
            int l;
            const T add = site_term<T>(P, si, val, part, feat, &l);
#pragma unroll
            for (int q = 0; q < kMaxL; ++q)
                if (q == l) {
                    Bl[q] += add;
                    if (stats) { Ns[q] += 1.0; Sm[q] += (double)add; Sq[q] += (double)add * (double)add; }
                }
        }
        if (nl > 1) { // butterfly over the lanes of this node inside a warp (an aligned group of min(nl, 32) lanes)
            for (int o = (nl > 32 ? 32 : nl) >> 1; o > 0; o >>= 1) {
#pragma unroll
                for (int l = 0; l < kMaxL; ++l) Bl[l] += __shfl_xor_sync(mask, Bl[l], o);
                Gs += __shfl_xor_sync(mask, Gs, o);
                if (stats) {
#pragma unroll
                    for (int l = 0; l <= kMaxL; ++l) {
                        Ns[l] += __shfl_xor_sync(mask, Ns[l], o);
                        Sm[l] += __shfl_xor_sync(mask, Sm[l], o);
                        Sq[l] += __shfl_xor_sync(mask, Sq[l], o);
                    }
                }
            }
        }
#ifndef HDP_HOST_EMU
        if (nl > 32) { // then over the warps of the block, in warp order
            const int w = lane >> 5, nw = nl >> 5;
            if ((lane & 31) == 0) {
                double* q = xwarp + w * kCmbStage;
                for (int l = 0; l < kMaxL; ++l) q[l] = (double)Bl[l];
                q[kMaxL] = (double)Gs;
                for (int l = 0; l <= kMaxL; ++l) {
                    q[kMaxL + 1 + 3 * l] = Ns[l]; q[kMaxL + 2 + 3 * l] = Sm[l]; q[kMaxL + 3 + 3 * l] = Sq[l];
                }
            }
            __syncthreads();
            if (lane == 0) {
                for (int ww = 1; ww < nw; ++ww) {
                    const double* q = xwarp + ww * kCmbStage;
                    for (int l = 0; l < kMaxL; ++l) Bl[l] += (T)q[l];
                    Gs += (T)q[kMaxL];
                    for (int l = 0; l <= kMaxL; ++l) {
                        Ns[l] += q[kMaxL + 1 + 3 * l]; Sm[l] += q[kMaxL + 2 + 3 * l]; Sq[l] += q[kMaxL + 3 + 3 * l];
                    }
                }
            }
        }
#endif
        T a = (T)0;
#pragma unroll
        for (int l = 0; l < kMaxL; ++l)
            if (l < nd.lv) a += ((pb.t1 - s) * Bl[l]) / (T)P.Mpow[nd.lv - l];
        const T a2 = Gs / (T)P.Mpow[nd.lv];
        result = a + a2;
    }
    if (lane != 0) return;
    if (mode == CMB_PARTIAL) xb[slot] = (double)result;
    else P.value[(u64)solve * P.n_nodes + nidx] = (mode == CMB_FINAL && !leaf) ? (T)(xb[slot] + (double)result) : result;
    if (stats) {
        double* xs = xb ? xb + P.n_slots : nullptr; // the root statistics travel behind the crown slots
        if (mode == CMB_PARTIAL) {
            for (int l = 0; l <= kMaxL; ++l) { xs[3 * l] = Ns[l]; xs[3 * l + 1] = Sm[l]; xs[3 * l + 2] = Sq[l]; }
            return;
        }
        // standard error of the root value: the samples of a level are i.i.d., the levels independent
        double var = 0.0;
        for (int l = 0; l <= P.L; ++l) {
            if (mode == CMB_FINAL) { Ns[l] += xs[3 * l]; Sm[l] += xs[3 * l + 1]; Sq[l] += xs[3 * l + 2]; }
            lv[3 * l] = Ns[l]; lv[3 * l + 1] = Sm[l]; lv[3 * l + 2] = Sq[l];
            const double n = Ns[l];
            const double w = l < P.L ? (double)(pb.t1 - pb.t0) : 1.0;
            if (n > 1.5) {
                double v1 = (Sq[l] - Sm[l] * Sm[l] / n) / (n - 1.0); // sample variance of the level
                if (v1 < 0.0) v1 = 0.0;
                var += w * w * v1 / n;
            }
        }
        if (P.std_err) P.std_err[solve] = (T)sqrt(var);
    }
}

// One thread's part (`tid` of `nthreads`) of share `share` of `nshare` of a node's pre-reduction: its units and — with
// `with_sites`, the crown partial pass of a sharded solve — its sites, summed per level into Bl / Gs (and, for the root
// of a solve that wants an error bar, the level statistics).  Shared by mlp_prereduce_kernel (a block's threads, then
// reduced over the block) and the test-only host emulation (one thread).
template <typename T, bool STATS>
__device__ __forceinline__ void prereduce_share(const MlpParams<T>& P, const PlanNode& nd, int nidx, int solve, int share, int nshare,
                                                bool with_sites, int mode, int tid, int nthreads, T* Bl, T& Gs, double* Ns,
                                                double* Sm, double* Sq)
{
    const int nu = nd.unit_end - nd.unit_begin;
    const double* part = P.partial + (u64)solve * P.n_units;
    const bool stats = STATS && nidx == 0 && P.lvstat != nullptr;
    const double* rsq = stats ? P.rootsq + (u64)solve * P.rootsq_stride : nullptr;
    const int per = (nu + nshare - 1) / nshare;
    const int u0 = nd.unit_begin + share * per;
    const int u1 = u0 + per < nd.unit_end ? u0 + per : nd.unit_end;
    for (int u = u0 + tid; u < u1; u += nthreads) {
        if (mode == CMB_PARTIAL && !P.unit_mine[u]) continue;
        const int tag = P.unit_tag[u];
        const T v = (T)part[u];
        const int sl = (tag >> 4) == UNIT_G ? P.L : (tag & 15);
        if ((tag >> 4) == UNIT_G) Gs += v;
        else {
#pragma unroll
            for (int l = 0; l < kMaxL; ++l)
                if (l == (tag & 15)) Bl[l] += v;
        }
        if (stats) {
            const PlanUnit un = P.units[u];
#pragma unroll
            for (int l = 0; l <= kMaxL; ++l)
                if (l == sl) { Ns[l] += (double)(un.k1 - un.k0); Sm[l] += rsq[2 * u]; Sq[l] += rsq[2 * u + 1]; }
        }
    }
    if (with_sites) { // this share of the node's sites
        const T* val = P.value + (u64)solve * P.n_nodes;
        const T* feat = P.feat + (u64)solve * P.n_points;
        const int ns = nd.site_end - nd.site_begin, sper = (ns + nshare - 1) / nshare;
        const int s0 = nd.site_begin + share * sper;
        const int s1 = s0 + sper < nd.site_end ? s0 + sper : nd.site_end;
        for (int si = s0 + tid; si < s1; si += nthreads) {
            if (mode != CMB_ALL) {
                const int cls = P.site_cls[si];
                if (mode == CMB_PARTIAL ? cls != SITE_MINE : cls != SITE_UPPER) continue;
            }
            int l;
            const T add = site_term<T>(P, si, val, part, feat, &l);
#pragma unroll
            for (int q = 0; q < kMaxL; ++q)
                if (q == l) {
                    Bl[q] += add;
                    if (stats) { Ns[q] += 1.0; Sm[q] += (double)add; Sq[q] += (double)add * (double)add; }
                }
        }
    }
}

// points[solve][0] <- x0 (padded with zeros)
template <typename T>
__device__ __forceinline__ void init_item(const MlpParams<T>& P, const T* x0, u64 q)
{
    const u64 solve = q / (u64)P.dpad;
    const int i = (int)(q % (u64)P.dpad);
    const u64 src = (P.n_x0 == 1 ? 0 : solve) * (u64)P.pb.d;
    const T v = i < P.pb.d ? x0[src + i] : (T)0;
    P.points[solve * (u64)P.n_points * (u64)P.dpad + i] = v;
    if (P.pb.has_bc && i < P.pb.d) { // every solve steps from x0 (src/reflect.jl:12)
        const T lo = P.pb.bc_is_scalar ? P.pb.lo_s : P.pb.lo[i], hi = P.pb.bc_is_scalar ? P.pb.hi_s : P.pb.hi[i];
        if (!(v >= lo && v <= hi)) atomicExch(&P.flags[0], 1);
    }
}

#ifndef HDP_HOST_EMU
// ---------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------
extern __shared__ double2 hdp_dyn_smem[]; // units kernel: blockDim.x * sizeof(Frame<T>) bytes

// A solve is a chain of ten to twenty dependent launches (geometry by depth, units, combine by depth) and a
// single solve — sharded or not — lasts 0.3-7 ms, so the gaps between the links count.  Every kernel of the
// chain is launched with programmatic stream serialisation (k_mlp_host.cu: launch_chained) and opens with this
// pair: `launch_dependents` lets the next link's blocks be scheduled onto whatever SM capacity this grid leaves
// free (its launch latency, parameter load and — fp64 — the staging of the generator tables then overlap this
// grid); `wait` blocks until every grid before this one has completed and its writes are visible.  All threads
// of all blocks execute the pair first thing, unconditionally: a grid that exits without `wait` would be taken as
// complete before its predecessors are.  Launched without the attribute (NVRTC-built kernels, host emulation)
// both instructions do nothing.
__device__ __forceinline__ void chain_enter()
{
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename T>
__global__ void mlp_init_kernel(const __grid_constant__ MlpParams<T> P, const T* __restrict__ x0)
{
    chain_enter();
    const u64 total = (u64)P.n_solves * (u64)P.dpad;
    for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (u64)gridDim.x * blockDim.x)
        init_item<T>(P, x0, q);
}

// (two lists per launch: a sharded solve expands the replicated upper sites of a depth and its own sites of
// that depth together — they do not depend on each other, and every launch of the chain costs 10-20 us)
template <typename T, int G>
__global__ void __launch_bounds__(128) mlp_expand_kernel(const __grid_constant__ MlpParams<T> P, const int32_t* list, int count,
                                                         int count_hits, const int32_t* list2, int count2, int count_hits2)
{
    asm volatile("griddepcontrol.launch_dependents;");
    if constexpr (sizeof(T) == 8) tables_init(); // (constant tables: staged while the previous link still runs)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    Ctx<T, G> c;
    ctx_init<T, G>(c, P, (u64)blockIdx.x * blockDim.x + threadIdx.x, threadIdx.x & 31);
    // a short pass (a link of the launch chain that cannot fill the GPU) takes one lane group per (site, point)
    const bool split = (u64)(count + count2) * (u64)P.n_solves < kExpandSplitBelow;
    const u64 per = (u64)P.n_solves * (u64)(split ? 1 + P.Kp : 1);
    const u64 n1 = (u64)count * per, n12 = n1 + (u64)count2 * per;
    // grid-stride: a large pass runs on a resident grid, so that staging the generator tables (6 KB per block)
    // is paid once per resident block and not once per 128 items
    const u64 stride = (u64)gridDim.x * blockDim.x / G;
    for (u64 gid = c.tid / G; gid < n12; gid += stride) {
        c.nb = 0;
        if (gid < n1) expand_item<T, G>(c, P, list, gid, count_hits != 0, split);
        else expand_item<T, G>(c, P, list2, gid - n1, count_hits2 != 0, split);
    }
}

// the work units: dynamic fetch, one warp takes 32/G consecutive work items at a time
// (lane groups of a warp then run equal-shaped units in lockstep)
template <typename T, int G, int U, int FM = -1>
__global__ void __launch_bounds__(HDP_MLP_BLOCK, Tune<U>::MB) mlp_units_kernel(const __grid_constant__ MlpParams<T> P)
{
    constexpr int GPW = 32 / G;
    asm volatile("griddepcontrol.launch_dependents;");
    if constexpr (sizeof(T) == 8) tables_init();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    Ctx<T, G> c;
    const int lane_w = threadIdx.x & 31;
    ctx_init<T, G>(c, P, (u64)blockIdx.x * blockDim.x + threadIdx.x, lane_w);
    if (FM >= 0 && c.fast != FM) { // the host and the device disagree about the mode: never compute with the wrong loop
        if (threadIdx.x == 0) atomicExch(&P.flags[1], 1);
        return;
    }
#ifdef HDP_FRAME_SHARED_EXPERIMENT
    c.frame_mem = reinterpret_cast<Frame<T>*>(hdp_dyn_smem) + threadIdx.x / G;
#else
    c.frame_mem = reinterpret_cast<Frame<T>*>(hdp_dyn_smem) + threadIdx.x; // one private frame per thread
#endif
    const int sub = lane_w / G;
    const u64 n_work = (u64)P.n_exec * (u64)P.n_solves; // this rank's units (all of them on one GPU)
    for (;;) {
        unsigned long long base = 0;
        if (lane_w == 0) base = atomicAdd(P.work_ctr, (unsigned long long)GPW);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_work) break;
        const u64 wl = base + (u64)sub;
        if (wl < n_work) unit_item<T, G, U, FM>(c, P, wl);
    }
    if (c.bad) atomicExch(&P.flags[0], 1);
}

// First stage of the combine pass for nodes with more than kBigNode units: grid (kStageBlocks, items); block b of
// an item sums its share of the node's unit partials (this rank's, when the solve is sharded: mode CMB_PARTIAL)
// per level — and, for the root, the sample statistics of the error bar — in a fixed order, and leaves one
// kCmbStage record.  A single block of 512 lanes walking the 10^5 units of the root of Allen-Cahn d=100, M=8, L=6
// through dependent loads took 0.23 ms of a 1.6 ms sharded solve.
// With `jobs` (crown of a sharded solve, P.stage_off): grid (jobs, solves); job j = (list index k, share b of the
// stage_off[k + 1] - stage_off[k] shares of that node) sums its share of the node's units and of its sites (the
// f-terms of this rank's ownership roots) — the combine launch that follows only adds the records up, so the
// root, whose thousands of sites one block walked through dependent loads in 33 us, costs what any node costs.
template <typename T, bool STATS = true>
__global__ void __launch_bounds__(256, 3) mlp_prereduce_kernel(const __grid_constant__ MlpParams<T> P, const int32_t* list, int mode,
                                                            double* stage, const int2* jobs)
{
    chain_enter();
    int share, nshare, solve, nidx;
    u64 rec;
    if (jobs) {
        const int2 jb = jobs[blockIdx.x];
        share = jb.y;
        nshare = P.stage_off[jb.x + 1] - P.stage_off[jb.x];
        solve = (int)blockIdx.y;
        nidx = list[jb.x];
        rec = (u64)solve * (u64)P.stage_jobs + (u64)blockIdx.x;
    } else {
        const u64 item = blockIdx.y;
        share = (int)blockIdx.x;
        nshare = (int)gridDim.x;
        solve = (int)(item % (u64)P.n_solves);
        nidx = list[item / (u64)P.n_solves];
        rec = item * (u64)gridDim.x + blockIdx.x;
    }
    const PlanNode nd = P.nodes[nidx];
    const int nu = nd.unit_end - nd.unit_begin;
    if (!jobs && nu <= kBigNode) return;
    const bool stats = STATS && nidx == 0 && P.lvstat != nullptr;
    T Bl[kMaxL];
    T Gs = (T)0;
    double Ns[kMaxL + 1], Sm[kMaxL + 1], Sq[kMaxL + 1];
#pragma unroll
    for (int l = 0; l < kMaxL; ++l) Bl[l] = (T)0;
#pragma unroll
    for (int l = 0; l <= kMaxL; ++l) { Ns[l] = 0.0; Sm[l] = 0.0; Sq[l] = 0.0; }
    prereduce_share<T, STATS>(P, nd, nidx, solve, share, nshare, jobs != nullptr, mode, (int)threadIdx.x, (int)blockDim.x, Bl, Gs, Ns,
                              Sm, Sq);
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int l = 0; l < kMaxL; ++l) Bl[l] += __shfl_xor_sync(0xffffffffu, Bl[l], o);
        Gs += __shfl_xor_sync(0xffffffffu, Gs, o);
        if (stats) {
#pragma unroll
            for (int l = 0; l <= kMaxL; ++l) {
                Ns[l] += __shfl_xor_sync(0xffffffffu, Ns[l], o);
                Sm[l] += __shfl_xor_sync(0xffffffffu, Sm[l], o);
                Sq[l] += __shfl_xor_sync(0xffffffffu, Sq[l], o);
            }
        }
    }
    __shared__ double xw[8 * kCmbStage];
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        double* q = xw + w * kCmbStage;
        for (int l = 0; l < kMaxL; ++l) q[l] = (double)Bl[l];
        q[kMaxL] = (double)Gs;
        for (int l = 0; l <= kMaxL; ++l) { q[kMaxL + 1 + 3 * l] = Ns[l]; q[kMaxL + 2 + 3 * l] = Sm[l]; q[kMaxL + 3 + 3 * l] = Sq[l]; }
    }
    __syncthreads();
    if (threadIdx.x < kCmbStage) { // entry by entry, the warps in order
        double a = 0.0;
        bool as_t = threadIdx.x <= kMaxL; // level sums and the g-term are sums of T values
        T at = (T)0;
        for (int ww = 0; ww < (int)(blockDim.x >> 5); ++ww) {
            const double v = xw[ww * kCmbStage + threadIdx.x];
            if (as_t) at += (T)v; else a += v;
        }
        stage[rec * kCmbStage + threadIdx.x] = as_t ? (double)at : a;
    }
}

// nl lanes (a power of two, 1..32) share one (node, solve): 32 where nodes have many children (the
// root of a deeply cut tree has thousands), 4 where they have a handful
template <typename T, bool STATS = true>
__global__ void __launch_bounds__(512) mlp_combine_kernel(const __grid_constant__ MlpParams<T> P, const int32_t* list, int count,
                                                          int nl, int mode)
{
    chain_enter();
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    // the call-wide tail of the exchange buffer: wall hits of this rank out, the all-reduced total back
    if (t == 0 && P.xbuf) {
        double* tail = P.xbuf + (u64)P.n_solves * P.xstride;
        if (mode == CMB_PARTIAL) tail[4] = (double)*P.bounce_ctr;
        if (mode == CMB_FINAL && list[0] == 0) *P.bounce_ctr = (unsigned long long)tail[4];
    }
    __shared__ double xwarp[16 * kCmbStage]; // nl > 32: one node per block of nl <= 512 threads
    const u64 gid = t / (unsigned)nl;
    if (gid >= (u64)count * (u64)P.n_solves) return; // (the nl lanes of a node leave together)
    // the shuffles of a node stay inside its own nl lanes: the root's statistics are reduced by its group alone
    const int lane = (int)(t % (unsigned)nl);
    const unsigned mask = nl >= 32 ? 0xffffffffu : (((1u << nl) - 1u) << ((threadIdx.x & 31) - lane));
    combine_item<T, STATS>(P, list, gid, lane, nl, mask, mode, xwarp);
}

// sharded solve: the all-reduced values of the whole-call leaves below upper sites go back to partial[],
// where the sites read their children
template <typename T>
__global__ void __launch_bounds__(128) mlp_scatter_kernel(const __grid_constant__ MlpParams<T> P, const int32_t* list, int count)
{
    chain_enter();
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (u64)count * (u64)P.n_solves) return;
    const int solve = (int)(t % (u64)P.n_solves);
    const int nidx = list[t / (u64)P.n_solves];
    P.partial[(u64)solve * P.n_units + P.nodes[nidx].unit_begin] = P.xbuf[(u64)solve * P.xstride + P.node_slot[nidx]];
}

// ---------------------------------------------------------------------
// The exchange of a sharded solve over peer memory (hdp_peer_window_*, include/hdpde.h): push — flag — sum —
// scatter in ONE single-block launch, in place of ncclAllReduce + mlp_scatter_kernel.
// Window of a rank: double data[2][kPeerMax][kPeerSlot], then u64 flag[2][kPeerMax]; slot [p][r] and flag [p][r]
// are written by rank r alone, in the exchanges of parity p.  Rank r's flag is the number of its latest exchange
// (they only grow), stored with release semantics after its doubles; a reader takes it with acquire semantics
// and then reads the slot past L1.  Two parities suffice: a rank can only be one exchange ahead of a peer (it
// cannot finish exchange n + 1 without the peer's flag n + 1, which the peer publishes after it has finished
// reading exchange n), and an exchange ahead writes the other parity.
// The sum runs over the slots in rank order on every rank: identical bits everywhere, no dependence on arrival.
// ---------------------------------------------------------------------
constexpr int kPeerMax = 16;       // == HDP_PEER_MAX_RANKS
constexpr int kPeerSlot = 8192;    // == HDP_PEER_SLOT_DOUBLES
constexpr unsigned long long kPeerTimeoutNs = 20ull * 1000000000ull;
constexpr size_t kPeerWindowBytes = (size_t)2 * kPeerMax * kPeerSlot * sizeof(double) + (size_t)2 * kPeerMax * sizeof(unsigned long long);
struct PeerWin {
    double* base[kPeerMax]; // every rank's window as mapped on this device
    int nranks, rank;
    unsigned long long epoch;
};
__device__ __forceinline__ unsigned long long* peer_flag(double* base, int par, int r)
{
    return reinterpret_cast<unsigned long long*>(base + (size_t)2 * kPeerMax * kPeerSlot) + par * kPeerMax + r;
}
template <typename T>
__global__ void __launch_bounds__(1024) mlp_exchange_kernel(const __grid_constant__ MlpParams<T> P, const __grid_constant__ PeerWin W,
                                                            int n, const int32_t* leaves, int n_leaves)
{
    chain_enter();
    const int par = (int)(W.epoch & 1ull);
    const size_t slot = ((size_t)par * kPeerMax + (size_t)W.rank) * kPeerSlot;
    // push: this rank's doubles into its slot of every window (the peers' over NVLink, its own locally)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = P.xbuf[i];
        for (int q = 0; q < W.nranks; ++q) {
            int r = W.rank + q; // (each rank starts with a different peer)
            if (r >= W.nranks) r -= W.nranks;
            W.base[r][slot + i] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < W.nranks) {
        const int r = threadIdx.x;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flag(W.base[r], par, W.rank)), "l"(W.epoch) : "memory");
        const unsigned long long* f = peer_flag(W.base[W.rank], par, r);
        unsigned long long v, t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= W.epoch) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > kPeerTimeoutNs) { // the peer never arrived: fail the call, do not hang the GPU
                atomicExch(&P.flags[2], 1);
                break;
            }
        }
    }
    __syncthreads();
    // sum of the slots of this rank's window, in rank order
    const double* mine = W.base[W.rank] + (size_t)par * kPeerMax * kPeerSlot;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double a = 0.0;
        for (int r = 0; r < W.nranks; ++r) a += __ldcg(mine + (size_t)r * kPeerSlot + i);
        P.xbuf[i] = a;
    }
    __syncthreads();
    // whole-call leaves below upper sites: their summed values go back to partial[], where the sites read them
    const u64 tot = (u64)n_leaves * (u64)P.n_solves;
    for (u64 t = threadIdx.x; t < tot; t += blockDim.x) {
        const int solve = (int)(t % (u64)P.n_solves);
        const int nidx = leaves[t / (u64)P.n_solves];
        P.partial[(u64)solve * P.n_units + P.nodes[nidx].unit_begin] = P.xbuf[(u64)solve * P.xstride + P.node_slot[nidx]];
    }
}
#endif // HDP_HOST_EMU

} // namespace hdp
