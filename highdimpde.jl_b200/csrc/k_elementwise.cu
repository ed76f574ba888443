// k_elementwise.cu — K1 (reflected Euler–Maruyama loop), K2 (samplers),
// K4 (reflection operator).  One thread owns VEC consecutive elements of the
// column-major d×B state (one 128-bit load/store, one Philox call) and keeps
// them in registers across all N steps; nothing but y0, y1 (and z) is written.
//
// Replaces: sde_loop! (src/DeepSplitting.jl:153-166), _reflect
// (src/reflect.jl:8-82), UniformSampling/NormalSampling call operators
// (src/MCSample.jl:28-33, 63-71).
#include <string.h>
#include "hdp_host_util.h"
#include "k_sde.cuh"
#include "user_kinds.h"

using namespace hdp;

namespace {

template <typename T> struct Bounds {
    int scalar;
    T lo_s, hi_s;
    const T* lo;
    const T* hi;
    __device__ __forceinline__ T lo_at(int i) const { return scalar ? lo_s : lo[i]; }
    __device__ __forceinline__ T hi_at(int i) const { return scalar ? hi_s : hi[i]; }
};

// ------------------------------ K4 -----------------------------------
template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(256)
reflect_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out,
               int32_t* __restrict__ cnt, size_t n_el, int d, Bounds<T> bd, int* flags)
{
    constexpr int VEC = Tr<T>::VEC;
    const size_t nv = (n_el + VEC - 1) / VEC;
    bool bad = false;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (size_t)gridDim.x * blockDim.x) {
        const size_t q0 = v * VEC;
        Vec<T> av = load_el<T, ALIGNED>(a, q0, n_el);
        Vec<T> bv = load_el<T, ALIGNED>(b, q0, n_el);
        int i = (int)(q0 % (size_t)d);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            if (q0 + e < n_el) {
                T lo = bd.lo_at(i), hi = bd.hi_at(i);
                if (!(av.v[e] >= lo && av.v[e] <= hi)) bad = true; // src/reflect.jl:12, :59
                int nb = 0;
                bv.v[e] = reflect_coord<T>(bv.v[e], lo, hi, nb);
                if (cnt) cnt[q0 + e] = nb;
            }
            if (++i == d) i = 0;
        }
        store_el<T, ALIGNED>(out, q0, n_el, bv);
    }
    if (bad) atomicExch(&flags[0], 1);
}

// ------------------------------ K2 -----------------------------------
template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(256)
mc_sample_kernel(T* __restrict__ z, const T* __restrict__ y1, const T* __restrict__ inject,
                 size_t n_el, int d, int K, DevSampler<T> sm, u64 seed, int stream)
{
    constexpr int VEC = Tr<T>::VEC;
    if constexpr (sizeof(T) == 8) tables_init();
    const size_t nv = (n_el + VEC - 1) / VEC;
    const bool normal = sm.kind == HDP_SAMPLER_NORMAL;
    const bool use_ref = normal && sm.shifted && y1 != nullptr;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (size_t)gridDim.x * blockDim.x) {
        const size_t q0 = v * VEC;
        Vec<T> ref;
#pragma unroll
        for (int e = 0; e < VEC; ++e) ref.v[e] = (T)0;
        if (use_ref) ref = load_el<T, ALIGNED>(y1, q0, n_el);
        const int i0 = (int)(q0 % (size_t)d);
        for (int k = 0; k < K; ++k) {
            Vec<T> raw = inject ? load_el<T, ALIGNED>(inject + (size_t)k * n_el, q0, n_el)
                                : draw_vec<T>(seed, stream, (u64)k, (u64)v, normal);
            int i = i0;
            Vec<T> o;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                o.v[e] = (use_ref || !normal) ? sample_one<T>(sm, raw.v[e], i, ref.v[e])
                                              : normal_map<T>(raw.v[e], sm.sigma, false, (T)0);
                if (++i == d) i = 0;
            }
            store_el<T, ALIGNED>(z + (size_t)k * n_el, q0, n_el, o);
        }
    }
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

inline int grid_for(hdp_context* h, size_t nv, int threads, int per_sm)
{
    size_t blocks = (nv + threads - 1) / threads;
    size_t cap = (size_t)h->prop.multiProcessorCount * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// stage a caller array on the device (HOST flag) or pass it through (DEVICE flag)
template <typename T>
int stage_in(hdp_context* h, const T* src, size_t n, unsigned mem_flags, int which, const T** dptr)
{
    if (!src) { *dptr = nullptr; return HDP_OK; }
    if (mem_flags & HDP_MEM_DEVICE) { *dptr = src; return HDP_OK; }
    T* d = (T*)ensure(h, which, n * sizeof(T));
    if (!d) return HDP_ERR_ALLOC;
    HDP_CUDA(h, cudaMemcpyAsync(d, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    *dptr = d;
    return HDP_OK;
}
template <typename T>
int stage_out(hdp_context* h, T* dst, size_t n, unsigned mem_flags, int which, T** dptr)
{
    if (!dst) { *dptr = nullptr; return HDP_OK; }
    if (mem_flags & HDP_MEM_DEVICE) { *dptr = dst; return HDP_OK; }
    T* d = (T*)ensure(h, which, n * sizeof(T));
    if (!d) return HDP_ERR_ALLOC;
    *dptr = d;
    return HDP_OK;
}
template <typename T>
int copy_back(hdp_context* h, T* dst, const T* d, size_t n, unsigned mem_flags)
{
    if (!dst || (mem_flags & HDP_MEM_DEVICE)) return HDP_OK;
    HDP_CUDA(h, cudaMemcpyAsync(dst, d, n * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    return HDP_OK;
}

int check_flag(hdp_context* h, const char* what)
{
    int f = 0;
    HDP_CUDA(h, cudaMemcpyAsync(&f, h->d_flags, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    HDP_CUDA(h, cudaStreamSynchronize(h->stream));
    if (f) {
        cudaMemsetAsync(h->d_flags, 0, sizeof(int), h->stream);
        return set_err(h, HDP_ERR_START_OUTSIDE_DOMAIN, "%s: start point not in hypercube", what);
    }
    return HDP_OK;
}

// ------------------------------ K4 host ------------------------------
template <typename T>
int reflect_impl(hdp_handle h, const T* a, const T* b, T* out, int d, int B, const T* lo, const T* hi,
                 int bounds_are_scalar, int32_t* cnt, unsigned mem_flags)
{
    if (!h || !a || !b || !out || !lo || !hi || d <= 0 || B < 0) return HDP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    const size_t n_el = (size_t)d * (size_t)B;
    if (n_el == 0) return HDP_OK;
    const T *da, *db, *dlo = nullptr, *dhi = nullptr;
    T* dout;
    int32_t* dcnt = nullptr;
    int rc;
    if ((rc = stage_in(h, a, n_el, mem_flags, hdp_context::B_A, &da)) != HDP_OK) return rc;
    if ((rc = stage_in(h, b, n_el, mem_flags, hdp_context::B_B, &db)) != HDP_OK) return rc;
    if ((rc = stage_out(h, out, n_el, mem_flags, hdp_context::B_C, &dout)) != HDP_OK) return rc;
    if ((rc = stage_out(h, cnt, n_el, mem_flags, hdp_context::B_D, &dcnt)) != HDP_OK) return rc;
    Bounds<T> bd{};
    bd.scalar = bounds_are_scalar;
    if (bounds_are_scalar) {
        T l, u;
        if (mem_flags & HDP_MEM_DEVICE) {
            HDP_CUDA(h, cudaMemcpyAsync(&l, lo, sizeof(T), cudaMemcpyDeviceToHost, h->stream));
            HDP_CUDA(h, cudaMemcpyAsync(&u, hi, sizeof(T), cudaMemcpyDeviceToHost, h->stream));
            HDP_CUDA(h, cudaStreamSynchronize(h->stream));
        } else { l = lo[0]; u = hi[0]; }
        bd.lo_s = l; bd.hi_s = u;
    } else {
        if ((rc = stage_in(h, lo, (size_t)d, mem_flags, hdp_context::B_E, &dlo)) != HDP_OK) return rc;
        if ((rc = stage_in(h, hi, (size_t)d, mem_flags, hdp_context::B_F, &dhi)) != HDP_OK) return rc;
        bd.lo = dlo; bd.hi = dhi;
    }
    constexpr int VEC = Tr<T>::VEC;
    const size_t nv = (n_el + VEC - 1) / VEC;
    const bool al = (n_el % VEC == 0) && aligned16(da) && aligned16(db) && aligned16(dout);
    const int grid = grid_for(h, nv, 256, 16);
    HDP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    if (al) reflect_kernel<T, true><<<grid, 256, 0, h->stream>>>(da, db, dout, dcnt, n_el, d, bd, h->d_flags);
    else reflect_kernel<T, false><<<grid, 256, 0, h->stream>>>(da, db, dout, dcnt, n_el, d, bd, h->d_flags);
    HDP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    h->launches++;
    HDP_CUDA(h, cudaGetLastError());
    if ((rc = copy_back(h, out, dout, n_el, mem_flags)) != HDP_OK) return rc;
    if ((rc = copy_back(h, cnt, dcnt, n_el, mem_flags)) != HDP_OK) return rc;
    return check_flag(h, "_reflect");
}

// ------------------------------ K2 host ------------------------------
template <typename T>
int mc_sample_impl(hdp_handle h, T* z, const T* y1, int d, int B, int K, const hdp_sampler* sampler,
                   uint64_t seed, const T* inject, unsigned mem_flags)
{
    if (!h || !z || !sampler || d <= 0 || B < 0 || K < 0) return HDP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    if (sampler->kind == HDP_SAMPLER_NONE) return HDP_OK; // NoSampling()(x...) = nothing (src/MCSample.jl:86)
    const size_t n_el = (size_t)d * (size_t)B;
    if (n_el == 0 || K == 0) return HDP_OK;
    DevSampler<T> sm;
    int rc = make_dev_sampler<T>(h, sampler, d, hdp_context::B_E, &sm);
    if (rc != HDP_OK) return rc;
    const T *dy1 = nullptr, *dinj = nullptr;
    T* dz;
    if (sampler->kind == HDP_SAMPLER_NORMAL && sampler->shifted && y1)
        if ((rc = stage_in(h, y1, n_el, mem_flags, hdp_context::B_A, &dy1)) != HDP_OK) return rc;
    if ((rc = stage_in(h, inject, n_el * (size_t)K, mem_flags, hdp_context::B_B, &dinj)) != HDP_OK) return rc;
    if ((rc = stage_out(h, z, n_el * (size_t)K, mem_flags, hdp_context::B_C, &dz)) != HDP_OK) return rc;
    constexpr int VEC = Tr<T>::VEC;
    const size_t nv = (n_el + VEC - 1) / VEC;
    if (!draw_address_ok(HDP_STREAM_MC, (u64)K, (u64)nv))
        return set_err(h, HDP_ERR_BAD_ARG, "mc_sample!: more than 2^24 samples or 2^37 vec-groups");
    const bool al = (n_el % VEC == 0) && aligned16(dz) && aligned16(dy1) && aligned16(dinj);
    const int grid = grid_for(h, nv, 256, 16);
    HDP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    if (al) mc_sample_kernel<T, true><<<grid, 256, 0, h->stream>>>(dz, dy1, dinj, n_el, d, K, sm, seed, HDP_STREAM_MC);
    else mc_sample_kernel<T, false><<<grid, 256, 0, h->stream>>>(dz, dy1, dinj, n_el, d, K, sm, seed, HDP_STREAM_MC);
    HDP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    h->launches++;
    HDP_CUDA(h, cudaGetLastError());
    if ((rc = copy_back(h, z, dz, n_el * (size_t)K, mem_flags)) != HDP_OK) return rc;
    if (!(mem_flags & HDP_MEM_DEVICE)) HDP_CUDA(h, cudaStreamSynchronize(h->stream));
    return HDP_OK;
}

// ------------------------------ K1 host ------------------------------
inline bool uses_user_coef(const hdp_problem* p) { return p->mu_kind == HDP_COEF_USER || p->sigma_kind == HDP_COEF_USER; }

// NVRTC-built K1 kernels for the expressions of hdp_set_user_exprs (once per handle and expression pair)
int get_sde_user_module(hdp_context* h, const hdp_problem* prob, bool f32, SdeUserModule** out)
{
    hdp_user_exprs u{};
    u.mu = h->has_user_mu ? h->user_mu.c_str() : nullptr;
    u.sigma = h->has_user_sigma ? h->user_sigma.c_str() : nullptr;
    // the module is specialised to the pair of kinds as well as to the expressions
    std::string key = (f32 ? "f32/" : "f64/") + std::to_string(prob->mu_kind) + "/" + std::to_string(prob->sigma_kind) + "/";
    key += (prob->mu_kind == HDP_COEF_USER && u.mu) ? u.mu : "-";
    key += "\x1f";
    key += (prob->sigma_kind == HDP_COEF_USER && u.sigma) ? u.sigma : "-";
    auto it = h->sde_user_modules.find(key);
    if (it != h->sde_user_modules.end()) { *out = &it->second; return HDP_OK; }
    SdeUserBuild b;
    int rc = user_compile_sde(prob, &u, f32, &b);
    if (rc != HDP_OK) return set_err(h, rc, "%s", b.log.c_str());
    SdeUserModule m;
    HDP_CUDA(h, cudaLibraryLoadData(&m.lib, b.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    int n = 0;
    for (int al = 0; al < 2; ++al)
        for (int inj = 0; inj < 2; ++inj) HDP_CUDA(h, cudaLibraryGetKernel(&m.sde[al][inj], m.lib, b.names[n++].c_str()));
    for (int al = 0; al < 2; ++al)
        for (int inj = 0; inj < 2; ++inj)
            for (int tr = 0; tr < 2; ++tr) HDP_CUDA(h, cudaLibraryGetKernel(&m.ens[al][inj][tr], m.lib, b.names[n++].c_str()));
    auto ins = h->sde_user_modules.emplace(key, m);
    *out = &ins.first->second;
    return HDP_OK;
}

template <typename T>
int sde_loop_impl(hdp_handle h, const hdp_problem* prob, T* y0, T* y1, int B, int nsteps, T dt, const T* ts,
                  const hdp_sampler* x0_sampler, uint64_t seed, const T* dWall_inject,
                  const hdp_sampler* mc, T* z, int K, int normal_mode, unsigned mem_flags, int mlp_assoc = 0, int n_epochs = 1)
{
    if (!h || !prob || !y0 || !y1 || B < 0 || nsteps < 0 || n_epochs < 1 || n_epochs > 65535) return HDP_ERR_BAD_ARG;
    if (normal_mode != HDP_NORMAL_DEFAULT && normal_mode != (sizeof(T) == 8 ? HDP_NORMAL_BM64 : HDP_NORMAL_BM32))
        return set_err(h, HDP_ERR_UNSUPPORTED, "normal_mode %d not supported for this element type", normal_mode);
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    DevProblem<T> pb;
    int rc = make_dev_problem<T>(h, prob, hdp_context::B_E, &pb, /*allow_user=*/true);
    if (rc != HDP_OK) return rc;
    const int d = prob->d;
    const size_t n_el = (size_t)d * (size_t)B; // per epoch
    const size_t E = (size_t)n_epochs;
    if (n_el == 0) return HDP_OK;
    // user μ/σ: the enumerated kinds are time-independent, expressions may read t = ts[step]
    SdeUserModule* um = nullptr;
    const T* dts = nullptr;
    if (uses_user_coef(prob)) {
        if (!ts && nsteps > 0) return set_err(h, HDP_ERR_BAD_ARG, "sde_loop!: user-defined mu/sigma need ts (t = ts[N+1-i])");
        if ((rc = get_sde_user_module(h, prob, sizeof(T) == 4, &um)) != HDP_OK) return rc;
        if (nsteps > 0) { // ts is a HOST array whatever mem_flags says (a Julia range, collected)
            T* t = (T*)ensure(h, hdp_context::B_I, (size_t)nsteps * sizeof(T));
            if (!t) return HDP_ERR_ALLOC;
            HDP_CUDA(h, cudaMemcpyAsync(t, ts, (size_t)nsteps * sizeof(T), cudaMemcpyHostToDevice, h->stream));
            dts = t;
        }
    }
    DevSampler<T> x0s, mcs;
    if ((rc = make_dev_sampler<T>(h, x0_sampler, d, hdp_context::B_F, &x0s)) != HDP_OK) return rc;
    if ((rc = make_dev_sampler<T>(h, mc, d, hdp_context::B_G, &mcs)) != HDP_OK) return rc;
    const bool want_z = z != nullptr && mcs.kind != HDP_SAMPLER_NONE && K > 0;
    T *dy0, *dy1, *dz = nullptr;
    const T* ddw = nullptr;
    if (mem_flags & HDP_MEM_DEVICE) { dy0 = y0; dy1 = y1; dz = z; ddw = dWall_inject; }
    else {
        dy0 = (T*)ensure(h, hdp_context::B_A, E * n_el * sizeof(T));
        dy1 = (T*)ensure(h, hdp_context::B_B, E * n_el * sizeof(T));
        if (!dy0 || !dy1) return HDP_ERR_ALLOC;
        HDP_CUDA(h, cudaMemcpyAsync(dy1, y1, E * n_el * sizeof(T), cudaMemcpyHostToDevice, h->stream));
        if (dWall_inject) {
            T* t = (T*)ensure(h, hdp_context::B_C, E * n_el * (size_t)nsteps * sizeof(T));
            if (!t) return HDP_ERR_ALLOC;
            HDP_CUDA(h, cudaMemcpyAsync(t, dWall_inject, E * n_el * (size_t)nsteps * sizeof(T), cudaMemcpyHostToDevice, h->stream));
            ddw = t;
        }
        if (want_z) {
            dz = (T*)ensure(h, hdp_context::B_D, E * n_el * (size_t)K * sizeof(T));
            if (!dz) return HDP_ERR_ALLOC;
        }
    }
    if (!want_z) dz = nullptr;
    constexpr int VEC = Tr<T>::VEC;
    const size_t nv = (n_el + VEC - 1) / VEC;
    if (!draw_address_ok(HDP_STREAM_SDE, (u64)nsteps, (u64)nv) || !draw_address_ok(HDP_STREAM_MC, (u64)(K > 0 ? K : 0), (u64)nv))
        return set_err(h, HDP_ERR_BAD_ARG, "sde_loop!: more than 2^24 steps/samples or 2^37 vec-groups");
    const bool al = (n_el % VEC == 0) && aligned16(dy0) && aligned16(dy1) && aligned16(dz) && aligned16(ddw);
    const int threads = 128;
    const int grid_x = grid_for(h, nv, threads, 64);
    const dim3 grid((unsigned)grid_x, (unsigned)n_epochs); // blockIdx.y = epoch (k_sde.cuh)
    const bool const_coef = prob->mu_kind == HDP_COEF_CONST && prob->sigma_kind == HDP_COEF_CONST && !mlp_assoc;
    int bc = !const_coef ? K1_GENERIC
             : !pb.has_bc ? K1_NOBC
             : (pb.bc_is_scalar && pb.lo_s == -pb.hi_s) ? K1_SYM : K1_GENERIC;
    if (bc != K1_GENERIC && pb.mu_c == (T)0) bc += 2; // K1_NOBC0 / K1_SYM0
#define LAUNCH3(AL, INJ, BC)                                                                              \
    sde_loop_kernel<T, AL, INJ, BC><<<grid, threads, 0, h->stream>>>(pb, dy0, dy1, n_el, nsteps, dt, x0s, \
                                                                     nullptr, seed, ddw, mcs, dz, K,       \
                                                                     h->d_flags, h->d_ctr + 1, mlp_assoc, 0, dts)
#define LAUNCH(AL, INJ)                                                                                   \
    do {                                                                                                  \
        if (bc == K1_SYM) LAUNCH3(AL, INJ, K1_SYM);                                                       \
        else if (bc == K1_NOBC) LAUNCH3(AL, INJ, K1_NOBC);                                                \
        else if (bc == K1_SYM0) LAUNCH3(AL, INJ, K1_SYM0);                                                \
        else if (bc == K1_NOBC0) LAUNCH3(AL, INJ, K1_NOBC0);                                              \
        else LAUNCH3(AL, INJ, K1_GENERIC);                                                                \
    } while (0)
    HDP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    if (um) {
        const T* x0_inject = nullptr;
        int* flags = h->d_flags;
        unsigned long long* bctr = h->d_ctr + 1;
        int zero = 0;
        size_t n_el_ = n_el;
        u64 seed_ = seed;
        void* args[] = {&pb, &dy0, &dy1, &n_el_, &nsteps, &dt, &x0s, &x0_inject, &seed_, &ddw, &mcs, &dz, &K,
                        &flags, &bctr, &mlp_assoc, &zero, &dts};
        HDP_CUDA(h, cudaLaunchKernel((const void*)um->sde[al ? 1 : 0][ddw ? 1 : 0], grid, dim3(threads), args, 0, h->stream));
    } else if (al && ddw) LAUNCH(true, true);
    else if (al) LAUNCH(true, false);
    else if (ddw) LAUNCH(false, true);
    else LAUNCH(false, false);
#undef LAUNCH3
#undef LAUNCH
    HDP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    h->launches++;
    HDP_CUDA(h, cudaGetLastError());
    if (!(mem_flags & HDP_MEM_DEVICE)) {
        HDP_CUDA(h, cudaMemcpyAsync(y0, dy0, E * n_el * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
        HDP_CUDA(h, cudaMemcpyAsync(y1, dy1, E * n_el * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
        if (want_z) HDP_CUDA(h, cudaMemcpyAsync(z, dz, E * n_el * (size_t)K * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    }
    if ((mem_flags & HDP_MEM_DEVICE) && (mem_flags & HDP_ASYNC)) return HDP_OK;
    if (prob->has_bc || !(mem_flags & HDP_MEM_DEVICE)) return check_flag(h, "sde_loop!");
    return HDP_OK;
}

// `_mlt_sde_loop!`: one step with the MLP association; x2 receives the result, x is untouched
template <typename T>
int mlt_sde_loop_impl(hdp_handle h, const hdp_problem* prob, T* x2, const T* x, int B, T s, T t, uint64_t seed,
                      const T* inject, unsigned mem_flags)
{
    if (!h || !prob || !x2 || !x || B < 0) return HDP_ERR_BAD_ARG;
    const size_t n_el = (size_t)prob->d * (size_t)B;
    if (n_el == 0) return HDP_OK;
    {
        // x2 <- x, then step x2 in place; the kernel's y0 output (the start state) goes to scratch
        std::lock_guard<std::mutex> lk(h->mu);
        DeviceGuard g(h->device);
        if (mem_flags & HDP_MEM_DEVICE)
            HDP_CUDA(h, cudaMemcpyAsync(x2, x, n_el * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
        else memcpy(x2, x, n_el * sizeof(T));
    }
    T* y0 = nullptr;
    std::vector<T> y0_host;
    if (mem_flags & HDP_MEM_DEVICE) {
        std::lock_guard<std::mutex> lk(h->mu);
        DeviceGuard g(h->device);
        y0 = (T*)ensure(h, hdp_context::B_H, n_el * sizeof(T));
        if (!y0) return HDP_ERR_ALLOC;
    } else {
        y0_host.resize(n_el);
        y0 = y0_host.data();
    }
    // μ(x, p, t), σ(x, p, t) take the END time of the step (src/MLP.jl:417)
    return sde_loop_impl<T>(h, prob, y0, x2, B, 1, (T)(t - s), &t, nullptr, seed, inject, nullptr, nullptr, 0,
                            HDP_NORMAL_DEFAULT, mem_flags & ~(unsigned)HDP_ASYNC, /*mlp_assoc=*/1);
}

// ------------------------------ K1e host -----------------------------
template <typename T>
int em_ensemble_impl(hdp_handle h, const hdp_problem* prob, T* x, int B, int nsteps, T dt, T t0, T* traj,
                     uint64_t seed, const T* inject, unsigned mem_flags)
{
    if (!h || !prob || !x || B < 0 || nsteps < 0) return HDP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    DevProblem<T> pb;
    hdp_problem pr = *prob; // no reflection on this path: the solvers it serves have no neumann_bc
    pr.has_bc = 0;
    int rc = make_dev_problem<T>(h, &pr, hdp_context::B_E, &pb, /*allow_user=*/true);
    if (rc != HDP_OK) return rc;
    const int d = prob->d;
    const size_t n_el = (size_t)d * (size_t)B;
    if (n_el == 0) return HDP_OK;
    SdeUserModule* um = nullptr;
    if (uses_user_coef(prob) && (rc = get_sde_user_module(h, prob, sizeof(T) == 4, &um)) != HDP_OK) return rc;
    const size_t n_traj = n_el * ((size_t)nsteps + 1);
    T *dx, *dtraj = nullptr;
    const T* ddw = nullptr;
    if (mem_flags & HDP_MEM_DEVICE) { dx = x; dtraj = traj; ddw = inject; }
    else {
        dx = (T*)ensure(h, hdp_context::B_A, n_el * sizeof(T));
        if (!dx) return HDP_ERR_ALLOC;
        HDP_CUDA(h, cudaMemcpyAsync(dx, x, n_el * sizeof(T), cudaMemcpyHostToDevice, h->stream));
        if (inject) {
            T* t = (T*)ensure(h, hdp_context::B_C, n_el * (size_t)nsteps * sizeof(T));
            if (!t) return HDP_ERR_ALLOC;
            HDP_CUDA(h, cudaMemcpyAsync(t, inject, n_el * (size_t)nsteps * sizeof(T), cudaMemcpyHostToDevice, h->stream));
            ddw = t;
        }
        if (traj) {
            dtraj = (T*)ensure(h, hdp_context::B_D, n_traj * sizeof(T));
            if (!dtraj) return HDP_ERR_ALLOC;
        }
    }
    constexpr int VEC = Tr<T>::VEC;
    const size_t nv = (n_el + VEC - 1) / VEC;
    if (!draw_address_ok(HDP_STREAM_SDE, (u64)nsteps, (u64)nv))
        return set_err(h, HDP_ERR_BAD_ARG, "em_ensemble: more than 2^24 steps or 2^37 vec-groups");
    // vector accesses need whole vec-groups inside one column of the trajectory array
    const bool al = (n_el % VEC == 0) && aligned16(dx) && aligned16(ddw) && (!dtraj || (d % VEC == 0 && aligned16(dtraj)));
    const int threads = 128;
    const int grid = grid_for(h, nv, threads, 64);
    HDP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    if (um) {
        size_t n_el_ = n_el;
        u64 seed_ = seed;
        int zero = 0;
        void* args[] = {&pb, &dx, &n_el_, &nsteps, &dt, &t0, &dtraj, &seed_, &ddw, &zero};
        HDP_CUDA(h, cudaLaunchKernel((const void*)um->ens[al ? 1 : 0][ddw ? 1 : 0][dtraj ? 1 : 0], dim3(grid), dim3(threads), args, 0, h->stream));
    } else {
#define LAUNCH_E(AL, INJ, TR) \
    em_ensemble_kernel<T, AL, INJ, TR><<<grid, threads, 0, h->stream>>>(pb, dx, n_el, nsteps, dt, t0, dtraj, seed, ddw, 0)
#define LAUNCH_E2(AL, INJ)              \
    do {                                \
        if (dtraj) LAUNCH_E(AL, INJ, true); \
        else LAUNCH_E(AL, INJ, false);  \
    } while (0)
        if (al && ddw) LAUNCH_E2(true, true);
        else if (al) LAUNCH_E2(true, false);
        else if (ddw) LAUNCH_E2(false, true);
        else LAUNCH_E2(false, false);
#undef LAUNCH_E2
#undef LAUNCH_E
    }
    HDP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    h->launches++;
    HDP_CUDA(h, cudaGetLastError());
    if (mem_flags & HDP_MEM_DEVICE) {
        if (!(mem_flags & HDP_ASYNC)) HDP_CUDA(h, cudaStreamSynchronize(h->stream));
        return HDP_OK;
    }
    HDP_CUDA(h, cudaMemcpyAsync(x, dx, n_el * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    if (traj) HDP_CUDA(h, cudaMemcpyAsync(traj, dtraj, n_traj * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    HDP_CUDA(h, cudaStreamSynchronize(h->stream));
    return HDP_OK;
}

} // namespace

extern "C" int hdp_set_user_exprs(hdp_handle h, const hdp_user_exprs* user)
{
    if (!h) return HDP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    h->has_user_mu = user && user->mu;
    h->has_user_sigma = user && user->sigma;
    h->user_mu = h->has_user_mu ? user->mu : "";
    h->user_sigma = h->has_user_sigma ? user->sigma : "";
    return HDP_OK;
}
extern "C" int hdp_em_ensemble_f64(hdp_handle h, const hdp_problem* prob, double* x, int B, int nsteps, double dt,
                                   double t0, double* traj, uint64_t seed, const double* normals_inject,
                                   unsigned mem_flags)
{
    return em_ensemble_impl<double>(h, prob, x, B, nsteps, dt, t0, traj, seed, normals_inject, mem_flags);
}
extern "C" int hdp_em_ensemble_f32(hdp_handle h, const hdp_problem* prob, float* x, int B, int nsteps, float dt,
                                   float t0, float* traj, uint64_t seed, const float* normals_inject,
                                   unsigned mem_flags)
{
    return em_ensemble_impl<float>(h, prob, x, B, nsteps, dt, t0, traj, seed, normals_inject, mem_flags);
}
extern "C" int hdp_mlt_sde_loop_f64(hdp_handle h, const hdp_problem* prob, double* x2, const double* x, int B,
                                    double s, double t, uint64_t seed, const double* normals_inject, unsigned mem_flags)
{
    return mlt_sde_loop_impl<double>(h, prob, x2, x, B, s, t, seed, normals_inject, mem_flags);
}
extern "C" int hdp_mlt_sde_loop_f32(hdp_handle h, const hdp_problem* prob, float* x2, const float* x, int B,
                                    float s, float t, uint64_t seed, const float* normals_inject, unsigned mem_flags)
{
    return mlt_sde_loop_impl<float>(h, prob, x2, x, B, s, t, seed, normals_inject, mem_flags);
}

extern "C" int hdp_reflect_f64(hdp_handle h, const double* a, const double* b, double* out, int d, int B,
                               const double* lo, const double* hi, int bounds_are_scalar,
                               int32_t* bounce_count_out, unsigned mem_flags)
{
    return reflect_impl<double>(h, a, b, out, d, B, lo, hi, bounds_are_scalar, bounce_count_out, mem_flags);
}
extern "C" int hdp_reflect_f32(hdp_handle h, const float* a, const float* b, float* out, int d, int B,
                               const float* lo, const float* hi, int bounds_are_scalar,
                               int32_t* bounce_count_out, unsigned mem_flags)
{
    return reflect_impl<float>(h, a, b, out, d, B, lo, hi, bounds_are_scalar, bounce_count_out, mem_flags);
}
extern "C" int hdp_mc_sample_f64(hdp_handle h, double* z, const double* y1, int d, int B, int K,
                                 const hdp_sampler* sampler, uint64_t seed, const double* inject, unsigned mem_flags)
{
    return mc_sample_impl<double>(h, z, y1, d, B, K, sampler, seed, inject, mem_flags);
}
extern "C" int hdp_mc_sample_f32(hdp_handle h, float* z, const float* y1, int d, int B, int K,
                                 const hdp_sampler* sampler, uint64_t seed, const float* inject, unsigned mem_flags)
{
    return mc_sample_impl<float>(h, z, y1, d, B, K, sampler, seed, inject, mem_flags);
}
extern "C" int hdp_sde_loop_f64(hdp_handle h, const hdp_problem* prob, double* y0, double* y1, int B, int nsteps,
                                double dt, const double* ts, const hdp_sampler* x0_sampler, uint64_t seed,
                                const double* dWall_inject, const hdp_sampler* mc, double* z, int K,
                                int normal_mode, unsigned mem_flags)
{
    return sde_loop_impl<double>(h, prob, y0, y1, B, nsteps, dt, ts, x0_sampler, seed, dWall_inject, mc, z, K,
                                 normal_mode, mem_flags);
}
extern "C" int hdp_sde_loop_f32(hdp_handle h, const hdp_problem* prob, float* y0, float* y1, int B, int nsteps,
                                float dt, const float* ts, const hdp_sampler* x0_sampler, uint64_t seed,
                                const float* dWall_inject, const hdp_sampler* mc, float* z, int K,
                                int normal_mode, unsigned mem_flags)
{
    return sde_loop_impl<float>(h, prob, y0, y1, B, nsteps, dt, ts, x0_sampler, seed, dWall_inject, mc, z, K,
                                normal_mode, mem_flags);
}
extern "C" int hdp_sde_loop_epochs_f64(hdp_handle h, const hdp_problem* prob, double* y0, double* y1, int B, int nsteps,
                                       double dt, const double* ts, const hdp_sampler* x0_sampler, uint64_t seed,
                                       int n_epochs, const hdp_sampler* mc, double* z, int K, unsigned mem_flags)
{
    return sde_loop_impl<double>(h, prob, y0, y1, B, nsteps, dt, ts, x0_sampler, seed, nullptr, mc, z, K,
                                 HDP_NORMAL_DEFAULT, mem_flags, 0, n_epochs);
}
extern "C" int hdp_sde_loop_epochs_f32(hdp_handle h, const hdp_problem* prob, float* y0, float* y1, int B, int nsteps,
                                       float dt, const float* ts, const hdp_sampler* x0_sampler, uint64_t seed,
                                       int n_epochs, const hdp_sampler* mc, float* z, int K, unsigned mem_flags)
{
    return sde_loop_impl<float>(h, prob, y0, y1, B, nsteps, dt, ts, x0_sampler, seed, nullptr, mc, z, K,
                                HDP_NORMAL_DEFAULT, mem_flags, 0, n_epochs);
}
