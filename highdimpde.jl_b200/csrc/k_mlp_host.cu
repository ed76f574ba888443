// k_mlp_host.cu — host driver of K3 (multilevel Picard): validates the
// descriptors, plans the tree cut, launches init / expand / units / combine,
// all-reduces the per-unit partial sums when a communicator is attached.
// Entry points replace `_ml_picard*` as called from solve(::MLP)
// (src/MLP.jl:68-98).
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "hdp_host_util.h"
#include "k_mlp_launch.h"
#include "mlp_plan.h"
#include "user_kinds.h"

using namespace hdp;

namespace hdp {
MlpPlan::~MlpPlan()
{
    void* ps[] = {d_nodes, d_sites, d_units, d_children, d_exec, d_sites_by_depth, d_nodes_by_depth, d_unit_tag};
    for (void* p : ps)
        if (p) cudaFree(p);
}
} // namespace hdp

namespace {

template <typename T> MlpLaunchers<T> launchers_for(int G);
#define HDP_CASE(T, TN, G) case G: return mlp_launchers_##TN##_g##G();
template <> MlpLaunchers<double> launchers_for<double>(int G)
{
    switch (G) { HDP_FOR_ALL_G(HDP_CASE, double, f64) }
    return MlpLaunchers<double>{};
}
template <> MlpLaunchers<float> launchers_for<float>(int G)
{
    switch (G) { HDP_FOR_ALL_G(HDP_CASE, float, f32) }
    return MlpLaunchers<float>{};
}

// Lane-group width (measured on B200, profiles/r1k_group_sweep.txt).
// Throughput regime (a batch with plenty of work): the narrowest group with at most 64 vec-groups
// per lane (d=100 → 1 lane, d=1000 → 8 lanes; C2 8.2 → 7.2 ms going from 4 lanes to 1, C5
// 9.9 → 8.9 ms from 32 to 8): fewest idle lanes in the last pass, fewest shuffles, and the per-step
// bookkeeping is spread over more coordinate work.
// Latency regime: a narrow group makes every work unit take longer, so when there are few
// path-steps per resident lane group (or only a handful of solves, whose trees cut into coarse
// units) the group is widened until each resident group has about a dozen path-steps to do.
int pick_group(int nvec, double path_steps_total, int n_solves, double sms)
{
    const char* env = getenv("HDP_MLP_GROUP");
    if (env) {
        int g = atoi(env);
        if (g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32) return g;
    }
    int G = 1;
    while (G < 32 && (nvec + G - 1) / G > (n_solves < 8 ? 16 : 64)) G *= 2;
    auto resident_groups = [&](int g) {
        const int npl = (nvec + g - 1) / g;
        return sms * (npl >= 12 ? 4 : 6) * 128.0 / g; // blocks per SM of the two kernel variants
    };
    // (a batch never widens to the point where lanes of a group idle: 2 G <= nvec)
    const int gmax = n_solves < 8 ? nvec : nvec / 2;
    while (G < 32 && G < gmax && path_steps_total < 12.0 * resident_groups(G)) G *= 2;
    return G;
}

template <typename V> int upload(hdp_context* h, const std::vector<V>& v, void** dptr)
{
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(V);
    HDP_CUDA(h, cudaMalloc(dptr, bytes));
    if (!v.empty()) HDP_CUDA(h, cudaMemcpy(*dptr, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice));
    return HDP_OK;
}

int get_plan(hdp_context* h, int M, int L, int K, bool nonlocal, double thresh, int compat,
             std::shared_ptr<MlpPlan>* out)
{
    char key[160];
    snprintf(key, sizeof(key), "%d/%d/%d/%d/%.6g/%d", M, L, K, (int)nonlocal, thresh, compat);
    auto it = h->plans.find(key);
    if (it != h->plans.end()) { *out = it->second; return HDP_OK; }
    auto plan = std::make_shared<MlpPlan>();
    mlp_build_plan(M, L, K, nonlocal, thresh, compat, plan.get());
    int rc;
    if ((rc = upload(h, plan->nodes, &plan->d_nodes)) != HDP_OK) return rc;
    if ((rc = upload(h, plan->sites, &plan->d_sites)) != HDP_OK) return rc;
    if ((rc = upload(h, plan->units, &plan->d_units)) != HDP_OK) return rc;
    if ((rc = upload(h, plan->children, &plan->d_children)) != HDP_OK) return rc;
    if ((rc = upload(h, plan->exec_order, &plan->d_exec)) != HDP_OK) return rc;
    if ((rc = upload(h, plan->unit_tag, &plan->d_unit_tag)) != HDP_OK) return rc;
    if ((rc = upload(h, plan->sites_by_depth, &plan->d_sites_by_depth)) != HDP_OK) return rc;
    if ((rc = upload(h, plan->nodes_by_depth, &plan->d_nodes_by_depth)) != HDP_OK) return rc;
    if (h->plans.size() > 64) h->plans.clear();
    h->plans[key] = plan;
    *out = plan;
    return HDP_OK;
}

// NVRTC-built kernels for a problem with *_USER kinds (compiled once per handle and expression set)
int get_user_module(hdp_context* h, const hdp_problem* prob, const hdp_user_exprs* user, bool f32, int G,
                    UserModule** out)
{
    if (!user) return set_err(h, HDP_ERR_BAD_ARG, "the problem uses a *_USER kind but cfg.user is NULL");
    const std::string key = user_key(prob, user, f32, G);
    auto it = h->user_modules.find(key);
    if (it != h->user_modules.end()) { *out = &it->second; return HDP_OK; }
    UserBuild b;
    int rc = user_compile(prob, user, f32, G, &b);
    if (rc != HDP_OK) return set_err(h, rc, "%s", b.log.c_str());
    UserModule m;
    HDP_CUDA(h, cudaLibraryLoadData(&m.lib, b.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    HDP_CUDA(h, cudaLibraryGetKernel(&m.units, m.lib, b.units_name.c_str()));
    HDP_CUDA(h, cudaLibraryGetKernel(&m.expand, m.lib, b.expand_name.c_str()));
    HDP_CUDA(h, cudaLibraryGetKernel(&m.combine, m.lib, b.combine_name.c_str()));
    auto ins = h->user_modules.emplace(key, m);
    *out = &ins.first->second;
    return HDP_OK;
}

void compat_counts(const MlpCountTables& t, int NT, uint64_t* C, uint64_t* R, uint64_t* S)
{
    uint64_t c = 0, r = 0, s = 0;
    for (int tid = 1; tid <= NT; ++tid)
        for (int l = 0; l < t.L; ++l) {
            uint64_t n = (uint64_t)mlp_get_loop_num(t.M, (int64_t)t.Mpow[t.L - l], tid, NT);
            c += n * t.cs[l]; r += n * t.rs[l]; s += n * t.ss[l];
        }
    *C = c + t.Mpow[t.L]; *R = r; *S = s;
}

template <typename T>
int mlp_solve_impl(hdp_handle h, const hdp_problem* prob, const hdp_mlp_cfg* cfg, const T* x0, T* u_out,
                   hdp_mlp_stats* stats)
{
    if (!h) return HDP_ERR_BAD_ARG;
    if (!prob || !cfg || !x0 || !u_out) return set_err(h, HDP_ERR_BAD_ARG, "mlp_solve: NULL argument");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    const int M = cfg->M, L = cfg->L, K = cfg->K;
    const int n_solves = cfg->n_solves;
    if (M < 1 || K < 1 || L < 0 || L > kMaxL)
        return set_err(h, HDP_ERR_BAD_ARG, "MLP(M=%d, L=%d, K=%d): need M>=1, K>=1, 0<=L<=%d", M, L, K, kMaxL);
    if (n_solves < 1) return set_err(h, HDP_ERR_BAD_ARG, "n_solves = %d", n_solves);
    if (cfg->n_x0 != 1 && cfg->n_x0 != n_solves) return set_err(h, HDP_ERR_BAD_ARG, "n_x0 must be 1 or n_solves");
    if (cfg->normal_mode != HDP_NORMAL_DEFAULT &&
        cfg->normal_mode != (sizeof(T) == 8 ? HDP_NORMAL_BM64 : HDP_NORMAL_BM32))
        return set_err(h, HDP_ERR_UNSUPPORTED, "normal_mode %d not supported for this element type", cfg->normal_mode);
    const bool nonlocal = cfg->sampler.kind != HDP_SAMPLER_NONE;
    const bool f_nonlocal = (prob->f_kind >= HDP_F_NL_FIRST && prob->f_kind < HDP_F_USER) || prob->f_kind == HDP_F_NL_USER;
    if (f_nonlocal && !nonlocal)
        return set_err(h, HDP_ERR_BAD_ARG, "non-local f kind needs a sampler (MLP(mc_sample = ...))");
    const bool taped = cfg->r_tape || cfg->n_tape || cfg->s_tape;
    if (taped && n_solves != 1) return set_err(h, HDP_ERR_BAD_ARG, "injected tapes require n_solves == 1");
    if (taped && (!cfg->r_tape || !cfg->n_tape || (nonlocal && !cfg->s_tape)))
        return set_err(h, HDP_ERR_BAD_ARG, "injected tapes: r_tape, n_tape (and s_tape with a sampler) must all be given");
    const int compat = cfg->compat_mlt_threads > 0 ? cfg->compat_mlt_threads : 0;

    MlpParams<T> P;
    memset(&P, 0, sizeof(P));
    int rc = make_dev_problem<T>(h, prob, hdp_context::B_E, &P.pb, /*allow_user=*/true);
    if (rc != HDP_OK) return rc;
    const bool uses_user = user_needed(prob) != 0;
    if (prob->f_kind == HDP_F_NL_USER && cfg->sampler.kind == HDP_SAMPLER_NONE)
        return set_err(h, HDP_ERR_BAD_ARG, "non-local user f needs a sampler (MLP(mc_sample = ...))");
    if ((rc = make_dev_sampler<T>(h, &cfg->sampler, prob->d, hdp_context::B_F, &P.sm)) != HDP_OK) return rc;

    MlpCountTables tab;
    mlp_count_tables(M, L, K, nonlocal, &tab);
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->per_solve.path_steps = tab.C[L];
        stats->per_solve.time_uniforms = tab.R[L];
        stats->per_solve.sampler_calls = tab.S[L];
        stats->per_solve.f_evals = tab.F[L];
        stats->per_solve.g_evals = tab.G[L];
        if (compat) compat_counts(tab, compat, &stats->per_solve.path_steps, &stats->per_solve.time_uniforms,
                                  &stats->per_solve.sampler_calls);
    }
    if (L == 0) { // `_ml_picard` returns zero(elxType) (src/MLP.jl:117-119)
        for (int i = 0; i < n_solves; ++i) u_out[i] = (T)0;
        return HDP_OK;
    }

    constexpr int VEC = Tr<T>::VEC;
    const int d = prob->d;
    const int nvec = (d + VEC - 1) / VEC;
    const double min_unit = 8.0; // smallest work unit (path-steps + half the sampler calls)
    const int G = pick_group(nvec, tab.W[L] * (double)n_solves, n_solves, (double)h->prop.multiProcessorCount);
    const MlpLaunchers<T> LN = launchers_for<T>(G);
    if (!LN.units[0][0]) return set_err(h, HDP_ERR_UNSUPPORTED, "no kernel for group size %d", G);
    const int mode = em_mode_of<T>(P.pb, cfg->n_tape != nullptr); // the units kernel compiled for this mode
    UserModule* um = nullptr;
    if (uses_user && (rc = get_user_module(h, prob, cfg->user, sizeof(T) == 4, G, &um)) != HDP_OK) return rc;
    const int block = HDP_MLP_BLOCK;
    // dynamic shared memory of the units kernel: one Frame<T> per lane group (k_mlp.cuh, run_unit_iter)
    const size_t dyn_smem = (size_t)(block / G) * sizeof(Frame<T>);
    // long rows: three vec-groups per pass (Tune<U>, k_mlp.cuh)
    const int long_rows = (nvec + G - 1) / G >= 12 ? 1 : 0;
    int bps = 0;
    if (um) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, (const void*)um->units, block, dyn_smem) != cudaSuccess) {
            (void)cudaGetLastError();
            bps = 0;
        }
    } else bps = LN.max_blocks_per_sm[long_rows][mode](block, dyn_smem);
    if (bps < 1) return set_err(h, HDP_ERR_CUDA, "occupancy query failed for the MLP kernel");
    const int sms = h->prop.multiProcessorCount;
    const int nranks = h->nccl_comm ? h->nccl_nranks : 1;
    const int rank = h->nccl_comm ? h->nccl_rank : 0;

    // ---- cut the tree: at most `thresh` work per unit ----
    const double resident_groups = (double)sms * bps * block / G;
    // units per resident lane group: enough for the LPT-ordered dynamic fetch to balance, few
    // enough that the replicated expand/combine passes and the all-reduce stay small
    // (measured, C4 L=5: 1 GPU best at 8; 8 GPUs: 2 per rank gives 0.9 M units and 3.7x, 8 per rank
    // cuts one level deeper — 16 M units — and the all-reduce + combine then dominate)
    double oversub = nranks > 1 ? 2.0 : 8.0;
    if (const char* ev = getenv("HDP_MLP_OVERSUB")) oversub = std::max(1.0, atof(ev));
    double target = cfg->target_units > 0 ? (double)cfg->target_units : resident_groups * oversub * nranks;
    const double total_work = tab.W[L] * (double)n_solves;
    double thresh = total_work / target;
    if (thresh < min_unit) thresh = min_unit;
    if (thresh >= tab.W[L]) thresh = tab.W[L]; // one WHOLE unit per solve
    {
        // quantise so that nearby batch sizes share a plan
        double q = pow(2.0, floor(log2(thresh) * 4.0) / 4.0);
        thresh = thresh >= tab.W[L] ? tab.W[L] : std::max(min_unit, q);
    }
    std::shared_ptr<MlpPlan> plan;
    if ((rc = get_plan(h, M, L, K, nonlocal, thresh, compat, &plan)) != HDP_OK) return rc;

    if (!draw_address_ok(HDP_STREAM_STEP, tab.C[L] + 1, (u64)nvec) || !draw_address_ok(HDP_STREAM_SAMPLER, tab.S[L] + 1, (u64)nvec))
        return set_err(h, HDP_ERR_BAD_ARG, "MLP: more than 2^40 path-steps per solve or 2^21 vec-groups per point");
    P.M = M; P.K = K; P.L = L; P.Kp = tab.Kp;
    P.nvec = nvec;
    P.npl = (nvec + G - 1) / G;
    P.dpad = P.npl * G * VEC; // every lane may load its full vec-group
    for (int i = 0; i <= kMaxL; ++i) {
        P.C[i] = tab.C[i]; P.R[i] = tab.R[i]; P.S[i] = tab.S[i];
        P.cs[i] = tab.cs[i]; P.rs[i] = tab.rs[i]; P.ss[i] = tab.ss[i];
    }
    for (int i = 0; i <= kMaxL + 1; ++i) P.Mpow[i] = tab.Mpow[i];
    P.seed0 = cfg->seed;
    P.n_solves = n_solves;
    P.n_x0 = cfg->n_x0;
    P.units = (const PlanUnit*)plan->d_units;
    P.exec = (const int32_t*)plan->d_exec;
    P.unit_tag = (const int8_t*)plan->d_unit_tag;
    P.nodes = (const PlanNode*)plan->d_nodes;
    P.sites = (const PlanSite*)plan->d_sites;
    P.children = (const int32_t*)plan->d_children;
    P.sites_by_depth = (const int32_t*)plan->d_sites_by_depth;
    P.nodes_by_depth = (const int32_t*)plan->d_nodes_by_depth;
    P.n_units = (int)plan->units.size();
    P.n_nodes = (int)plan->nodes.size();
    P.n_sites = (int)plan->sites.size();
    P.n_points = plan->n_points;
    P.rank = rank;
    P.nranks = nranks;
    P.work_ctr = h->d_ctr;
    P.bounce_ctr = h->d_ctr + 1;
    P.flags = h->d_flags;

    // ---- launch geometry of the unit kernel ----
    const uint64_t n_work = (uint64_t)P.n_units * (uint64_t)n_solves;
    const uint64_t n_local = n_work > (uint64_t)rank ? (n_work - rank + nranks - 1) / nranks : 0;
    const uint64_t groups_per_block = block / G;
    uint64_t want_blocks = (n_local + groups_per_block - 1) / groups_per_block;
    int grid = (int)std::min<uint64_t>(std::max<uint64_t>(want_blocks, 1), (uint64_t)sms * bps);
    // the expand kernel is launched with one group per (site, solve); size scratch for the larger
    P.total_threads = (uint64_t)grid * block;

    // ---- per-call device state ----
    const size_t n_pts = (size_t)n_solves * P.n_points;
    T* d_points = (T*)ensure(h, hdp_context::B_A, n_pts * P.dpad * sizeof(T));
    T* d_feat = (T*)ensure(h, hdp_context::B_B, n_pts * sizeof(T));
    T* d_site_r = (T*)ensure(h, hdp_context::B_C, (size_t)n_solves * std::max(P.n_sites, 1) * sizeof(T));
    double* d_partial = (double*)ensure(h, hdp_context::B_D, (size_t)n_solves * P.n_units * sizeof(double));
    T* d_value = (T*)ensure(h, hdp_context::B_G, (size_t)n_solves * P.n_nodes * sizeof(T));
    T* d_scratch = (T*)ensure(h, hdp_context::B_H, (size_t)(2 * L) * P.npl * P.total_threads * VEC * sizeof(T));
    T* d_x0 = (T*)ensure(h, hdp_context::B_I, (size_t)cfg->n_x0 * d * sizeof(T));
    if (!d_points || !d_feat || !d_site_r || !d_partial || !d_value || !d_scratch || !d_x0) return HDP_ERR_ALLOC;
    P.points = d_points; P.feat = d_feat; P.site_r = d_site_r; P.partial = d_partial; P.value = d_value;
    P.scratch = d_scratch;

    if (taped) {
        uint64_t nC = tab.C[L], nR = tab.R[L], nS = tab.S[L];
        if (compat) compat_counts(tab, compat, &nC, &nR, &nS);
        T* tr = (T*)ensure(h, hdp_context::B_J, std::max<size_t>(nR, 1) * sizeof(T));
        T* tn = (T*)ensure(h, hdp_context::B_K, (size_t)nC * d * sizeof(T));
        T* ts = (T*)ensure(h, hdp_context::B_L, std::max<size_t>((size_t)nS * d, 1) * sizeof(T));
        if (!tr || !tn || !ts) return HDP_ERR_ALLOC;
        HDP_CUDA(h, cudaMemcpyAsync(tr, cfg->r_tape, (size_t)nR * sizeof(T), cudaMemcpyHostToDevice, h->stream));
        HDP_CUDA(h, cudaMemcpyAsync(tn, cfg->n_tape, (size_t)nC * d * sizeof(T), cudaMemcpyHostToDevice, h->stream));
        if (nonlocal)
            HDP_CUDA(h, cudaMemcpyAsync(ts, cfg->s_tape, (size_t)nS * d * sizeof(T), cudaMemcpyHostToDevice, h->stream));
        P.r_tape = tr; P.n_tape = tn; P.s_tape = nonlocal ? ts : nullptr;
    }

    const int launches0 = h->launches;
    HDP_CUDA(h, cudaMemcpyAsync(d_x0, x0, (size_t)cfg->n_x0 * d * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    HDP_CUDA(h, cudaMemsetAsync(h->d_ctr, 0, 2 * sizeof(unsigned long long), h->stream));
    if (nranks > 1)
        HDP_CUDA(h, cudaMemsetAsync(d_partial, 0, (size_t)n_solves * P.n_units * sizeof(double), h->stream));
    HDP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    {
        uint64_t tot = (uint64_t)n_solves * P.dpad;
        int gi = (int)std::min<uint64_t>((tot + 255) / 256, (uint64_t)sms * 8);
        mlp_init_kernel<T><<<gi, 256, 0, h->stream>>>(P, d_x0);
        h->launches++;
        HDP_CUDA(h, cudaGetLastError());
    }
    // top-down geometry of the explicit sites
    for (int dep = 0; dep <= plan->max_depth; ++dep) {
        const int off = plan->site_depth_off[dep];
        const int cnt = plan->site_depth_off[dep + 1] - off;
        if (cnt == 0) continue;
        uint64_t groups = (uint64_t)cnt * n_solves;
        int ge = (int)((groups * G + block - 1) / block);
        if (um) {
            int off_ = off, cnt_ = cnt;
            void* args[3] = {(void*)&P, (void*)&off_, (void*)&cnt_};
            HDP_CUDA(h, cudaLaunchKernel((const void*)um->expand, dim3(ge), dim3(block), args, 0, h->stream));
        } else HDP_CUDA(h, LN.expand(P, off, cnt, ge, block, h->stream));
        h->launches++;
    }
    HDP_CUDA(h, cudaEventRecord(h->evp[0], h->stream));
    // the work units
    if (n_local > 0) {
        if (um) {
            void* args[1] = {(void*)&P};
            HDP_CUDA(h, cudaLaunchKernel((const void*)um->units, dim3(grid), dim3(block), args, dyn_smem, h->stream));
        } else HDP_CUDA(h, LN.units[long_rows][mode](P, grid, block, dyn_smem, h->stream));
        h->launches++;
    }
    HDP_CUDA(h, cudaEventRecord(h->evp[1], h->stream));
    if (nranks > 1) {
        rc = nccl_allreduce_f64_device(h, d_partial, (size_t)n_solves * P.n_units);
        if (rc != HDP_OK) return rc;
    }
    HDP_CUDA(h, cudaEventRecord(h->evp[2], h->stream));
    // bottom-up combine
    for (int dep = plan->max_depth; dep >= 0; --dep) {
        const int off = plan->node_depth_off[dep];
        const int cnt = plan->node_depth_off[dep + 1] - off;
        if (cnt == 0) continue;
        // lanes per (node, solve): a warp where some node of this depth has many children, 4 lanes otherwise
        int most = 0;
        for (int q = off; q < off + cnt; ++q) {
            const PlanNode& nd = plan->nodes[plan->nodes_by_depth[q]];
            most = std::max(most, (nd.unit_end - nd.unit_begin) + (nd.site_end - nd.site_begin));
        }
        int nl = most > 16 ? 32 : 4;
        uint64_t tot = (uint64_t)cnt * n_solves * nl;
        int gc = (int)((tot + 127) / 128);
        if (um) { // f is applied in the combine too: the user module's build of it
            int off_ = off, cnt_ = cnt;
            void* args[4] = {(void*)&P, (void*)&off_, (void*)&cnt_, (void*)&nl};
            HDP_CUDA(h, cudaLaunchKernel((const void*)um->combine, dim3(gc), dim3(128), args, 0, h->stream));
        } else mlp_combine_kernel<T><<<gc, 128, 0, h->stream>>>(P, off, cnt, nl);
        h->launches++;
        HDP_CUDA(h, cudaGetLastError());
    }
    HDP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    // root values: value[solve][0]
    HDP_CUDA(h, cudaMemcpy2DAsync(u_out, sizeof(T), d_value, (size_t)P.n_nodes * sizeof(T), sizeof(T), n_solves,
                                  cudaMemcpyDeviceToHost, h->stream));
    int flags2[2] = {0, 0}; // [0] start point outside the box, [1] kernel / mode mismatch (internal)
    int& flag = flags2[0];
    unsigned long long ctr[2] = {0, 0};
    HDP_CUDA(h, cudaMemcpyAsync(flags2, h->d_flags, sizeof(flags2), cudaMemcpyDeviceToHost, h->stream));
    HDP_CUDA(h, cudaMemcpyAsync(ctr, h->d_ctr, sizeof(ctr), cudaMemcpyDeviceToHost, h->stream));
    HDP_CUDA(h, cudaStreamSynchronize(h->stream));
    if (stats) {
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        stats->device_ms = ms;
        cudaEvent_t seq[5] = {h->ev0, h->evp[0], h->evp[1], h->evp[2], h->ev1};
        for (int i = 0; i < 4; ++i) {
            float pm = 0;
            cudaEventElapsedTime(&pm, seq[i], seq[i + 1]);
            stats->phase_ms[i] = pm;
        }
        stats->bounces = ctr[1];
        stats->n_units = P.n_units;
        stats->n_nodes = P.n_nodes;
        stats->n_sites = P.n_sites;
        stats->group_size = G;
        stats->n_launches = h->launches - launches0;
    }
    if (flags2[1]) {
        cudaMemsetAsync(h->d_flags, 0, sizeof(flags2), h->stream);
        return set_err(h, HDP_ERR_UNSUPPORTED, "internal: the units kernel launched does not match the mode of the problem");
    }
    if (flag) {
        cudaMemsetAsync(h->d_flags, 0, sizeof(int), h->stream);
        return set_err(h, HDP_ERR_START_OUTSIDE_DOMAIN, "_reflect: a start point is not in the hypercube (src/reflect.jl:12)");
    }
    return HDP_OK;
}

} // namespace

extern "C" int hdp_mlp_counts_get(int M, int L, int K, int nonlocal, int compat_mlt_threads, hdp_mlp_counts* out)
{
    if (!out || M < 1 || K < 1 || L < 0 || L > kMaxL || compat_mlt_threads < 0) return HDP_ERR_BAD_ARG;
    MlpCountTables t;
    mlp_count_tables(M, L, K, nonlocal != 0, &t);
    out->path_steps = t.C[L];
    out->time_uniforms = t.R[L];
    out->sampler_calls = t.S[L];
    out->f_evals = t.F[L];
    out->g_evals = t.G[L];
    if (compat_mlt_threads > 0 && L > 0)
        compat_counts(t, compat_mlt_threads, &out->path_steps, &out->time_uniforms, &out->sampler_calls);
    return HDP_OK;
}

// HDP_COEF_QUADRATIC inside the recursion: served as the user expression c*(x*x) through the NVRTC build, so
// the statically compiled kernels stay exactly as they are (the reference's MLP tests never use this kind)
template <typename T>
int mlp_solve_entry(hdp_handle h, const hdp_problem* prob, const hdp_mlp_cfg* cfg, const T* x0, T* u_out,
                    hdp_mlp_stats* stats)
{
    if (!h || !prob || !cfg || (prob->mu_kind != HDP_COEF_QUADRATIC && prob->sigma_kind != HDP_COEF_QUADRATIC))
        return mlp_solve_impl<T>(h, prob, cfg, x0, u_out, stats);
    hdp_problem p2 = *prob;
    hdp_mlp_cfg c2 = *cfg;
    hdp_user_exprs ue = cfg->user ? *cfg->user : hdp_user_exprs{nullptr, nullptr, nullptr, nullptr};
    if (prob->mu_kind == HDP_COEF_QUADRATIC) { p2.mu_kind = HDP_COEF_USER; ue.mu = "p[0]*(x*x)"; }
    if (prob->sigma_kind == HDP_COEF_QUADRATIC) { p2.sigma_kind = HDP_COEF_USER; ue.sigma = "p[0]*(x*x)"; }
    c2.user = &ue;
    return mlp_solve_impl<T>(h, &p2, &c2, x0, u_out, stats);
}

extern "C" int hdp_mlp_solve_f64(hdp_handle h, const hdp_problem* prob, const hdp_mlp_cfg* cfg, const double* x0,
                                 double* u_out, hdp_mlp_stats* stats)
{
    return mlp_solve_entry<double>(h, prob, cfg, x0, u_out, stats);
}
extern "C" int hdp_mlp_solve_f32(hdp_handle h, const hdp_problem* prob, const hdp_mlp_cfg* cfg, const float* x0,
                                 float* u_out, hdp_mlp_stats* stats)
{
    return mlp_solve_entry<float>(h, prob, cfg, x0, u_out, stats);
}
