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

static_assert(kPeerMax == HDP_PEER_MAX_RANKS && kPeerSlot == HDP_PEER_SLOT_DOUBLES, "peer window constants (include/hdpde.h)");

namespace hdp {
MlpPlan::~MlpPlan()
{
    void* ps[] = {d_nodes, d_sites, d_units, d_children, d_exec, d_sites_by_depth, d_nodes_by_depth, d_unit_tag};
    for (void* p : ps)
        if (p) cudaFree(p);
}
MlpSharding::~MlpSharding()
{
    void* ps[] = {d_exec_own, d_upper_sites, d_own_sites, d_owned_nodes, d_crown_nodes, d_crown_leaves,
                  d_crown_final, d_unit_mine, d_site_cls, d_node_slot, d_crown_jobs, d_crown_job_off};
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
// units) the group is widened until each resident group has about five path-steps to do.
// `forced` (cfg->group_size) overrides the rule; results never depend on the width.
int pick_group(int forced, int nvec, double path_steps_total, int n_solves, double sms)
{
    if (forced > 0) return forced;
    auto resident_groups = [&](int g) {
        const int npl = (nvec + g - 1) / g;
        return sms * (npl >= 12 ? 4 : 6) * 128.0 / g; // blocks per SM of the two kernel variants
    };
    // (a handful of solves that are each large — Allen-Cahn d=100, M=8, L=6: 19.5 M path-steps, 257 per resident
    // lane — are in the throughput regime like a batch: 7.85 ms with one lane per point against 9.29 ms with four)
    // In between (8..64 path-steps per resident lane: the same solve on 8 GPUs, 2.4 M path-steps per rank) rows of
    // up to 32 vec-groups per lane: 1.40 ms with two lanes per point, 1.54 with four, 1.66 with one.
    // (re-measured on the shares of one rank of 2, 4 and 8, profiles/r2z_fake*_minunit.txt: from 32 path-steps per
    // resident lane one lane per d=100 point once the cut may go down to four path-steps per unit — 1.19 -> 1.14 ms for
    // a rank of eight of M=8, L=6 —, and a full warp per point only for rows beyond 32 vec-groups per lane or almost
    // no work per lane: d=1000 with 16 lanes is 3-8 % faster than with 32 on 2, 4 and 8 ranks)
    const double per_lane = path_steps_total / resident_groups(1);
    const int cap = n_solves >= 8 || per_lane >= 32.0 ? 64 : (per_lane >= 8.0 ? 32 : 16);
    int G = 1;
    while (G < 32 && (nvec + G - 1) / G > ((G >= 16 && per_lane >= 1.5) ? std::max(cap, 32) : cap)) G *= 2;
    // (a batch never widens to the point where lanes of a group idle: 2 G <= nvec)
    const int gmax = n_solves < 8 ? nvec : nvec / 2;
    // (five path-steps per resident lane group: re-measured on the share of one rank of eight of Allen-Cahn d=100, M=8,
    // L=5 — 148 k path-steps — where the earlier twelve widened to 16 lanes: 0.31 ms against 0.26 with four,
    // profiles/r2z_fake8_sweep.txt; the batch sizes of profiles/r1k_group_sweep.txt keep their best or second-best width)
    while (G < 32 && G < gmax && path_steps_total < 5.0 * resident_groups(G)) G *= 2;
    return G;
}

template <typename V> int upload(hdp_context* h, const std::vector<V>& v, void** dptr)
{
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(V);
    HDP_CUDA(h, cudaMalloc(dptr, bytes));
    if (!v.empty()) HDP_CUDA(h, cudaMemcpy(*dptr, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice));
    return HDP_OK;
}

// The planner materialises every unit (56 bytes each on host and device): a cut finer than kMaxPlanUnits
// units per solve is refused and retried with twice the threshold (a plan that failed is remembered as an
// alias of the one that replaced it).
// A sharded solve holds the whole plan on every rank but works on 1/nranks of it: it may be cut finer (and so
// may a caller who names the number of units, cfg->target_units).
constexpr size_t kMaxPlanUnits = 3u << 20, kMaxPlanUnitsSharded = 8u << 20;
int get_plan(hdp_context* h, int M, int L, int K, bool nonlocal, double thresh, int compat, size_t max_units,
             std::shared_ptr<MlpPlan>* out, std::string* key_out)
{
    char key[160];
    snprintf(key, sizeof(key), "%d/%d/%d/%d/%.6g/%d", M, L, K, (int)nonlocal, thresh, compat);
    *key_out = key;
    auto it = h->plans.find(key);
    if (it != h->plans.end()) {
        *out = it->second;
        snprintf(key, sizeof(key), "%d/%d/%d/%d/%.6g/%d", M, L, K, (int)nonlocal, it->second->thresh, compat);
        *key_out = key; // (the key of the plan actually built: shardings hang off it)
        return HDP_OK;
    }
    const std::string asked = key;
    auto plan = std::make_shared<MlpPlan>();
    while (!mlp_build_plan(M, L, K, nonlocal, thresh, compat, plan.get(), max_units)) thresh *= 2.0;
    snprintf(key, sizeof(key), "%d/%d/%d/%d/%.6g/%d", M, L, K, (int)nonlocal, thresh, compat);
    *key_out = key;
    it = h->plans.find(key);
    if (it != h->plans.end()) { h->plans[asked] = it->second; *out = it->second; return HDP_OK; }
    int rc;
    if ((rc = upload(h, plan->nodes, &plan->d_nodes)) != HDP_OK) return rc;
    if ((rc = upload(h, plan->sites, &plan->d_sites)) != HDP_OK) return rc;
    if ((rc = upload(h, plan->units, &plan->d_units)) != HDP_OK) return rc;
    if ((rc = upload(h, plan->children, &plan->d_children)) != HDP_OK) return rc;
    if ((rc = upload(h, plan->exec_order, &plan->d_exec)) != HDP_OK) return rc;
    if ((rc = upload(h, plan->unit_tag, &plan->d_unit_tag)) != HDP_OK) return rc;
    if ((rc = upload(h, plan->sites_by_depth, &plan->d_sites_by_depth)) != HDP_OK) return rc;
    if ((rc = upload(h, plan->nodes_by_depth, &plan->d_nodes_by_depth)) != HDP_OK) return rc;
    if (h->plans.size() > 64) { h->shardings.clear(); h->plans.clear(); }
    h->plans[key] = plan;
    h->plans[asked] = plan;
    *out = plan;
    return HDP_OK;
}

int get_sharding(hdp_context* h, const std::string& plan_key, const MlpPlan& plan, int nranks, int rank,
                 std::shared_ptr<MlpSharding>* out)
{
    const std::string key = plan_key + "#" + std::to_string(nranks) + "/" + std::to_string(rank);
    auto it = h->shardings.find(key);
    if (it != h->shardings.end()) { *out = it->second; return HDP_OK; }
    auto sh = std::make_shared<MlpSharding>();
    mlp_build_sharding(plan, nranks, rank, sh.get());
    int rc;
    if ((rc = upload(h, sh->exec_own, &sh->d_exec_own)) != HDP_OK) return rc;
    if ((rc = upload(h, sh->upper_sites, &sh->d_upper_sites)) != HDP_OK) return rc;
    if ((rc = upload(h, sh->own_sites, &sh->d_own_sites)) != HDP_OK) return rc;
    if ((rc = upload(h, sh->owned_nodes, &sh->d_owned_nodes)) != HDP_OK) return rc;
    if ((rc = upload(h, sh->crown_nodes, &sh->d_crown_nodes)) != HDP_OK) return rc;
    if ((rc = upload(h, sh->crown_leaves, &sh->d_crown_leaves)) != HDP_OK) return rc;
    if ((rc = upload(h, sh->crown_final, &sh->d_crown_final)) != HDP_OK) return rc;
    if ((rc = upload(h, sh->unit_mine, &sh->d_unit_mine)) != HDP_OK) return rc;
    if ((rc = upload(h, sh->site_cls, &sh->d_site_cls)) != HDP_OK) return rc;
    if ((rc = upload(h, sh->node_slot, &sh->d_node_slot)) != HDP_OK) return rc;
    mlp_build_crown_jobs(plan, sh.get()); // shares of the crown partial pass
    if ((rc = upload(h, sh->crown_jobs, &sh->d_crown_jobs)) != HDP_OK) return rc;
    if ((rc = upload(h, sh->crown_job_off, &sh->d_crown_job_off)) != HDP_OK) return rc;
    h->shardings[key] = sh;
    *out = sh;
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

// What the ranks of a sharded solve must agree on: the configuration, the problem and (a sample of) the start
// points.  FNV-1a; the four 16-bit chunks travel as doubles in the tail of the exchange buffer.
uint64_t fnv(uint64_t hsh, const void* p, size_t n)
{
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { hsh ^= b[i]; hsh *= 0x100000001b3ull; }
    return hsh;
}
template <typename T>
uint64_t call_hash(const hdp_problem* prob, const hdp_mlp_cfg* cfg, const T* x0_host, int n_units, double thresh)
{
    uint64_t hsh = 0xcbf29ce484222325ull;
    const int32_t ints[] = {cfg->M, cfg->L, cfg->K, cfg->n_solves, cfg->n_x0, cfg->compat_mlt_threads, cfg->sampler.kind,
                            cfg->sampler.shifted, prob->d, prob->mu_kind, prob->sigma_kind, prob->g_kind, prob->f_kind,
                            prob->has_bc, n_units, (int32_t)sizeof(T)};
    hsh = fnv(hsh, ints, sizeof(ints));
    hsh = fnv(hsh, &cfg->seed, sizeof(cfg->seed));
    hsh = fnv(hsh, &thresh, sizeof(thresh));
    hsh = fnv(hsh, &prob->t0, 2 * sizeof(double));
    hsh = fnv(hsh, prob->mu_p, sizeof(prob->mu_p));
    hsh = fnv(hsh, prob->sigma_p, sizeof(prob->sigma_p));
    hsh = fnv(hsh, prob->g_p, sizeof(prob->g_p));
    hsh = fnv(hsh, prob->f_p, sizeof(prob->f_p));
    if (x0_host) { // host start points: the first and the last 4 KB
        const size_t nb = (size_t)cfg->n_x0 * prob->d * sizeof(T), cap = 4096;
        hsh = fnv(hsh, x0_host, std::min(nb, cap));
        if (nb > cap) hsh = fnv(hsh, (const char*)x0_host + nb - cap, cap);
    }
    return hsh;
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
    const int gs = cfg->group_size;
    if (gs != 0 && gs != 1 && gs != 2 && gs != 4 && gs != 8 && gs != 16 && gs != 32)
        return set_err(h, HDP_ERR_BAD_ARG, "group_size = %d: must be 0 (automatic), 1, 2, 4, 8, 16 or 32", gs);
    if (cfg->mem_flags & ~(unsigned)(HDP_MEM_DEVICE | HDP_ASYNC))
        return set_err(h, HDP_ERR_BAD_ARG, "mem_flags = %u", cfg->mem_flags);
    const bool on_device = (cfg->mem_flags & HDP_MEM_DEVICE) != 0;
    const bool async = on_device && (cfg->mem_flags & HDP_ASYNC) != 0;
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
    T* se_out = (T*)cfg->std_err_out;
    if (L == 0) { // `_ml_picard` returns zero(elxType) (src/MLP.jl:117-119)
        if (on_device) {
            HDP_CUDA(h, cudaMemsetAsync(u_out, 0, (size_t)n_solves * sizeof(T), h->stream));
            if (se_out) HDP_CUDA(h, cudaMemsetAsync(se_out, 0, (size_t)n_solves * sizeof(T), h->stream));
            if (!async) HDP_CUDA(h, cudaStreamSynchronize(h->stream));
        } else
            for (int i = 0; i < n_solves; ++i) {
                u_out[i] = (T)0;
                if (se_out) se_out[i] = (T)0;
            }
        return HDP_OK;
    }

    constexpr int VEC = Tr<T>::VEC;
    const int d = prob->d;
    const int nvec = (d + VEC - 1) / VEC;
    double min_unit = 8.0; // smallest work unit (path-steps + half the sampler calls)
#ifdef HDP_PHASE_TRACE
    if (const char* mu = getenv("HDP_TRACE_MIN_UNIT")) min_unit = atof(mu); // debugging build only
#endif
    const int sms = h->prop.multiProcessorCount;
    int nranks = h->nccl_comm ? h->nccl_nranks : 1;
    int rank = h->nccl_comm ? h->nccl_rank : 0;
#ifdef HDP_PHASE_TRACE
    // debugging build only: the launch chain of rank r of n on ONE GPU, the exchange left out (the value is then
    // this rank's share, not the solution) — per-launch times of an 8-rank chain at the price of one GPU
    bool fake_shard = false;
    if (const char* fr = getenv("HDP_TRACE_FAKE_RANKS")) {
        int n = 1, r = 0;
        if (!h->nccl_comm && sscanf(fr, "%d/%d", &r, &n) == 2 && n > 1 && r >= 0 && r < n) { nranks = n; rank = r; fake_shard = true; }
    }
#endif
    // a sharded solve is in the latency regime by construction: the work of one rank decides the lane-group width
    const int G = pick_group(gs, nvec, tab.W[L] * (double)n_solves / nranks, n_solves, (double)sms);
    const MlpLaunchers<T> LN = launchers_for<T>(G);
    if (!LN.units[0][0]) return set_err(h, HDP_ERR_UNSUPPORTED, "no kernel for group size %d", G);
    const int mode = em_mode_of<T>(P.pb, cfg->n_tape != nullptr); // the units kernel compiled for this mode
    UserModule* um = nullptr;
    if (uses_user && (rc = get_user_module(h, prob, cfg->user, sizeof(T) == 4, G, &um)) != HDP_OK) return rc;
    const int block = HDP_MLP_BLOCK;
    // dynamic shared memory of the units kernel: one private Frame<T> per thread (k_mlp.cuh, run_unit_iter)
    const size_t dyn_smem = (size_t)block * sizeof(Frame<T>);
    // long rows: three vec-groups per pass, very long rows four (Tune<U>, k_mlp.cuh)
    const int npl_ = (nvec + G - 1) / G;
    int long_rows = npl_ >= 48 ? 2 : (npl_ >= 12 ? 1 : 0);
#ifdef HDP_PHASE_TRACE
    if (const char* lr = getenv("HDP_TRACE_LONG_ROWS")) long_rows = atoi(lr); // debugging build only: force the variant
#endif
    int bps = 0;
    if (um) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, (const void*)um->units, block, dyn_smem) != cudaSuccess) {
            (void)cudaGetLastError();
            bps = 0;
        }
    } else bps = LN.max_blocks_per_sm[long_rows][mode](block, dyn_smem);
    if (bps < 1) return set_err(h, HDP_ERR_CUDA, "occupancy query failed for the MLP kernel");

    // ---- cut the tree: at most `thresh` work per unit ----
    const double resident_groups = (double)sms * bps * block / G;
    // units per resident lane group (cfg->unit_oversub): enough for the LPT-ordered dynamic fetch to balance.
    // On one GPU 8 (measured, C4 L=5).  A sharded solve exchanges one double per crown node, whatever the
    // number of units below, so it can afford the same fine cut per rank.
    double oversub = cfg->unit_oversub > 0 ? std::max(1.0, cfg->unit_oversub) : 8.0;
    double target = cfg->target_units > 0 ? (double)cfg->target_units : resident_groups * oversub * nranks;
    const double total_work = tab.W[L] * (double)n_solves;
    double thresh = total_work / target;
    // A single solve (or a handful) whose wanted cut is finer than the smallest unit: where the coarser cut still gives
    // every resident lane group a unit, halving the smallest unit balances the groups better (a rank of eight of
    // Allen-Cahn d=1000, M=8: 0.71 -> 0.64 ms with 16 lanes; of two or four of d=100, M=8, L=5: 8-12 %); where it
    // does not (fewer units than groups: the time is one unit's), smaller units only add per-unit work.
    // (units of the coarser cut: ~ total_work / (1.36 * min_unit), measured on the M=8 trees)
    if (n_solves < 8 && thresh < min_unit && cfg->target_units <= 0 &&
        total_work / (1.36 * min_unit) >= resident_groups * nranks)
        min_unit *= 0.5;
    if (thresh < min_unit) thresh = min_unit;
    if (thresh >= tab.W[L]) thresh = tab.W[L]; // one WHOLE unit per solve
    {
        // quantise so that nearby batch sizes share a plan
        double q = pow(2.0, floor(log2(thresh) * 4.0) / 4.0);
        thresh = thresh >= tab.W[L] ? tab.W[L] : std::max(min_unit, q);
    }
    std::shared_ptr<MlpPlan> plan;
    std::string plan_key;
    if ((rc = get_plan(h, M, L, K, nonlocal, thresh, compat, (nranks > 1 || cfg->target_units > 0) ? kMaxPlanUnitsSharded : kMaxPlanUnits, &plan,
                       &plan_key)) != HDP_OK) return rc;
    std::shared_ptr<MlpSharding> sh;
    if (nranks > 1 && (rc = get_sharding(h, plan_key, *plan, nranks, rank, &sh)) != HDP_OK) return rc;

    if (!draw_address_ok(HDP_STREAM_STEP, tab.C[L] + 1, (u64)nvec) || !draw_address_ok(HDP_STREAM_SAMPLER, tab.S[L] + 1, (u64)nvec))
        return set_err(h, HDP_ERR_BAD_ARG, "MLP: more than 2^40 path-steps per solve or 2^21 vec-groups per point");
    P.M = M; P.K = K; P.L = L; P.Kp = tab.Kp;
    P.nvec = nvec;
    P.npl = (nvec + G - 1) / G;
    // Geometry passes: a pass of the chain that does not fill the GPU lasts as long as ONE path-step of one lane
    // group — d = 100 with one lane per point is 50 dependent passes over the row (25 us per launch), with a warp
    // it is two — so every pass picks its own width (launch_expand below), whatever the width of the units kernel.
    // Rows are plain arrays of vec-groups: only their padding depends on the width.  Short rows (d = 10) and
    // NVRTC-built modules (compiled for one width) keep the units kernel's.
    // (a batch keeps its rows as they were — its geometry passes fill the GPU at any width, and padding d = 100 to
    // 128 put the start points of the 32 units of a warp 1024 bytes apart, one L1 set: C2 lost 6 %)
    const int Ge = (!um && nvec >= 16 && G < 32 && n_solves < 8) ? 32 : G; // the widest a pass may take
    const int npl_e = (nvec + Ge - 1) / Ge;
    P.dpad = std::max(P.npl * G, npl_e * Ge) * VEC; // every lane may load its full vec-group
    if (Ge != G && ((size_t)P.dpad * sizeof(T)) % 512 == 0) P.dpad += 16 / (int)sizeof(T); // rows never a power of two apart
    for (int i = 0; i <= kMaxL; ++i) {
        P.C[i] = tab.C[i]; P.R[i] = tab.R[i]; P.S[i] = tab.S[i];
        P.cs[i] = tab.cs[i]; P.rs[i] = tab.rs[i]; P.ss[i] = tab.ss[i];
    }
    for (int i = 0; i <= kMaxL + 1; ++i) P.Mpow[i] = tab.Mpow[i];
    P.seed0 = cfg->seed;
    P.n_solves = n_solves;
    P.n_x0 = cfg->n_x0;
    P.units = (const PlanUnit*)plan->d_units;
    P.exec = sh ? (const int32_t*)sh->d_exec_own : (const int32_t*)plan->d_exec;
    P.n_exec = sh ? (int)sh->exec_own.size() : (int)plan->units.size();
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
    const uint64_t n_local = (uint64_t)P.n_exec * (uint64_t)n_solves;
    const uint64_t groups_per_block = block / G;
    uint64_t want_blocks = (n_local + groups_per_block - 1) / groups_per_block;
    int grid = (int)std::min<uint64_t>(std::max<uint64_t>(want_blocks, 1), (uint64_t)sms * bps);
    P.total_threads = (uint64_t)grid * block;

    // ---- per-call device state ----
    const size_t n_pts = (size_t)n_solves * P.n_points;
    T* d_points = (T*)ensure(h, hdp_context::B_A, n_pts * P.dpad * sizeof(T));
    T* d_feat = (T*)ensure(h, hdp_context::B_B, n_pts * sizeof(T));
    T* d_site_r = (T*)ensure(h, hdp_context::B_C, (size_t)n_solves * std::max(P.n_sites, 1) * sizeof(T));
    double* d_partial = (double*)ensure(h, hdp_context::B_D, (size_t)n_solves * P.n_units * sizeof(double));
    T* d_value = (T*)ensure(h, hdp_context::B_G, (size_t)n_solves * P.n_nodes * sizeof(T));
    T* d_scratch = (T*)ensure(h, hdp_context::B_H, (size_t)(2 * L) * P.npl * P.total_threads * VEC * sizeof(T));
    if (!d_points || !d_feat || !d_site_r || !d_partial || !d_value || !d_scratch) return HDP_ERR_ALLOC;
    const T* d_x0 = x0;
    if (!on_device) {
        T* buf = (T*)ensure(h, hdp_context::B_I, (size_t)cfg->n_x0 * d * sizeof(T));
        if (!buf) return HDP_ERR_ALLOC;
        d_x0 = buf;
    }
    P.points = d_points; P.feat = d_feat; P.site_r = d_site_r; P.partial = d_partial; P.value = d_value;
    P.scratch = d_scratch;

    // error bar: per root unit (or per root level when the solve is one unit) a (sum, sum of squares) pair
    const bool want_se = se_out != nullptr;
    T* d_se = nullptr;
    if (want_se) {
        const PlanNode& root = plan->nodes[0];
        P.rootsq_stride = 2 * std::max(root.unit_end - root.unit_begin, L + 1);
        P.rootsq = (double*)ensure(h, hdp_context::B_M, (size_t)n_solves * P.rootsq_stride * sizeof(double));
        P.lvstat = (double*)ensure(h, hdp_context::B_O, (size_t)n_solves * 3 * (kMaxL + 1) * sizeof(double));
        d_se = on_device ? se_out : (T*)ensure(h, hdp_context::B_P, (size_t)n_solves * sizeof(T));
        if (!P.rootsq || !P.lvstat || !d_se) return HDP_ERR_ALLOC;
        P.std_err = d_se;
    }
    // exchange buffer of a sharded solve: one double per crown node and solve, the root statistics, the call tail
    size_t x_count = 0;
    uint64_t my_hash = 0;
    if (sh) {
        P.unit_mine = (const int8_t*)sh->d_unit_mine;
        P.site_cls = (const int8_t*)sh->d_site_cls;
        P.node_slot = (const int32_t*)sh->d_node_slot;
        P.n_slots = sh->n_slots;
        P.xstride = sh->n_slots + 3 * (kMaxL + 1);
        x_count = (size_t)n_solves * P.xstride + kXTail;
        P.xbuf = (double*)ensure(h, hdp_context::B_Q, x_count * sizeof(double));
        if (!P.xbuf) return HDP_ERR_ALLOC;
        my_hash = call_hash<T>(prob, cfg, on_device ? nullptr : x0, P.n_units, thresh);
    }
    // peer windows attached on every rank and the exchange fits a slot: the library's own exchange kernel
    const bool use_peer = sh && h->peer_nranks == nranks && h->peer_rank == rank && x_count <= (size_t)kPeerSlot;

    if (taped) {
        uint64_t nC = tab.C[L], nR = tab.R[L], nS = tab.S[L];
        if (compat) compat_counts(tab, compat, &nC, &nR, &nS);
        if (on_device) {
            P.r_tape = (const T*)cfg->r_tape; P.n_tape = (const T*)cfg->n_tape;
            P.s_tape = nonlocal ? (const T*)cfg->s_tape : nullptr;
        } else {
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
    }

    const int launches0 = h->launches;
#ifdef HDP_PHASE_TRACE
    // debugging build (make EXTRA=-DHDP_PHASE_TRACE): an event after every launch, the per-launch times on stderr
    static cudaEvent_t tr_ev[64];
    static const char* tr_name[64];
    static int tr_cnt[64];
    int tr_n = 0;
    auto trace = [&](const char* name, int cnt) {
        if (tr_n >= 64) return;
        if (!tr_ev[tr_n]) cudaEventCreate(&tr_ev[tr_n]);
        cudaEventRecord(tr_ev[tr_n], h->stream);
        tr_name[tr_n] = name; tr_cnt[tr_n] = cnt;
        ++tr_n;
    };
#else
    auto trace = [](const char*, int) {};
#endif
    double tail_host[kXTail] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (!on_device)
        HDP_CUDA(h, cudaMemcpyAsync((void*)d_x0, x0, (size_t)cfg->n_x0 * d * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    HDP_CUDA(h, cudaMemsetAsync(h->d_ctr, 0, 2 * sizeof(unsigned long long), h->stream));
    if (sh) {
        for (int q = 0; q < 4; ++q) tail_host[q] = (double)((my_hash >> (16 * q)) & 0xffffu);
        HDP_CUDA(h, cudaMemcpyAsync(P.xbuf + (size_t)n_solves * P.xstride, tail_host, sizeof(tail_host),
                                    cudaMemcpyHostToDevice, h->stream));
    }
    HDP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    {
        uint64_t tot = (uint64_t)n_solves * P.dpad;
        int gi = (int)std::min<uint64_t>((tot + 255) / 256, (uint64_t)sms * 8);
        trace("start", 0);
        mlp_init_kernel<T><<<gi, 256, 0, h->stream>>>(P, d_x0);
        h->launches++;
        trace("init", gi);
        HDP_CUDA(h, cudaGetLastError());
    }
    auto launch_expand = [&](const int32_t* list, int cnt, int count_hits, const int32_t* list2 = nullptr, int cnt2 = 0,
                             int count_hits2 = 0) -> int {
        if (cnt + cnt2 == 0) return HDP_OK;
        const uint64_t site_items = (uint64_t)(cnt + cnt2) * n_solves;
        // one lane group per (site, solve, point) in a short pass, per (site, solve) in a long one (mlp_expand_kernel)
        uint64_t groups = site_items * (uint64_t)(site_items < kExpandSplitBelow ? 1 + tab.Kp : 1);
        // width of this pass: waves of resident lane groups x passes over a lane's share of the row (three
        // vec-groups per pass on long shares), the narrower on a tie
        int Gx = G;
        {
            const double cap = (double)sms * 12 * block;
            double best = 0;
            for (int g = G; g <= Ge; g *= 2) {
                const int npl_g = (nvec + g - 1) / g;
                const double cost = ceil((double)groups * g / cap) * (npl_g >= 12 ? (npl_g + 2) / 3 * 1.5 : (double)npl_g);
                if (g == G || cost < best) { best = cost; Gx = g; }
            }
        }
        int ge = (int)std::min<uint64_t>((groups * Gx + block - 1) / block, (uint64_t)sms * 12);
        if (um) {
            void* args[7] = {(void*)&P, (void*)&list, (void*)&cnt, (void*)&count_hits, (void*)&list2, (void*)&cnt2, (void*)&count_hits2};
            HDP_CUDA(h, launch_chained_module((const void*)um->expand, dim3(ge), dim3(block), 0, h->stream, args));
        } else if (Gx != G) {
            MlpParams<T> Px = P; // the same state, its own partition of the rows
            Px.npl = (nvec + Gx - 1) / Gx;
            HDP_CUDA(h, launchers_for<T>(Gx).expand(Px, list, cnt, count_hits, list2, cnt2, count_hits2, ge, block, h->stream));
        } else HDP_CUDA(h, LN.expand(P, list, cnt, count_hits, list2, cnt2, count_hits2, ge, block, h->stream));
        h->launches++;
        trace("expand", cnt + cnt2);
        return HDP_OK;
    };
    // lanes per (node, solve) of a combine launch: a warp where some node has many children, 4 lanes otherwise
    auto launch_combine = [&](const int32_t* list, const int32_t* host_list, int cnt, int cmode) -> int {
        if (cnt == 0) return HDP_OK;
        // the error bar is gathered at the root: only a launch that holds node 0 carries the level statistics
        bool with_stats = false;
        if (want_se)
            for (int q = 0; q < cnt; ++q) with_stats |= host_list[q] == 0;
        int most = 0, most_units = 0, most_weight = 0;
        for (int q = 0; q < cnt; ++q) most_units = std::max(most_units, plan->nodes[host_list[q]].unit_end - plan->nodes[host_list[q]].unit_begin);
        // nodes with very many units: several blocks each sum a share first (mlp_prereduce_kernel)
        const uint64_t items = (uint64_t)cnt * n_solves;
        const bool staged = cmode != CMB_FINAL && most_units > kBigNode && items <= 1024;
        MlpParams<T> Pc = P;
        if (staged) {
            double* stage = (double*)ensure(h, hdp_context::B_R, items * kStageBlocks * kCmbStage * sizeof(double));
            if (!stage) return HDP_ERR_ALLOC;
            if (with_stats)
                HDP_CUDA(h, launch_chained(mlp_prereduce_kernel<T, true>, dim3(kStageBlocks, (unsigned)items), dim3(256), 0, h->stream, P,
                                           list, cmode, stage, (const int2*)nullptr));
            else
                HDP_CUDA(h, launch_chained(mlp_prereduce_kernel<T, false>, dim3(kStageBlocks, (unsigned)items), dim3(256), 0, h->stream, P,
                                           list, cmode, stage, (const int2*)nullptr));
            h->launches++;
            trace("prereduce", cnt);
            Pc.stage = stage;
            Pc.stage_nb = kStageBlocks;
        }
        for (int q = 0; q < cnt; ++q) {
            const PlanNode& nd = plan->nodes[host_list[q]];
            const int nu = nd.unit_end - nd.unit_begin;
            // (CMB_FINAL walks only the replicated upper sites of a node, but sizing its lanes by those alone was
            // slower — Allen-Cahn M=8, L=6 on 8 ranks: 18 -> 36 us and 20 -> 50 us for the last two passes: these
            // passes are latency-bound walks, lanes are free)
            const int nus = (staged && nu > kBigNode) ? kStageBlocks : nu, ns = nd.site_end - nd.site_begin;
            most = std::max(most, nus + ns);
            most_weight = std::max(most_weight, nus + ns * (1 + tab.Kp));
        }
        // a handful of children: 4 lanes; many: a warp; thousands (the root of a finely cut tree): a block of 512 —
        // but only where the list is too short to fill the GPU with warps (Allen-Cahn d=100, M=8, L=6: 5264 calls of
        // depth 1 with a few thousand slice units each took 310 us as blocks of 512, a warp each is plenty)
        // In a short list — the root, alone — lanes are free, and a non-local site costs 1 + K evaluations of f with
        // their dependent loads: the root of C4 (300 units, 340 explicit sites) walked by one warp, eleven sites
        // per lane, held the chain for 90 us.
        const bool few = (uint64_t)cnt * n_solves <= 2 * (uint64_t)sms;
        const int nl = few ? (most_weight > 256 ? 512 : (most_weight > 16 ? 32 : 4)) : (most > 16 ? 32 : 4);
        const int cb = nl > 128 ? nl : 128; // threads per block (one node per block when nl > 32)
        uint64_t tot = (uint64_t)cnt * n_solves * nl;
        int gc = (int)((tot + cb - 1) / cb);
        if (um) { // f is applied in the combine too: the user module's build of it
            void* args[5] = {(void*)&Pc, (void*)&list, (void*)&cnt, (void*)&nl, (void*)&cmode};
            HDP_CUDA(h, launch_chained_module((const void*)um->combine, dim3(gc), dim3(cb), 0, h->stream, args));
        } else if (with_stats) HDP_CUDA(h, launch_chained(mlp_combine_kernel<T, true>, dim3(gc), dim3(cb), 0, h->stream, Pc, list, cnt, nl, cmode));
        else HDP_CUDA(h, launch_chained(mlp_combine_kernel<T, false>, dim3(gc), dim3(cb), 0, h->stream, Pc, list, cnt, nl, cmode));
        h->launches++;
        trace(cmode == CMB_ALL ? "combine_all" : (cmode == CMB_PARTIAL ? "combine_partial" : "combine_final"), cnt);
        HDP_CUDA(h, cudaGetLastError());
        return HDP_OK;
    };
    // ---- top-down geometry of the explicit sites ----
    if (!sh) {
        for (int dep = 0; dep <= plan->max_depth; ++dep) {
            const int off = plan->site_depth_off[dep];
            if ((rc = launch_expand(P.sites_by_depth + off, plan->site_depth_off[dep + 1] - off, 1)) != HDP_OK) return rc;
        }
    } else {
        // depth by depth: the upper sites on every rank (rank 0 counts their wall hits) together with the sites at
        // and below this rank's ownership roots (a site's start point comes from a site one depth up)
        for (int dep = 0; dep <= plan->max_depth; ++dep) {
            const int off = sh->upper_depth_off[dep], off2 = sh->own_depth_off[dep];
            if ((rc = launch_expand((const int32_t*)sh->d_upper_sites + off, sh->upper_depth_off[dep + 1] - off, rank == 0 ? 1 : 0,
                                    (const int32_t*)sh->d_own_sites + off2, sh->own_depth_off[dep + 1] - off2, 1)) != HDP_OK)
                return rc;
        }
    }
    HDP_CUDA(h, cudaEventRecord(h->evp[0], h->stream));
    // ---- the work units ----
    if (n_local > 0) {
        if (um) {
            void* args[1] = {(void*)&P};
            HDP_CUDA(h, launch_chained_module((const void*)um->units, dim3(grid), dim3(block), dyn_smem, h->stream, args));
        } else HDP_CUDA(h, LN.units[long_rows][mode](P, grid, block, dyn_smem, h->stream));
        h->launches++;
        trace("units", (int)n_local);
    }
    HDP_CUDA(h, cudaEventRecord(h->evp[1], h->stream));
    if (!sh) {
        HDP_CUDA(h, cudaEventRecord(h->evp[2], h->stream));
        // ---- bottom-up combine ----
        for (int dep = plan->max_depth; dep >= 0; --dep) {
            const int off = plan->node_depth_off[dep];
            if ((rc = launch_combine(P.nodes_by_depth + off, plan->nodes_by_depth.data() + off,
                                     plan->node_depth_off[dep + 1] - off, CMB_ALL)) != HDP_OK)
                return rc;
        }
    } else {
        // the calls below this rank's ownership roots, bottom-up, then this rank's share of every crown node
        for (int dep = plan->max_depth; dep >= 0; --dep) {
            const int off = sh->owned_depth_off[dep];
            if ((rc = launch_combine((const int32_t*)sh->d_owned_nodes + off, sh->owned_nodes.data() + off,
                                     sh->owned_depth_off[dep + 1] - off, CMB_ALL)) != HDP_OK)
                return rc;
        }
        const int n_jobs = (int)(sh->crown_jobs.size() / 2);
        // (one heavy node only — the root of C4, its other 4312 crown nodes are light — gains nothing over root first, rest after)
        if (!um && n_solves <= 65535 && sh->crown_heavy >= 2) {
            // every crown node in one pass: its units and this rank's sites summed by the node's share of blocks
            // (crown_jobs), then a warp per node adds the records up and writes the node's slot of the exchange buffer
            double* stage = (double*)ensure(h, hdp_context::B_R, (size_t)n_solves * n_jobs * kCmbStage * sizeof(double));
            if (!stage) return HDP_ERR_ALLOC;
            MlpParams<T> Pj = P;
            Pj.stage = stage;
            Pj.stage_off = (const int32_t*)sh->d_crown_job_off;
            Pj.stage_jobs = n_jobs;
            const int32_t* clist = (const int32_t*)sh->d_crown_nodes;
            const int cnt = (int)sh->crown_nodes.size();
            if (want_se)
                HDP_CUDA(h, launch_chained(mlp_prereduce_kernel<T, true>, dim3((unsigned)n_jobs, (unsigned)n_solves), dim3(256), 0, h->stream,
                                           Pj, clist, (int)CMB_PARTIAL, stage, (const int2*)sh->d_crown_jobs));
            else
                HDP_CUDA(h, launch_chained(mlp_prereduce_kernel<T, false>, dim3((unsigned)n_jobs, (unsigned)n_solves), dim3(256), 0, h->stream,
                                           Pj, clist, (int)CMB_PARTIAL, stage, (const int2*)sh->d_crown_jobs));
            h->launches++;
            trace("prereduce_jobs", n_jobs);
            const int nl = 32, cb = 128;
            const uint64_t tot = (uint64_t)cnt * n_solves * nl;
            if (want_se)
                HDP_CUDA(h, launch_chained(mlp_combine_kernel<T, true>, dim3((unsigned)((tot + cb - 1) / cb)), dim3(cb), 0, h->stream, Pj, clist,
                                           cnt, nl, (int)CMB_PARTIAL));
            else
                HDP_CUDA(h, launch_chained(mlp_combine_kernel<T, false>, dim3((unsigned)((tot + cb - 1) / cb)), dim3(cb), 0, h->stream, Pj, clist,
                                           cnt, nl, (int)CMB_PARTIAL));
            h->launches++;
            trace("combine_partial", cnt);
        } else {
            // (the root first and alone: it has thousands of children and takes a block of lanes, the other crown
            // nodes a few lanes each — one launch sized for the root spent 95 us on 4313 blocks of 512 threads)
            if ((rc = launch_combine((const int32_t*)sh->d_crown_nodes, sh->crown_nodes.data(), 1, CMB_PARTIAL)) != HDP_OK) return rc;
            if ((rc = launch_combine((const int32_t*)sh->d_crown_nodes + 1, sh->crown_nodes.data() + 1,
                                     (int)sh->crown_nodes.size() - 1, CMB_PARTIAL)) != HDP_OK)
                return rc;
        }
        HDP_CUDA(h, cudaEventRecord(h->evp[1], h->stream)); // (the local combine counts as unit work)
        if (use_peer) {
            // the library's own exchange over peer memory: push, flag, sum in rank order and the scatter of the
            // whole-call leaves in one launch (mlp_exchange_kernel)
            PeerWin W;
            memset(&W, 0, sizeof(W));
            for (int r = 0; r < nranks; ++r) W.base[r] = (double*)h->peer_base[r];
            W.nranks = nranks;
            W.rank = rank;
            W.epoch = ++h->peer_epoch;
            HDP_CUDA(h, launch_chained(mlp_exchange_kernel<T>, dim3(1), dim3(1024), 0, h->stream, P, W, (int)x_count,
                                       (const int32_t*)sh->d_crown_leaves, (int)sh->crown_leaves.size()));
            h->launches++;
            trace("exchange", (int)x_count);
            HDP_CUDA(h, cudaEventRecord(h->evp[2], h->stream));
        } else {
#ifdef HDP_PHASE_TRACE
            if (!fake_shard)
#endif
            rc = nccl_allreduce_f64_device(h, P.xbuf, x_count);
            if (rc != HDP_OK) return rc;
            trace("allreduce", (int)x_count);
            HDP_CUDA(h, cudaEventRecord(h->evp[2], h->stream));
            if (!sh->crown_leaves.empty()) {
                const int cnt = (int)sh->crown_leaves.size();
                const uint64_t tot = (uint64_t)cnt * n_solves;
                HDP_CUDA(h, launch_chained(mlp_scatter_kernel<T>, dim3((unsigned)((tot + 127) / 128)), dim3(128), 0, h->stream, P,
                                           (const int32_t*)sh->d_crown_leaves, cnt));
                h->launches++;
                trace("scatter", cnt);
            }
        }
        for (int dep = plan->max_depth; dep >= 0; --dep) {
            const int off = sh->crown_final_off[dep];
            if ((rc = launch_combine((const int32_t*)sh->d_crown_final + off, sh->crown_final.data() + off,
                                     sh->crown_final_off[dep + 1] - off, CMB_FINAL)) != HDP_OK)
                return rc;
        }
    }
    HDP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    // ---- results: value[solve][0] ----
    HDP_CUDA(h, cudaMemcpy2DAsync(u_out, sizeof(T), d_value, (size_t)P.n_nodes * sizeof(T), sizeof(T), n_solves,
                                  on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    if (want_se && !on_device)
        HDP_CUDA(h, cudaMemcpyAsync(se_out, d_se, (size_t)n_solves * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    if (stats) {
        stats->exchange_kind = sh ? (use_peer ? 2 : 1) : 0;
        stats->n_units = P.n_units;
        stats->n_nodes = P.n_nodes;
        stats->n_sites = P.n_sites;
        stats->group_size = G;
        stats->n_launches = h->launches - launches0;
        stats->exchange_bytes = x_count * sizeof(double);
    }
    if (async) return HDP_OK; // a start point outside the box surfaces at hdp_sync
    int flags2[3] = {0, 0, 0}; // [0] start point outside the box, [1] kernel / mode mismatch (internal), [2] peer exchange timed out
    int& flag = flags2[0];
    unsigned long long ctr[2] = {0, 0};
    double lv0[3 * (kMaxL + 1)];
    T se0 = (T)0;
    HDP_CUDA(h, cudaMemcpyAsync(flags2, h->d_flags, sizeof(flags2), cudaMemcpyDeviceToHost, h->stream));
    HDP_CUDA(h, cudaMemcpyAsync(ctr, h->d_ctr, sizeof(ctr), cudaMemcpyDeviceToHost, h->stream));
    if (sh)
        HDP_CUDA(h, cudaMemcpyAsync(tail_host, P.xbuf + (size_t)n_solves * P.xstride, sizeof(tail_host),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (want_se && stats) {
        HDP_CUDA(h, cudaMemcpyAsync(lv0, P.lvstat, sizeof(lv0), cudaMemcpyDeviceToHost, h->stream));
        HDP_CUDA(h, cudaMemcpyAsync(&se0, d_se, sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    }
    HDP_CUDA(h, cudaStreamSynchronize(h->stream));
#ifdef HDP_PHASE_TRACE
    if (getenv("HDP_PHASE_TRACE")) {
        fprintf(stderr, "[trace rank %d]", rank);
        for (int i = 1; i < tr_n; ++i) {
            float pm = 0;
            cudaEventElapsedTime(&pm, tr_ev[i - 1], tr_ev[i]);
            fprintf(stderr, " %s(%d)=%.1fus", tr_name[i], tr_cnt[i], pm * 1e3f);
        }
        fprintf(stderr, "\n");
    }
#endif
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
        if (want_se) {
            for (int l = 0; l <= L; ++l) {
                stats->level_n[l] = (uint64_t)(lv0[3 * l] + 0.5);
                stats->level_sum[l] = lv0[3 * l + 1];
                stats->level_sumsq[l] = lv0[3 * l + 2];
            }
            stats->std_error = (double)se0;
        }
    }
    if (flags2[2]) {
        cudaMemsetAsync(h->d_flags, 0, sizeof(flags2), h->stream);
        return set_err(h, HDP_ERR_NCCL, "sharded MLP solve: a peer did not reach the exchange within 20 s (every rank must make the "
                                        "same sequence of sharded calls; detach the peer windows to fall back to NCCL)");
    }
    if (flags2[1]) {
        cudaMemsetAsync(h->d_flags, 0, sizeof(flags2), h->stream);
        return set_err(h, HDP_ERR_UNSUPPORTED, "internal: the units kernel launched does not match the mode of the problem");
    }
#ifdef HDP_PHASE_TRACE
    if (fake_shard) return HDP_OK;
#endif
    if (sh) { // every rank must have described the same call: the all-reduced hash is nranks times mine
        for (int q = 0; q < 4; ++q)
            if (tail_host[q] != (double)nranks * (double)((my_hash >> (16 * q)) & 0xffffu))
                return set_err(h, HDP_ERR_BAD_ARG,
                               "sharded MLP solve: the ranks disagree on seed, configuration, problem or start points "
                               "(every rank must pass the same arguments; draw the seed once and broadcast it)");
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
