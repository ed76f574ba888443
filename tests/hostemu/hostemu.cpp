// hostemu.cpp — TEST INFRASTRUCTURE ONLY: runs the product's MLP device code
// (k_mlp.cuh, compiled for the host through hdp_host_emu_shim.h with G = 1) and
// the product's planner (mlp_plan.cu) on the CPU.  Used by tests/ to check the
// tree cut / ordinal / combine logic against the oracle without a GPU.
// It is NOT a fallback: libhdpde_b200.so never contains or calls it.
#define HDP_HOST_EMU 1
#include <vector>
#include <algorithm>
#include "k_mlp.cuh"
#include "k_sde.cuh"
#include "mlp_plan.h"
#include "hdp_convert.h"

using namespace hdp;

namespace hdp {
MlpPlan::~MlpPlan() {} // no device copies in the emulation
}

namespace hdp {
MlpSharding::~MlpSharding() {}
}
// ownership roots per rank the emulated shardings aim for (0: the library's default); tests lower it so that
// explicit sites nest below an ownership root on small trees
static double g_own_items = 0;
extern "C" void hdp_emu_set_own_items(double v) { g_own_items = v; }
// shares of the crown partial pass (mlp_build_crown_jobs): the library's values unless a test lowers them to force
// several shares per node on a small tree
static long g_job_light = 512, g_job_share = 2048;
extern "C" void hdp_emu_set_crown_jobs(long light_below, long per_share) { g_job_light = light_below; g_job_share = per_share; }
static long g_jobs_used = 0; // crown partial passes that went through the shares (tests assert the path was taken)
extern "C" long hdp_emu_crown_jobs_used(void) { return g_jobs_used; }

// nranks > 1 emulates a solve sharded over several GPUs (MlpSharding).  In one process the ranks run one after
// the other on private copies of the per-call state, their exchange buffers are summed as the all-reduce would,
// and every rank finishes the crown; the results of all ranks must agree bit for bit (checked), rank 0's is
// returned.  The two phases are also exported separately (hdp_emu_shard_*), so that a world_size-2 `gloo` test can
// run one rank per process and do the all-reduce with torch.distributed.
template <typename T> struct EmuSolve {
    static constexpr int VEC = Tr<T>::VEC;
    static constexpr int G = 1;
    struct RankState {
        std::vector<T> feat, site_r, value, se;
        std::vector<double> partial, rootsq, lvstat, xbuf;
        std::vector<double2> scratch_raw, points_raw;
        unsigned long long ctr[2] = {0, 0};
        int flags[2] = {0, 0};
        MlpSharding sh;
    };
    MlpParams<T> P;
    MlpCountTables tab;
    MlpPlan plan;
    std::vector<T> lohi, ab, x0v;
    std::vector<RankState> rs; // indexed by rank (only the ranks this process emulates are sized)
    Ctx<T, G> c;
    int nranks = 1, n_solves = 1, L = 1, xstride = 0;
    bool want_se = false;

    int setup(const hdp_problem* prob, const hdp_mlp_cfg* cfg, const T* x0, int nranks_, int only_rank)
    {
        const int M = cfg->M, K = cfg->K;
        L = cfg->L; n_solves = cfg->n_solves; nranks = nranks_;
        if (M < 1 || K < 1 || L < 1 || L > kMaxL || n_solves < 1 || nranks < 1) return HDP_ERR_BAD_ARG;
        const bool nonlocal = cfg->sampler.kind != HDP_SAMPLER_NONE;
        memset(&P, 0, sizeof(P));
        fill_dev_problem<T>(prob, &P.pb);
        fill_dev_sampler<T>(&cfg->sampler, &P.sm);
        const int d = prob->d;
        x0v.assign(x0, x0 + (size_t)cfg->n_x0 * d);
        if (prob->has_bc && !prob->bc_is_scalar) {
            lohi.resize(2 * d);
            for (int i = 0; i < d; ++i) { lohi[i] = (T)prob->lo[i]; lohi[d + i] = (T)prob->hi[i]; }
            P.pb.lo = lohi.data(); P.pb.hi = lohi.data() + d;
        }
        if (cfg->sampler.kind == HDP_SAMPLER_UNIFORM && !cfg->sampler.bounds_are_scalar) {
            ab.resize(2 * d);
            for (int i = 0; i < d; ++i) { ab[i] = (T)cfg->sampler.a[i]; ab[d + i] = (T)cfg->sampler.b[i]; }
            P.sm.a = ab.data(); P.sm.b = ab.data() + d;
        }
        mlp_count_tables(M, L, K, nonlocal, &tab);
        const int compat = cfg->compat_mlt_threads > 0 ? cfg->compat_mlt_threads : 0;
        double thresh = tab.W[L];
        if (cfg->target_units > 0) thresh = std::max(1.0, tab.W[L] * n_solves / (double)cfg->target_units);
        if (thresh > tab.W[L]) thresh = tab.W[L];
        mlp_build_plan(M, L, K, nonlocal, thresh, compat, &plan);
        const int nvec = (d + VEC - 1) / VEC;
        P.M = M; P.K = K; P.L = L; P.Kp = tab.Kp;
        P.nvec = nvec; P.npl = nvec; P.dpad = nvec * VEC;
        for (int i = 0; i <= kMaxL; ++i) {
            P.C[i] = tab.C[i]; P.R[i] = tab.R[i]; P.S[i] = tab.S[i];
            P.cs[i] = tab.cs[i]; P.rs[i] = tab.rs[i]; P.ss[i] = tab.ss[i];
        }
        for (int i = 0; i <= kMaxL + 1; ++i) P.Mpow[i] = tab.Mpow[i];
        P.seed0 = cfg->seed; P.n_solves = n_solves; P.n_x0 = cfg->n_x0;
        P.units = plan.units.data(); P.unit_tag = plan.unit_tag.data(); P.nodes = plan.nodes.data();
        P.sites = plan.sites.data(); P.children = plan.children.data();
        P.sites_by_depth = plan.sites_by_depth.data(); P.nodes_by_depth = plan.nodes_by_depth.data();
        P.n_units = (int)plan.units.size(); P.n_nodes = (int)plan.nodes.size();
        P.n_sites = (int)plan.sites.size(); P.n_points = plan.n_points;
        P.nranks = nranks;
        P.r_tape = (const T*)cfg->r_tape; P.n_tape = (const T*)cfg->n_tape; P.s_tape = (const T*)cfg->s_tape;
        want_se = cfg->std_err_out != nullptr;
        const PlanNode& root = plan.nodes[0];
        P.rootsq_stride = 2 * std::max(root.unit_end - root.unit_begin, L + 1);
        rs.resize(nranks);
        for (int r = 0; r < nranks; ++r) {
            if (only_rank >= 0 && r != only_rank) continue;
            RankState& S = rs[r];
            S.feat.resize((size_t)n_solves * P.n_points);
            S.site_r.resize((size_t)n_solves * std::max(P.n_sites, 1));
            S.value.resize((size_t)n_solves * P.n_nodes);
            S.se.resize(n_solves);
            S.partial.assign((size_t)n_solves * P.n_units, 0.0);
            S.rootsq.assign((size_t)n_solves * P.rootsq_stride, 0.0);
            S.lvstat.assign((size_t)n_solves * 3 * (kMaxL + 1), 0.0);
            S.scratch_raw.resize((size_t)(2 * L) * P.npl * VEC * sizeof(T) / sizeof(double2) + 2); // 16-byte aligned
            S.points_raw.resize(((size_t)n_solves * P.n_points * P.dpad * sizeof(T)) / sizeof(double2) + 2);
            if (nranks > 1) {
                mlp_build_sharding(plan, nranks, r, &S.sh, g_own_items);
                mlp_build_crown_jobs(plan, &S.sh, g_job_light, g_job_share);
                xstride = S.sh.n_slots + 3 * (kMaxL + 1);
                S.xbuf.assign((size_t)n_solves * xstride + kXTail, 0.0);
            }
        }
        return HDP_OK;
    }
    size_t xcount() const { return (size_t)n_solves * xstride + kXTail; }
    void bind(int r)
    {
        RankState& S = rs[r];
        P.rank = r;
        P.points = (T*)S.points_raw.data(); P.feat = S.feat.data(); P.site_r = S.site_r.data();
        P.partial = S.partial.data(); P.value = S.value.data(); P.scratch = (T*)S.scratch_raw.data();
        P.total_threads = 1; P.work_ctr = &S.ctr[0]; P.bounce_ctr = &S.ctr[1]; P.flags = S.flags;
        P.rootsq = want_se ? S.rootsq.data() : nullptr;
        P.lvstat = want_se ? S.lvstat.data() : nullptr;
        P.std_err = want_se ? S.se.data() : nullptr;
        if (nranks > 1) {
            P.exec = S.sh.exec_own.data(); P.n_exec = (int)S.sh.exec_own.size();
            P.unit_mine = S.sh.unit_mine.data(); P.site_cls = S.sh.site_cls.data(); P.node_slot = S.sh.node_slot.data();
            P.n_slots = S.sh.n_slots; P.xstride = xstride;
            P.xbuf = S.xbuf.data();
        } else {
            P.exec = plan.exec_order.data(); P.n_exec = P.n_units;
        }
    }
    void expand_list(const int32_t* list, int cnt, bool hits)
    {
        // both forms of the geometry pass (mlp_expand_kernel picks by the size of the pass): short lists one item per
        // (site, point), longer ones one item per site
        const bool split = cnt < 8;
        for (u64 it = 0; it < (u64)cnt * n_solves * (u64)(split ? 1 + P.Kp : 1); ++it) { c.nb = 0; expand_item<T, G>(c, P, list, it, hits, split); }
    }
    void combine_list(const int32_t* list, int cnt, int mode)
    {
        for (u64 it = 0; it < (u64)cnt * n_solves; ++it) combine_item<T>(P, list, it, 0, 1, 0xffffffffu, mode);
    }
    // geometry, units and the local combine of one rank; fills its exchange buffer
    void phase1(int r)
    {
        bind(r);
        RankState& S = rs[r];
        for (u64 q = 0; q < (u64)n_solves * P.dpad; ++q) init_item<T>(P, x0v.data(), q);
        ctx_init<T, G>(c, P, 0, 0);
        if (nranks == 1) {
            for (int dep = 0; dep <= plan.max_depth; ++dep)
                expand_list(P.sites_by_depth + plan.site_depth_off[dep], plan.site_depth_off[dep + 1] - plan.site_depth_off[dep], true);
        } else {
            for (int dep = 0; dep <= plan.max_depth; ++dep)
                expand_list(S.sh.upper_sites.data() + S.sh.upper_depth_off[dep],
                            S.sh.upper_depth_off[dep + 1] - S.sh.upper_depth_off[dep], r == 0);
            for (int dep = 0; dep <= plan.max_depth; ++dep)
                expand_list(S.sh.own_sites.data() + S.sh.own_depth_off[dep], S.sh.own_depth_off[dep + 1] - S.sh.own_depth_off[dep], true);
        }
        for (u64 w = 0; w < (u64)P.n_exec * n_solves; ++w) unit_item<T, G>(c, P, w);
        if (c.bad) S.flags[0] = 1;
        if (nranks == 1) {
            for (int dep = plan.max_depth; dep >= 0; --dep)
                combine_list(P.nodes_by_depth + plan.node_depth_off[dep], plan.node_depth_off[dep + 1] - plan.node_depth_off[dep], CMB_ALL);
        } else {
            for (int dep = plan.max_depth; dep >= 0; --dep)
                combine_list(S.sh.owned_nodes.data() + S.sh.owned_depth_off[dep],
                             S.sh.owned_depth_off[dep + 1] - S.sh.owned_depth_off[dep], CMB_ALL);
            if (S.sh.crown_heavy >= 2) {
                // the library's crown partial pass by shares (k_mlp_host.cu): every job sums its share of a node's units
                // and of this rank's sites into a record (mlp_prereduce_kernel: here one thread per block), then the
                // combine pass adds the records of a node up; light nodes are walked directly
                const int n_jobs = (int)(S.sh.crown_jobs.size() / 2);
                std::vector<double> stage((size_t)n_solves * n_jobs * kCmbStage, 0.0);
                for (int sv = 0; sv < n_solves; ++sv)
                    for (int j = 0; j < n_jobs; ++j) {
                        const int k = S.sh.crown_jobs[2 * j], share = S.sh.crown_jobs[2 * j + 1];
                        const int nshare = S.sh.crown_job_off[k + 1] - S.sh.crown_job_off[k];
                        const int nidx = S.sh.crown_nodes[k];
                        T Bl[kMaxL], Gs = (T)0;
                        double Ns[kMaxL + 1], Sm[kMaxL + 1], Sq[kMaxL + 1];
                        for (int l = 0; l < kMaxL; ++l) Bl[l] = (T)0;
                        for (int l = 0; l <= kMaxL; ++l) { Ns[l] = 0.0; Sm[l] = 0.0; Sq[l] = 0.0; }
                        prereduce_share<T, true>(P, P.nodes[nidx], nidx, sv, share, nshare, true, CMB_PARTIAL, 0, 1, Bl, Gs, Ns, Sm, Sq);
                        double* q = stage.data() + ((size_t)sv * n_jobs + j) * kCmbStage;
                        for (int l = 0; l < kMaxL; ++l) q[l] = (double)Bl[l];
                        q[kMaxL] = (double)Gs;
                        for (int l = 0; l <= kMaxL; ++l) { q[kMaxL + 1 + 3 * l] = Ns[l]; q[kMaxL + 2 + 3 * l] = Sm[l]; q[kMaxL + 3 + 3 * l] = Sq[l]; }
                    }
                P.stage = stage.data();
                P.stage_off = S.sh.crown_job_off.data();
                P.stage_jobs = n_jobs;
                combine_list(S.sh.crown_nodes.data(), (int)S.sh.crown_nodes.size(), CMB_PARTIAL);
                P.stage = nullptr;
                P.stage_off = nullptr;
                P.stage_jobs = 0;
                ++g_jobs_used;
            } else
                combine_list(S.sh.crown_nodes.data(), (int)S.sh.crown_nodes.size(), CMB_PARTIAL);
            S.xbuf[(size_t)n_solves * xstride + 4] = (double)S.ctr[1];
            S.xbuf[(size_t)n_solves * xstride + 5] = (double)S.flags[0]; // (the library reads its flag per rank; here it rides along)
        }
    }
    // after the all-reduce: scatter the crown leaves, finish the crown
    void phase2(int r, const double* sum)
    {
        bind(r);
        RankState& S = rs[r];
        S.xbuf.assign(sum, sum + xcount());
        P.xbuf = S.xbuf.data();
        for (int32_t nidx : S.sh.crown_leaves)
            for (int sv = 0; sv < n_solves; ++sv)
                S.partial[(size_t)sv * P.n_units + plan.nodes[nidx].unit_begin] = S.xbuf[(size_t)sv * xstride + S.sh.node_slot[nidx]];
        for (int dep = plan.max_depth; dep >= 0; --dep)
            combine_list(S.sh.crown_final.data() + S.sh.crown_final_off[dep],
                         S.sh.crown_final_off[dep + 1] - S.sh.crown_final_off[dep], CMB_FINAL);
        S.ctr[1] = (unsigned long long)sum[(size_t)n_solves * xstride + 4];
        if (sum[(size_t)n_solves * xstride + 5] > 0.0) S.flags[0] = 1;
    }
    int finish(int r, const hdp_mlp_cfg* cfg, T* u_out, hdp_mlp_stats* stats)
    {
        RankState& S = rs[r];
        for (int i = 0; i < n_solves; ++i) u_out[i] = S.value[(size_t)i * P.n_nodes];
        if (want_se)
            for (int i = 0; i < n_solves; ++i) ((T*)cfg->std_err_out)[i] = S.se[i];
        if (stats) {
            memset(stats, 0, sizeof(*stats));
            stats->per_solve.path_steps = tab.C[L];
            stats->per_solve.time_uniforms = tab.R[L];
            stats->per_solve.sampler_calls = tab.S[L];
            stats->per_solve.f_evals = tab.F[L];
            stats->per_solve.g_evals = tab.G[L];
            stats->bounces = S.ctr[1];
            stats->n_units = P.n_units; stats->n_nodes = P.n_nodes; stats->n_sites = P.n_sites;
            stats->group_size = 1;
            stats->exchange_bytes = nranks > 1 ? xcount() * sizeof(double) : 0;
            if (want_se) {
                for (int l = 0; l <= L; ++l) {
                    stats->level_n[l] = (uint64_t)(S.lvstat[3 * l] + 0.5);
                    stats->level_sum[l] = S.lvstat[3 * l + 1];
                    stats->level_sumsq[l] = S.lvstat[3 * l + 2];
                }
                stats->std_error = (double)S.se[0];
            }
        }
        return S.flags[0] ? HDP_ERR_START_OUTSIDE_DOMAIN : HDP_OK;
    }
};

template <typename T>
static int emu_solve(const hdp_problem* prob, const hdp_mlp_cfg* cfg, const T* x0, T* u_out, hdp_mlp_stats* stats,
                     int nranks = 1)
{
    EmuSolve<T> E;
    int rc = E.setup(prob, cfg, x0, nranks, -1);
    if (rc != HDP_OK) return rc;
    for (int r = 0; r < nranks; ++r) E.phase1(r);
    if (nranks > 1) {
        std::vector<double> sum(E.xcount(), 0.0); // the all-reduce
        for (int r = 0; r < nranks; ++r)
            for (size_t i = 0; i < sum.size(); ++i) sum[i] += E.rs[r].xbuf[i];
        for (int r = 0; r < nranks; ++r) E.phase2(r, sum.data());
        for (int r = 1; r < nranks; ++r)
            for (int i = 0; i < E.n_solves; ++i)
                if (E.rs[r].value[(size_t)i * E.P.n_nodes] != E.rs[0].value[(size_t)i * E.P.n_nodes]) return HDP_ERR_UNSUPPORTED;
    }
    return E.finish(0, cfg, u_out, stats);
}

// one rank per process (f64): create + phase 1 -> exchange buffer; all-reduce it outside; phase 2 -> result
extern "C" void* hdp_emu_shard_begin(const hdp_problem* prob, const hdp_mlp_cfg* cfg, const double* x0, int nranks, int rank,
                                     int64_t* xcount)
{
    auto* E = new EmuSolve<double>();
    if (nranks < 2 || rank < 0 || rank >= nranks || E->setup(prob, cfg, x0, nranks, rank) != HDP_OK) { delete E; return nullptr; }
    E->phase1(rank);
    *xcount = (int64_t)E->xcount();
    return E;
}
extern "C" void hdp_emu_shard_xbuf(void* e, int rank, double* out)
{
    auto* E = (EmuSolve<double>*)e;
    std::copy(E->rs[rank].xbuf.begin(), E->rs[rank].xbuf.end(), out);
}
extern "C" int hdp_emu_shard_end(void* e, int rank, const hdp_mlp_cfg* cfg, const double* xsum, double* u_out, hdp_mlp_stats* stats)
{
    auto* E = (EmuSolve<double>*)e;
    E->phase2(rank, xsum);
    int rc = E->finish(rank, cfg, u_out, stats);
    delete E;
    return rc;
}

extern "C" int hdp_emu_mlp_solve_sharded_f64(const hdp_problem* prob, const hdp_mlp_cfg* cfg, const double* x0,
                                             double* u_out, hdp_mlp_stats* stats, int nranks)
{
    return emu_solve<double>(prob, cfg, x0, u_out, stats, nranks);
}
extern "C" int hdp_emu_mlp_solve_sharded_f32(const hdp_problem* prob, const hdp_mlp_cfg* cfg, const float* x0,
                                             float* u_out, hdp_mlp_stats* stats, int nranks)
{
    return emu_solve<float>(prob, cfg, x0, u_out, stats, nranks);
}
// The share table of the crown partial pass (mlp_build_crown_jobs) of one rank, checked structurally:
// returns 0 when every invariant holds, else the number of the first one that fails.
// out: [heavy nodes, jobs, light non-leaf nodes, leaves, units covered by shares, sites covered by shares]
extern "C" int hdp_emu_crown_jobs_check(int M, int L, int K, int nonlocal, double thresh, int nranks, int rank, long light_below,
                                        long per_share, int64_t out[6])
{
    MlpPlan plan;
    mlp_build_plan(M, L, K, nonlocal != 0, thresh, 0, &plan);
    MlpSharding sh;
    mlp_build_sharding(plan, nranks, rank, &sh, g_own_items);
    mlp_build_crown_jobs(plan, &sh, light_below, per_share);
    for (int q = 0; q < 6; ++q) out[q] = 0;
    const size_t nc = sh.crown_nodes.size();
    if (sh.crown_job_off.size() != nc + 1 || sh.crown_job_off[0] != 0) return 1;
    if (sh.crown_jobs.size() != 2 * (size_t)sh.crown_job_off[nc]) return 2;
    out[1] = sh.crown_job_off[nc];
    for (size_t k = 0; k < nc; ++k) {
        const PlanNode& nd = plan.nodes[sh.crown_nodes[k]];
        const int nb = sh.crown_job_off[k + 1] - sh.crown_job_off[k];
        if (nb < 0 || nb > kStageBlocks) return 3;
        const int nu = nd.unit_end - nd.unit_begin, ns = nd.site_end - nd.site_begin;
        const bool leaf = nu == 1 && ns == 0 && (plan.unit_tag[nd.unit_begin] >> 4) == UNIT_WHOLE;
        if (leaf) { if (nb != 0) return 4; out[3]++; continue; }
        if (nb == 0) { out[2]++; continue; }
        out[0]++;
        // the jobs of node k are (k, 0) .. (k, nb - 1), in order
        for (int b = 0; b < nb; ++b) {
            const int j = sh.crown_job_off[k] + b;
            if (sh.crown_jobs[2 * j] != (int32_t)k || sh.crown_jobs[2 * j + 1] != b) return 5;
        }
        // the shares as prereduce_share cuts them tile the node's units and sites exactly once
        std::vector<int> ucover(nu, 0), scover(ns, 0);
        for (int b = 0; b < nb; ++b) {
            const int per = (nu + nb - 1) / nb, u0 = nd.unit_begin + b * per, u1 = std::min(u0 + per, nd.unit_end);
            for (int u = u0; u < u1; ++u) ucover[u - nd.unit_begin]++;
            const int sper = (ns + nb - 1) / nb, s0 = nd.site_begin + b * sper, s1 = std::min(s0 + sper, nd.site_end);
            for (int si = s0; si < s1; ++si) scover[si - nd.site_begin]++;
        }
        for (int c : ucover) if (c != 1) return 6;
        for (int c : scover) if (c != 1) return 7;
        out[4] += nu;
        out[5] += ns;
    }
    if (sh.crown_heavy != (int)out[0]) return 8;
    return 0;
}

// the sharding of a plan, summarised: [n_slots, crown nodes, crown leaves, upper sites, sites at or below an ownership root (all ranks),
// owned nodes (this rank), units (this rank), 1000 * busiest rank's modelled load / mean load]
extern "C" void hdp_emu_sharding_summary(int M, int L, int K, int nonlocal, double thresh, int nranks, int rank, int64_t out[8])
{
    MlpPlan plan;
    mlp_build_plan(M, L, K, nonlocal != 0, thresh, 0, &plan);
    MlpSharding sh;
    mlp_build_sharding(plan, nranks, rank, &sh, g_own_items);
    int64_t bottoms = 0;
    for (int8_t c : sh.site_cls) bottoms += c != 0;
    out[0] = sh.n_slots; out[1] = (int64_t)sh.crown_nodes.size(); out[2] = (int64_t)sh.crown_leaves.size();
    out[3] = (int64_t)sh.upper_sites.size(); out[4] = bottoms; out[5] = (int64_t)sh.owned_nodes.size();
    out[6] = (int64_t)sh.exec_own.size(); out[7] = (int64_t)(1000.0 * sh.load_max / std::max(sh.load_mean, 1e-300));
}

extern "C" int hdp_emu_mlp_solve_f64(const hdp_problem* prob, const hdp_mlp_cfg* cfg, const double* x0,
                                     double* u_out, hdp_mlp_stats* stats)
{
    return emu_solve<double>(prob, cfg, x0, u_out, stats);
}
extern "C" int hdp_emu_mlp_solve_f32(const hdp_problem* prob, const hdp_mlp_cfg* cfg, const float* x0,
                                     float* u_out, hdp_mlp_stats* stats)
{
    return emu_solve<float>(prob, cfg, x0, u_out, stats);
}
// raw draw of the emulated generator (libm transcendentals): for generator sanity tests
extern "C" void hdp_emu_draw_f64(uint64_t seed, int stream, uint64_t ord, uint32_t group, int normal, double out[2])
{
    Vec<double> v = draw_vec<double>(seed, stream, ord, group, normal != 0);
    out[0] = v.v[0]; out[1] = v.v[1];
}
extern "C" void hdp_emu_plan_summary(int M, int L, int K, int nonlocal, double thresh, int compat, int64_t out[8])
{
    MlpPlan plan;
    mlp_build_plan(M, L, K, nonlocal != 0, thresh, compat, &plan);
    out[0] = (int64_t)plan.nodes.size(); out[1] = (int64_t)plan.sites.size(); out[2] = (int64_t)plan.units.size();
    out[3] = plan.n_points; out[4] = plan.max_depth;
    out[5] = (int64_t)plan.tab.C[L]; out[6] = (int64_t)plan.tab.R[L]; out[7] = (int64_t)plan.tab.S[L];
}

// ------------------------------------------------------------------------------------------
// K1 (k_sde.cuh): the `sde_loop!` / `_mlt_sde_loop!` kernel and the EM-ensemble kernel, one emulated
// thread walking the whole grid-stride loop.  The mode is chosen as sde_loop_impl (k_elementwise.cu) does.
// ------------------------------------------------------------------------------------------
template <typename T>
static int emu_sde_loop(const hdp_problem* prob, T* y0, T* y1, int B, int nsteps, T dt, const T* ts,
                        const hdp_sampler* x0s_, uint64_t seed, const T* dW, const hdp_sampler* mc_, T* z, int K,
                        int mlp_assoc, int force_generic, unsigned long long* bounces)
{
    DevProblem<T> pb;
    fill_dev_problem<T>(prob, &pb);
    const int d = prob->d;
    std::vector<T> lohi, ab0, ab1;
    if (prob->has_bc && !prob->bc_is_scalar) {
        lohi.resize(2 * d);
        for (int i = 0; i < d; ++i) { lohi[i] = (T)prob->lo[i]; lohi[d + i] = (T)prob->hi[i]; }
        pb.lo = lohi.data(); pb.hi = lohi.data() + d;
    }
    DevSampler<T> x0s, mcs;
    fill_dev_sampler<T>(x0s_, &x0s);
    fill_dev_sampler<T>(mc_, &mcs);
    auto vec_bounds = [&](const hdp_sampler* s, DevSampler<T>& q, std::vector<T>& ab) {
        if (s && s->kind == HDP_SAMPLER_UNIFORM && !s->bounds_are_scalar) {
            ab.resize(2 * d);
            for (int i = 0; i < d; ++i) { ab[i] = (T)s->a[i]; ab[d + i] = (T)s->b[i]; }
            q.a = ab.data(); q.b = ab.data() + d;
        }
    };
    vec_bounds(x0s_, x0s, ab0);
    vec_bounds(mc_, mcs, ab1);
    const size_t n_el = (size_t)d * (size_t)B;
    if (n_el == 0) return HDP_OK;
    int flags[2] = {0, 0};
    unsigned long long ctr = 0;
    const bool const_coef = prob->mu_kind == HDP_COEF_CONST && prob->sigma_kind == HDP_COEF_CONST && !mlp_assoc && !force_generic;
    int bc = !const_coef ? K1_GENERIC : !pb.has_bc ? K1_NOBC : (pb.bc_is_scalar && pb.lo_s == -pb.hi_s) ? K1_SYM : K1_GENERIC;
    if (bc != K1_GENERIC && pb.mu_c == (T)0) bc += 2;
    T* zz = (z && mcs.kind != HDP_SAMPLER_NONE && K > 0) ? z : nullptr;
#define EMU_K1(INJ, BC) sde_loop_kernel<T, false, INJ, BC>(pb, y0, y1, n_el, nsteps, dt, x0s, nullptr, seed, dW, mcs, zz, K, flags, &ctr, mlp_assoc, 0, ts)
    if (dW) { if (bc == K1_SYM) EMU_K1(true, K1_SYM); else if (bc == K1_NOBC) EMU_K1(true, K1_NOBC); else if (bc == K1_SYM0) EMU_K1(true, K1_SYM0); else if (bc == K1_NOBC0) EMU_K1(true, K1_NOBC0); else EMU_K1(true, K1_GENERIC); }
    else { if (bc == K1_SYM) EMU_K1(false, K1_SYM); else if (bc == K1_NOBC) EMU_K1(false, K1_NOBC); else if (bc == K1_SYM0) EMU_K1(false, K1_SYM0); else if (bc == K1_NOBC0) EMU_K1(false, K1_NOBC0); else EMU_K1(false, K1_GENERIC); }
#undef EMU_K1
    if (bounces) *bounces = ctr;
    return flags[0] ? HDP_ERR_START_OUTSIDE_DOMAIN : HDP_OK;
}

template <typename T>
static int emu_em_ensemble(const hdp_problem* prob, T* x, int B, int nsteps, T dt, T t0, T* traj, uint64_t seed, const T* xi)
{
    DevProblem<T> pb;
    hdp_problem pr = *prob;
    pr.has_bc = 0;
    fill_dev_problem<T>(&pr, &pb);
    const size_t n_el = (size_t)prob->d * (size_t)B;
    if (n_el == 0) return HDP_OK;
    if (xi) { if (traj) em_ensemble_kernel<T, false, true, true>(pb, x, n_el, nsteps, dt, t0, traj, seed, xi, 0);
              else em_ensemble_kernel<T, false, true, false>(pb, x, n_el, nsteps, dt, t0, traj, seed, xi, 0); }
    else { if (traj) em_ensemble_kernel<T, false, false, true>(pb, x, n_el, nsteps, dt, t0, traj, seed, xi, 0);
           else em_ensemble_kernel<T, false, false, false>(pb, x, n_el, nsteps, dt, t0, traj, seed, xi, 0); }
    return HDP_OK;
}

#define EMU_K1_EXPORT(REAL, SFX)                                                                                     \
    extern "C" int hdp_emu_sde_loop##SFX(const hdp_problem* prob, REAL* y0, REAL* y1, int B, int nsteps, REAL dt,    \
                                         const REAL* ts, const hdp_sampler* x0s, uint64_t seed, const REAL* dW,      \
                                         const hdp_sampler* mc, REAL* z, int K, int mlp_assoc, int force_generic,    \
                                         unsigned long long* bounces)                                                \
    {                                                                                                                \
        return emu_sde_loop<REAL>(prob, y0, y1, B, nsteps, dt, ts, x0s, seed, dW, mc, z, K, mlp_assoc, force_generic, \
                                  bounces);                                                                          \
    }                                                                                                                \
    extern "C" int hdp_emu_em_ensemble##SFX(const hdp_problem* prob, REAL* x, int B, int nsteps, REAL dt, REAL t0,   \
                                            REAL* traj, uint64_t seed, const REAL* xi)                               \
    {                                                                                                                \
        return emu_em_ensemble<REAL>(prob, x, B, nsteps, dt, t0, traj, seed, xi);                                    \
    }
EMU_K1_EXPORT(double, _f64)
EMU_K1_EXPORT(float, _f32)

