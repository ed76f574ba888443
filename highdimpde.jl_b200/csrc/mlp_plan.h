// mlp_plan.h — host-side planner that cuts the multilevel-Picard recursion
// tree (src/MLP.jl:102-208) into GPU work units.
//
// The geometry of the tree (sample times, points) never depends on a value of
// u — only the f(...) combinations do — and every random draw is addressed by
// its consumption ordinal in the reference's depth-first order.  So the tree
// can be cut anywhere: the top is expanded into explicit *nodes* (calls
// U_lv(x,s)) and *sites* (one (l,k) iteration of the level loops); everything
// below the cut is evaluated on-chip, depth-first, by one lane group per
// *unit*.  A call's value is a plain sum over its sites, so a unit is a slice
// of one node's site list (or a whole small call); the nonlinearity f is
// applied where complete child values meet, in the bottom-up combine.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "mlp_plan_pod.h"

namespace hdp {

struct MlpCountTables {
    int M, L, K, Kp; // Kp = K if a sampler integrates, else 0
    uint64_t C[kMaxL + 1], R[kMaxL + 1], S[kMaxL + 1]; // per call U_l
    uint64_t F[kMaxL + 1], G[kMaxL + 1];
    uint64_t cs[kMaxL + 1], rs[kMaxL + 1], ss[kMaxL + 1]; // per level-l site
    uint64_t Mpow[kMaxL + 2];
    double W[kMaxL + 1];  // work model of a call
    double Ws[kMaxL + 1]; // work model of a level-l site
};

void mlp_count_tables(int M, int L, int K, bool nonlocal, MlpCountTables* t);

struct MlpPlan {
    MlpCountTables tab;
    double thresh = 0;
    int compat_threads = 0;
    std::vector<PlanNode> nodes;
    std::vector<PlanSite> sites;
    std::vector<PlanUnit> units;
    std::vector<int32_t> children;
    std::vector<float> unit_cost;       // work model of each unit (path-steps + half the sampler calls)
    std::vector<int8_t> unit_tag;       // (kind << 4) | level, the only unit fields the combine pass reads
    std::vector<int32_t> exec_order;    // unit ids, most expensive first
    std::vector<int32_t> sites_by_depth;  // site ids grouped by depth
    std::vector<int32_t> site_depth_off;  // size max_depth+2
    std::vector<int32_t> nodes_by_depth;  // nodes that need a combine (not whole-call leaves)
    std::vector<int32_t> node_depth_off;
    int n_points = 1;
    int max_depth = 0;
    // device copies (owned; freed in destructor)
    void* d_nodes = nullptr;
    void* d_sites = nullptr;
    void* d_units = nullptr;
    void* d_children = nullptr;
    void* d_unit_tag = nullptr;
    void* d_exec = nullptr;
    void* d_sites_by_depth = nullptr;
    void* d_nodes_by_depth = nullptr;
    ~MlpPlan();
};

// thresh: maximum work (path-steps + sampler calls) of one unit.
// compat_threads > 0 reproduces the `_ml_picard_mlt` / `_get_loop_num`
// enumeration of the root call (src/MLP.jl:218-404).
// max_units > 0: give up (return false, plan unusable) as soon as the cut needs more units than that.
bool mlp_build_plan(int M, int L, int K, bool nonlocal, double thresh, int compat_threads, MlpPlan* plan,
                    size_t max_units = 0);

int64_t mlp_get_loop_num(int M, int64_t num, int thread_id, int nthreads);

// One solve (or a small batch) sharded over `nranks` GPUs — the analogue of the reference's
// `Threads.@spawn` / `fetch` / `sum` (src/MLP.jl:235-259), cut where a GPU needs it.
//
// f is nonlinear, so a site can only be finished where all its child values are complete.  The plan is
// therefore split into a *crown* — the root, the *upper* sites and the calls directly below them — which every
// rank holds, and the subtrees below the *ownership roots*: walking down from the root, the first site whose
// modelled cost is at most 1/64 of a rank's share (or below which no call has explicit sites).  Each ownership
// root belongs to ONE rank: that rank generates the geometry of everything below it (deeper explicit sites
// included), evaluates its units, combines its calls and applies f.  What crosses the wire is linear: per crown
// node, this rank's share of the node's value (its slice units, the f-terms of its ownership roots; the value of
// a whole-call leaf below an upper site) — one double per crown node and solve, summed by a single
// ncclAllReduce.  The upper sites are then finished redundantly.
struct MlpSharding {
    int nranks = 1, rank = 0;
    std::vector<int32_t> exec_own;            // units this rank evaluates, most expensive first
    std::vector<int32_t> upper_sites;         // replicated sites grouped by depth
    std::vector<int32_t> upper_depth_off;     // size max_depth+2
    std::vector<int32_t> own_sites;           // this rank's ownership roots and the explicit sites below them, grouped by depth
    std::vector<int32_t> own_depth_off;       // size max_depth+2
    std::vector<int32_t> owned_nodes;         // calls below this rank's ownership roots, grouped by depth (combine: CMB_ALL)
    std::vector<int32_t> owned_depth_off;     // size max_depth+2
    std::vector<int32_t> crown_nodes;         // every crown node, leaves included (combine: CMB_PARTIAL)
    std::vector<int32_t> crown_leaves;        // whole-call leaves below upper sites (scatter after the all-reduce)
    std::vector<int32_t> crown_final;         // crown nodes that are not leaves, grouped by depth (combine: CMB_FINAL)
    std::vector<int32_t> crown_final_off;     // size max_depth+2
    std::vector<int8_t> unit_mine;            // per unit
    std::vector<int8_t> site_cls;             // per site: 0 upper, 1 at or below an ownership root of mine, 2 of another rank
    std::vector<int32_t> node_slot;           // per node: slot in the exchange buffer, -1 below an ownership root
    int n_slots = 0;
    // crown partial pass: node crown_nodes[k] is summed by crown_job_off[k + 1] - crown_job_off[k] blocks (its units
    // and this rank's sites shared out among them; 0 for a whole-call leaf), job j = (k, share) — k_mlp_host.cu
    std::vector<int32_t> crown_job_off;       // size crown_nodes.size() + 1
    std::vector<int32_t> crown_jobs;          // pairs (k, share)
    int crown_heavy = 0;                      // crown nodes with jobs
    double load_max = 0, load_mean = 0;       // modelled work of the busiest rank / the mean
    // device copies (owned by the handle's cache; freed in the destructor)
    void *d_exec_own = nullptr, *d_upper_sites = nullptr, *d_own_sites = nullptr, *d_owned_nodes = nullptr,
         *d_crown_nodes = nullptr, *d_crown_leaves = nullptr, *d_crown_final = nullptr, *d_unit_mine = nullptr,
         *d_site_cls = nullptr, *d_node_slot = nullptr, *d_crown_jobs = nullptr, *d_crown_job_off = nullptr;
    ~MlpSharding();
};
// own_items: ownership roots per rank to aim for (0: the default, 64; tests lower it to nest explicit sites below a root)
void mlp_build_sharding(const MlpPlan& plan, int nranks, int rank, MlpSharding* out, double own_items = 0);
// Shares of the crown partial pass (crown_job_off, crown_jobs, crown_heavy): about `per_share` units per block of
// 256 threads — a site of this rank counts as sixteen: 1 + K evaluations of f behind three dependent loads —, at most
// kStageBlocks per node; a node of at most `light_below` such units gets none and is walked by its warp of the
// combine launch (C4: 4312 crown nodes of a few dozen sites).  (Tests lower both to force shares on small trees.)
void mlp_build_crown_jobs(const MlpPlan& plan, MlpSharding* sh, long light_below = 512, long per_share = 2048);

} // namespace hdp
