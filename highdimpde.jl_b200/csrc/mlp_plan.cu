// mlp_plan.cu — host-only planner (see mlp_plan.h).  Work-count recurrences:
//   C(0)=0;  C(L) = Σ_{l<L} M^{L-l}·(1 + (1+K')·(C(l) + [l≥2]·C(l-1))) + M^L
// and the analogous R (time uniforms) and S (sampler calls), which give the
// closed-form consumption ordinal of every node of the reference's recursion.
#include "mlp_plan.h"
#include <math.h>
#include <algorithm>
#include <numeric>

namespace hdp {

void mlp_count_tables(int M, int L, int K, bool nonlocal, MlpCountTables* t)
{
    t->M = M; t->L = L; t->K = K; t->Kp = nonlocal ? K : 0;
    const uint64_t Kp = (uint64_t)t->Kp;
    t->Mpow[0] = 1;
    for (int i = 1; i <= kMaxL + 1; ++i) t->Mpow[i] = t->Mpow[i - 1] * (uint64_t)M;
    for (int i = 0; i <= kMaxL; ++i) {
        t->C[i] = t->R[i] = t->S[i] = t->F[i] = t->G[i] = 0;
        t->cs[i] = t->rs[i] = t->ss[i] = 0;
        t->W[i] = t->Ws[i] = 0;
    }
    for (int lv = 0; lv <= L && lv <= kMaxL; ++lv) {
        // per-site consumption at level lv (needs C[lv], C[lv-1], filled below for lv>=1)
        if (lv >= 1) {
            uint64_t c = 0, r = 0, s = 0, f = 0, g = 0;
            for (int l = 0; l < lv; ++l) {
                uint64_t num = t->Mpow[lv - l];
                c += num * t->cs[l];
                r += num * t->rs[l];
                s += num * t->ss[l];
                uint64_t ff = t->F[l] + (l >= 2 ? t->F[l - 1] : 0);
                uint64_t gg = t->G[l] + (l >= 2 ? t->G[l - 1] : 0);
                f += num * ((uint64_t)K * (l >= 2 ? 2 : 1) + (1 + Kp) * ff);
                g += num * ((1 + Kp) * gg);
            }
            t->C[lv] = c + t->Mpow[lv];
            t->R[lv] = r;
            t->S[lv] = s;
            t->F[lv] = f;
            t->G[lv] = g + t->Mpow[lv];
        }
        uint64_t cc = t->C[lv] + (lv >= 2 ? t->C[lv - 1] : 0);
        uint64_t rr = t->R[lv] + (lv >= 2 ? t->R[lv - 1] : 0);
        uint64_t sc = t->S[lv] + (lv >= 2 ? t->S[lv - 1] : 0);
        t->cs[lv] = 1 + (1 + Kp) * cc;
        t->rs[lv] = 1 + (1 + Kp) * rr;
        t->ss[lv] = Kp + (1 + Kp) * sc;
        t->W[lv] = (double)t->C[lv] + 0.5 * (double)t->S[lv];
        t->Ws[lv] = (double)t->cs[lv] + 0.5 * (double)t->ss[lv];
    }
}

// `_get_loop_num(M, num, thread_id, NUM_THREADS)`  (src/MLP.jl:391-404); `1:loop_num`
// with a Float64 loop_num iterates floor(loop_num) times.
int64_t mlp_get_loop_num(int M, int64_t num, int thread_id, int nthreads)
{
    if (num < nthreads) return thread_id > num ? 0 : 1;
    int64_t remainder = (int64_t)M % num;
    double ln = (double)num / (double)nthreads;
    if (remainder > 0 && thread_id <= remainder) ln += 1.0;
    return (int64_t)floor(ln);
}

namespace {
constexpr double kMinExplicitCall = 64.0, kMinExplicitCallNonlocal = 16.0, kLongSite = 6.0;
constexpr double kOwnItems = 64.0; // ownership roots per rank a sharded solve aims for (at least)
struct Builder {
    MlpPlan* p;
    const MlpCountTables& t;
    double thresh;
    int compat;
    std::vector<double> unit_cost;
    size_t max_units = 0; // 0: no limit
    bool overflow = false;

    void add_unit(int node, int kind, int l, int64_t k0, int64_t k1, uint64_t c, uint64_t r, uint64_t s, double cost)
    {
        if (max_units && p->units.size() >= max_units) { overflow = true; return; }
        PlanUnit u{};
        u.node = node; u.kind = kind; u.l = l; u.k0 = k0; u.k1 = k1; u.c = c; u.r = r; u.s = s;
        p->units.push_back(u);
        unit_cost.push_back(cost);
    }

    // one run of `nsites` consecutive level-l sites of node `id`, starting at (c,r,s)
    void do_level(int id, int depth, int l, int64_t nsites, uint64_t& c, uint64_t& r, uint64_t& s)
    {
        if (nsites <= 0) return;
        // Sites become explicit (their child calls separate units, f applied in the combine pass) only where the
        // child call is worth a launch-level detour: below kMinExplicitCall path-steps (U_1: 2M steps) the
        // geometry and combine passes cost more than the finer units buy (measured: Allen-Cahn d=100, M=8, L=5
        // one solve 0.99 -> 0.73 ms; and the C4 plan one level deeper has 16 M units), so such sites stay in
        // slice units, one site per unit at the finest.  A non-local site fans out into 1 + K child calls per
        // level, so there the detour pays from smaller calls on (replicator-mutator d=10, L=4: 1.16 vs 2.69 ms).
        // A site whose child calls each fit a unit can still be far too long for one lane to walk: a non-local site
        // is 1 + K' (or twice that) child calls, and with fewer such sites than resident lanes the units phase lasts
        // as long as one of them however many GPUs share the solve (C4 on 8 ranks: 57 k level-1 sites of 89
        // path-steps per rank, 0.7 ms each, for 114 k lanes).  Such a site becomes explicit too; its children are
        // then whole-call units, read straight from partial[] by the combine pass.
        const bool long_site = t.Ws[l] > kLongSite * thresh;
        if (l >= 1 && (t.W[l] > thresh || long_site) &&
            t.W[l] >= (t.Kp > 0 ? kMinExplicitCallNonlocal : kMinExplicitCall)) {
            for (int64_t k = 0; k < nsites; ++k) {
                PlanSite st{};
                st.node = id; st.l = l; st.depth = depth;
                st.c = c; st.r = r; st.s = s;
                st.x2_point = p->n_points++;
                st.x3_point0 = -1;
                if (t.Kp > 0) { st.x3_point0 = p->n_points; p->n_points += t.Kp; }
                st.child0 = (int32_t)p->children.size();
                p->children.resize(p->children.size() + 2 * (size_t)(1 + t.Kp), -1);
                p->sites.push_back(st);
                c += t.cs[l]; r += t.rs[l]; s += t.ss[l];
            }
        } else {
            int64_t chunk = (int64_t)floor(thresh / t.Ws[l]);
            if (chunk < 1) chunk = 1;
            for (int64_t k0 = 0; k0 < nsites; k0 += chunk) {
                int64_t k1 = std::min(nsites, k0 + chunk);
                add_unit(id, UNIT_LEVEL, l, k0, k1, c, r, s, (double)(k1 - k0) * t.Ws[l]);
                c += (uint64_t)(k1 - k0) * t.cs[l];
                r += (uint64_t)(k1 - k0) * t.rs[l];
                s += (uint64_t)(k1 - k0) * t.ss[l];
            }
        }
    }

    // a child that is a single whole-call unit is read straight from partial[] by its parent site
    int32_t child_code(int node) const
    {
        const PlanNode& n = p->nodes[node];
        if (n.unit_end - n.unit_begin == 1 && n.site_end == n.site_begin && p->units[n.unit_begin].kind == UNIT_WHOLE)
            return -2 - n.unit_begin;
        return node;
    }

    int build_node(int lv, int point, int time_site, int depth, uint64_t c0, uint64_t r0, uint64_t s0, bool root)
    {
        const int id = (int)p->nodes.size();
        {
            PlanNode n{};
            n.lv = lv; n.point = point; n.time_site = time_site; n.depth = depth;
            p->nodes.push_back(n);
        }
        p->max_depth = std::max(p->max_depth, depth);
        const int unit_begin = (int)p->units.size();
        const int site_begin = (int)p->sites.size();
        const bool compat_root = root && compat > 0;
        if (t.W[lv] <= thresh && !compat_root) {
            add_unit(id, UNIT_WHOLE, lv, 0, 1, c0, r0, s0, t.W[lv]);
        } else {
            uint64_t c = c0, r = r0, s = s0;
            if (!compat_root) {
                for (int l = 0; l < lv; ++l) do_level(id, depth, l, (int64_t)t.Mpow[lv - l], c, r, s);
            } else {
                // `_ml_picard_mlt`: tasks tid = 1..NT in turn, each its levels (src/MLP.jl:235-243, 262-388)
                for (int tid = 1; tid <= compat; ++tid)
                    for (int l = 0; l < lv; ++l)
                        do_level(id, depth, l, mlp_get_loop_num(t.M, (int64_t)t.Mpow[lv - l], tid, compat), c, r, s);
            }
            // g-term: M^lv plain steps (src/MLP.jl:198-205 / :246-254)
            int64_t ng = (int64_t)t.Mpow[lv];
            int64_t chunk = (int64_t)floor(thresh);
            if (chunk < 1) chunk = 1;
            for (int64_t k0 = 0; k0 < ng; k0 += chunk) {
                int64_t k1 = std::min(ng, k0 + chunk);
                add_unit(id, UNIT_G, 0, k0, k1, c, r, s, (double)(k1 - k0));
                c += (uint64_t)(k1 - k0);
            }
        }
        const int unit_end = (int)p->units.size();
        const int site_end = (int)p->sites.size();
        p->nodes[id].unit_begin = unit_begin; p->nodes[id].unit_end = unit_end;
        p->nodes[id].site_begin = site_begin; p->nodes[id].site_end = site_end;
        // children of the explicit sites (consumption order inside a site:
        // r, step, U_l(x2), [U_{l-1}(x2)], then per h: sample, U_l(x3), [U_{l-1}(x3)])
        for (int si = site_begin; si < site_end && !overflow; ++si) {
            const PlanSite st = p->sites[si];
            const int l = st.l;
            uint64_t c = st.c + 1, r = st.r + 1, s = st.s;
            int ch = st.child0;
            int id0 = build_node(l, st.x2_point, si, depth + 1, c, r, s, false);
            p->children[ch++] = child_code(id0);
            c += t.C[l]; r += t.R[l]; s += t.S[l];
            if (l >= 2) {
                int id1 = build_node(l - 1, st.x2_point, si, depth + 1, c, r, s, false);
                p->children[ch++] = child_code(id1);
                c += t.C[l - 1]; r += t.R[l - 1]; s += t.S[l - 1];
            } else ch++;
            for (int h = 0; h < t.Kp; ++h) {
                s += 1; // the sampler call itself
                int idh = build_node(l, st.x3_point0 + h, si, depth + 1, c, r, s, false);
                p->children[ch++] = child_code(idh);
                c += t.C[l]; r += t.R[l]; s += t.S[l];
                if (l >= 2) {
                    int idh1 = build_node(l - 1, st.x3_point0 + h, si, depth + 1, c, r, s, false);
                    p->children[ch++] = child_code(idh1);
                    c += t.C[l - 1]; r += t.R[l - 1]; s += t.S[l - 1];
                } else ch++;
            }
        }
        return id;
    }
};
} // namespace

bool mlp_build_plan(int M, int L, int K, bool nonlocal, double thresh, int compat_threads, MlpPlan* plan, size_t max_units)
{
    mlp_count_tables(M, L, K, nonlocal, &plan->tab);
    plan->thresh = thresh;
    plan->compat_threads = compat_threads;
    plan->nodes.clear(); plan->sites.clear(); plan->units.clear(); plan->children.clear();
    plan->n_points = 1;
    plan->max_depth = 0;
    Builder b{plan, plan->tab, thresh, compat_threads, {}};
    b.max_units = max_units;
    b.build_node(L, 0, -1, 0, 0, 0, 0, true);
    if (b.overflow) return false; // the cut is finer than the caller can hold: try a coarser one
    plan->unit_cost.assign(b.unit_cost.begin(), b.unit_cost.end());
    plan->unit_tag.resize(plan->units.size());
    for (size_t i = 0; i < plan->units.size(); ++i)
        plan->unit_tag[i] = (int8_t)((plan->units[i].kind << 4) | (plan->units[i].l & 15));
    // execution order: most expensive first (LPT); ties keep creation order so that
    // equal-shaped units stay adjacent (lane groups of one warp run in lockstep)
    plan->exec_order.resize(plan->units.size());
    std::iota(plan->exec_order.begin(), plan->exec_order.end(), 0);
    std::stable_sort(plan->exec_order.begin(), plan->exec_order.end(),
                     [&](int a, int c) { return b.unit_cost[a] > b.unit_cost[c]; });
    // depth buckets
    const int nd = plan->max_depth + 1;
    plan->site_depth_off.assign(nd + 1, 0);
    plan->node_depth_off.assign(nd + 1, 0);
    for (auto& s : plan->sites) plan->site_depth_off[s.depth + 1]++;
    auto needs_combine = [&](size_t i) {
        const PlanNode& n = plan->nodes[i];
        const bool leaf = n.unit_end - n.unit_begin == 1 && n.site_end == n.site_begin &&
                          plan->units[n.unit_begin].kind == UNIT_WHOLE;
        return i == 0 || !leaf;
    };
    for (size_t i = 0; i < plan->nodes.size(); ++i)
        if (needs_combine(i)) plan->node_depth_off[plan->nodes[i].depth + 1]++;
    for (int i = 0; i < nd; ++i) {
        plan->site_depth_off[i + 1] += plan->site_depth_off[i];
        plan->node_depth_off[i + 1] += plan->node_depth_off[i];
    }
    plan->sites_by_depth.resize(plan->sites.size());
    plan->nodes_by_depth.resize((size_t)plan->node_depth_off[nd]);
    std::vector<int32_t> so(plan->site_depth_off.begin(), plan->site_depth_off.end() - 1);
    std::vector<int32_t> no(plan->node_depth_off.begin(), plan->node_depth_off.end() - 1);
    for (size_t i = 0; i < plan->sites.size(); ++i) plan->sites_by_depth[so[plan->sites[i].depth]++] = (int32_t)i;
    for (size_t i = 0; i < plan->nodes.size(); ++i)
        if (needs_combine(i)) plan->nodes_by_depth[no[plan->nodes[i].depth]++] = (int32_t)i;
    return true;
}


// ---------------------------------------------------------------------
// sharding of one plan over several ranks (see MlpSharding in mlp_plan.h)
// ---------------------------------------------------------------------
void mlp_build_sharding(const MlpPlan& plan, int nranks, int rank, MlpSharding* out, double own_items)
{
    const int n_nodes = (int)plan.nodes.size(), n_sites = (int)plan.sites.size(), n_units = (int)plan.units.size();
    const int Kp = plan.tab.Kp;
    const int nch = 2 * (1 + Kp);
    out->nranks = nranks; out->rank = rank;
    auto is_leaf = [&](int n) {
        const PlanNode& nd = plan.nodes[n];
        return nd.unit_end - nd.unit_begin == 1 && nd.site_end == nd.site_begin && plan.units[nd.unit_begin].kind == UNIT_WHOLE;
    };
    // explicit child nodes of a site: the entries of children[] that name a node; a whole-call leaf is named by
    // its unit there (-2 - unit), so map units of leaves back to their node
    std::vector<int32_t> leaf_of_unit(n_units, -1);
    for (int n = 0; n < n_nodes; ++n)
        if (is_leaf(n)) leaf_of_unit[plan.nodes[n].unit_begin] = n;
    auto child_node = [&](int32_t code) -> int { return code >= 0 ? code : (code == -1 ? -1 : leaf_of_unit[-2 - code]); };
    // Ownership roots: walking down from the root, the first site that is small enough to be one item of the
    // balance (at most 1/kOwnItems of a rank's share: the longest-first assignment below is then within a
    // few per cent of even), or below which no call has explicit sites.  Everything under such a site —
    // deeper explicit sites included — belongs to its owner; the sites above are the upper sites.
    const double own_cap = plan.tab.W[plan.tab.L] / ((double)nranks * (own_items > 0 ? own_items : kOwnItems));
    std::vector<int8_t> structural_bottom(n_sites, 1);
    for (int si = 0; si < n_sites; ++si)
        for (int q = 0; q < nch; ++q) {
            const int32_t code = plan.children[plan.sites[si].child0 + q];
            if (code >= 0 && plan.nodes[code].site_end > plan.nodes[code].site_begin) structural_bottom[si] = 0;
        }
    std::vector<int32_t> root_of_node(n_nodes, -1); // ownership root above a node; -1: crown node
    std::vector<int32_t> root_of_site(n_sites, -1); // ownership root of a site (itself for a root); -1: upper site
    for (int si = 0; si < n_sites; ++si) { // sites are created parent-first: root_of_node[] of the parent call is final here
        const PlanSite& st = plan.sites[si];
        const int above = root_of_node[st.node];
        if (above >= 0) root_of_site[si] = above;
        else if (structural_bottom[si] || plan.tab.Ws[st.l] <= own_cap) root_of_site[si] = si;
        for (int q = 0; q < nch; ++q) {
            const int cn = child_node(plan.children[st.child0 + q]);
            if (cn >= 0) root_of_node[cn] = root_of_site[si];
        }
    }
    // items to balance: ownership roots (with everything below them) and the units attached to crown nodes
    struct Item { double cost; int kind, id; }; // kind 0: ownership root, 1: crown-attached unit
    std::vector<Item> items;
    for (int si = 0; si < n_sites; ++si)
        if (root_of_site[si] == si) items.push_back({plan.tab.Ws[plan.sites[si].l], 0, si});
    std::vector<int32_t> node_of_unit(n_units, 0);
    for (int n = 0; n < n_nodes; ++n)
        for (int u = plan.nodes[n].unit_begin; u < plan.nodes[n].unit_end; ++u) node_of_unit[u] = n;
    for (int u = 0; u < n_units; ++u)
        if (root_of_node[node_of_unit[u]] < 0) items.push_back({(double)plan.unit_cost[u], 1, u});
    std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.cost > b.cost; });
    std::vector<double> load(nranks, 0.0);
    std::vector<int32_t> root_owner(n_sites, -1), unit_owner(n_units, -1);
    int rr = 0;
    for (const Item& it : items) { // longest first, each to the least loaded rank (ties: round robin)
        int best = rr % nranks;
        for (int q = 0; q < nranks; ++q) {
            const int r = (rr + q) % nranks;
            if (load[r] < load[best]) best = r;
        }
        ++rr;
        load[best] += it.cost;
        if (it.kind == 0) root_owner[it.id] = best;
        else unit_owner[it.id] = best;
    }
    for (int u = 0; u < n_units; ++u)
        if (unit_owner[u] < 0) unit_owner[u] = root_owner[root_of_node[node_of_unit[u]]];
    out->load_max = *std::max_element(load.begin(), load.end());
    out->load_mean = std::accumulate(load.begin(), load.end(), 0.0) / nranks;
    // per-rank lists
    out->unit_mine.assign(n_units, 0);
    out->exec_own.clear();
    for (int32_t u : plan.exec_order)
        if (unit_owner[u] == rank) { out->exec_own.push_back(u); out->unit_mine[u] = 1; }
    const int nd = plan.max_depth + 1;
    out->site_cls.assign(n_sites, 0);
    out->upper_depth_off.assign(nd + 1, 0);
    out->own_depth_off.assign(nd + 1, 0);
    out->upper_sites.clear();
    out->own_sites.clear();
    for (int dep = 0; dep < nd; ++dep) {
        for (int q = plan.site_depth_off[dep]; q < plan.site_depth_off[dep + 1]; ++q) {
            const int si = plan.sites_by_depth[q];
            const int root = root_of_site[si];
            out->site_cls[si] = root < 0 ? 0 : (root_owner[root] == rank ? 1 : 2);
            if (root < 0) out->upper_sites.push_back(si);
            else if (root_owner[root] == rank) out->own_sites.push_back(si);
        }
        out->upper_depth_off[dep + 1] = (int32_t)out->upper_sites.size();
        out->own_depth_off[dep + 1] = (int32_t)out->own_sites.size();
    }
    out->node_slot.assign(n_nodes, -1);
    out->crown_nodes.clear(); out->crown_leaves.clear();
    out->n_slots = 0;
    std::vector<std::vector<int32_t>> fin(nd), own(nd);
    for (int n = 0; n < n_nodes; ++n) {
        if (root_of_node[n] < 0) {
            out->node_slot[n] = out->n_slots++;
            out->crown_nodes.push_back(n);
            if (is_leaf(n)) out->crown_leaves.push_back(n);
            if (!is_leaf(n) || n == 0) fin[plan.nodes[n].depth].push_back(n); // (a root that is one unit still gets its value)
        } else if (!is_leaf(n) && root_owner[root_of_node[n]] == rank) own[plan.nodes[n].depth].push_back(n);
    }
    out->crown_final.clear();
    out->owned_nodes.clear();
    out->crown_final_off.assign(nd + 1, 0);
    out->owned_depth_off.assign(nd + 1, 0);
    for (int dep = 0; dep < nd; ++dep) {
        out->crown_final.insert(out->crown_final.end(), fin[dep].begin(), fin[dep].end());
        out->crown_final_off[dep + 1] = (int32_t)out->crown_final.size();
        out->owned_nodes.insert(out->owned_nodes.end(), own[dep].begin(), own[dep].end());
        out->owned_depth_off[dep + 1] = (int32_t)out->owned_nodes.size();
    }
}

void mlp_build_crown_jobs(const MlpPlan& plan, MlpSharding* sh, long light_below, long per_share)
{
    sh->crown_jobs.clear();
    sh->crown_job_off.assign(1, 0);
    sh->crown_heavy = 0;
    for (size_t k = 0; k < sh->crown_nodes.size(); ++k) {
        const PlanNode& nd = plan.nodes[sh->crown_nodes[k]];
        const int nu = nd.unit_end - nd.unit_begin;
        const bool leaf = nu == 1 && nd.site_end == nd.site_begin && (plan.unit_tag[nd.unit_begin] >> 4) == UNIT_WHOLE;
        int nb = 0;
        if (!leaf) {
            long work = nu;
            for (int si = nd.site_begin; si < nd.site_end; ++si)
                if (sh->site_cls[si] == SITE_MINE) work += 16;
            nb = work <= light_below ? 0 : (int)std::min<long>(std::max<long>((work + per_share - 1) / per_share, 1), kStageBlocks);
        }
        sh->crown_heavy += nb > 0;
        for (int b = 0; b < nb; ++b) { sh->crown_jobs.push_back((int32_t)k); sh->crown_jobs.push_back(b); }
        sh->crown_job_off.push_back((int32_t)(sh->crown_jobs.size() / 2));
    }
}


} // namespace hdp
