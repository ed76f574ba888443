// mlp_plan_pod.h — the POD records of the tree plan shared by host planner and device
// kernels (also compiled by NVRTC for user kinds: no host headers here).
#pragma once
#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

namespace hdp {

constexpr int kMaxL = 6; // == HDP_MAX_LEVELS (include/hdpde.h)

enum { UNIT_LEVEL = 0, UNIT_G = 1, UNIT_WHOLE = 2 };
// class of a site in a sharded solve (MlpSharding::site_cls): replicated above the ownership roots / at or below an
// ownership root of this rank / of another rank
enum { SITE_UPPER = 0, SITE_MINE = 1, SITE_OTHER = 2 };
constexpr int kStageBlocks = 64; // most blocks that share the pre-reduction of one node (mlp_prereduce_kernel)

struct PlanUnit {
    int32_t node;
    int32_t kind;     // UNIT_*
    int32_t l;        // level of the sites in the slice (UNIT_LEVEL)
    int32_t _pad;
    int64_t k0, k1;   // site range [k0,k1) within the level (or g-term range)
    uint64_t c, r, s; // consumption ordinals at the first site: path-step / time / sampler
};

struct PlanNode {
    int32_t lv;        // call level: this node is U_lv(x, s)
    int32_t point;     // index of x in the per-solve point table
    int32_t time_site; // site whose r is this node's start time s (-1: root, s = t0)
    int32_t depth;
    int32_t unit_begin, unit_end; // units of this node (creation order)
    int32_t site_begin, site_end; // explicit sites of this node
};

struct PlanSite {
    int32_t node;      // parent call
    int32_t l;         // level of this site
    int32_t x2_point;  // point index of x2
    int32_t x3_point0; // first of K' consecutive x3 points (-1 if local)
    int32_t child0;    // index into children[]: [U_l(x2), U_{l-1}(x2), {U_l(x3_h), U_{l-1}(x3_h)}_h];
                       // entry >= 0: explicit node (value[node]); -1: identically 0;
                       // <= -2: a whole-call leaf, its value is partial[-2 - entry] (no combine needed)
    int32_t depth;     // depth of the parent node
    uint64_t c, r, s;  // ordinals at the start of the site
};


} // namespace hdp
