// user_kinds.h — NVRTC build of the K3 kernels for *_USER kinds.
#pragma once
#include <mutex>
#include <string>
#include <vector>
#include "../../include/hdpde.h"

namespace hdp {

struct UserBuild {
    std::vector<char> cubin;
    std::string log;
    std::string units_name, expand_name, combine_name; // lowered (mangled) kernel names
};

int user_needed(const hdp_problem* p);
std::string user_key(const hdp_problem* p, const hdp_user_exprs* u, bool f32, int G);
// HDP_OK, HDP_ERR_BAD_ARG (bad expression: log has the compiler output) or HDP_ERR_NVRTC
int user_compile(const hdp_problem* p, const hdp_user_exprs* u, bool f32, int G, UserBuild* out);

// K1 (k_sde.cuh): sde_loop_kernel<T, AL, INJ, K1_GENERIC> and em_ensemble_kernel<T, AL, INJ, TRAJ>
// for user μ/σ; lowered names in the order sde[AL][INJ] (4), ens[AL][INJ][TRAJ] (8)
struct SdeUserBuild {
    std::vector<char> cubin;
    std::string log;
    std::vector<std::string> names;
};
int user_compile_sde(const hdp_problem* p, const hdp_user_exprs* u, bool f32, SdeUserBuild* out);

} // namespace hdp
