// hdp_context.h — host-side handle of libhdpde_b200.so: stream, scratch,
// error reporting, optional NCCL communicator (dlopen'ed), MLP plan cache.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/hdpde.h"

namespace hdp {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct MlpPlan;     // mlp_plan.h
struct MlpSharding; // mlp_plan.h

// K3 kernels compiled by NVRTC for one set of user expressions (user_kinds.cu)
struct UserModule {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t units = nullptr, expand = nullptr, combine = nullptr;
};

// K1 kernels (k_sde.cuh) compiled by NVRTC for one pair of user μ/σ expressions
struct SdeUserModule {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t sde[2][2] = {};    // [ALIGNED][INJECT], BC = K1_GENERIC
    cudaKernel_t ens[2][2][2] = {}; // [ALIGNED][INJECT][TRAJ]
};

} // namespace hdp

struct hdp_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t evp[3] = {nullptr, nullptr, nullptr}; // phase boundaries inside hdp_mlp_solve_*
    cudaDeviceProp prop{};
    std::mutex mu;
    std::string err;
    // grow-only scratch, indexed by role
    enum { B_A = 0, B_B, B_C, B_D, B_E, B_F, B_G, B_H, B_I, B_J, B_K, B_L, B_M, B_O, B_P, B_Q, B_R, B_N };
    hdp::DevBuf buf[B_N];
    int* d_flags = nullptr;          // [0] start point outside the box, [1] kernel/mode mismatch, [2] peer exchange timed out
    unsigned long long* d_ctr = nullptr; // [0] work counter, [1] bounce counter
    // NCCL (optional)
    void* nccl_comm = nullptr;
    int nccl_rank = 0, nccl_nranks = 1;
    // peer exchange window (optional, hdp_peer_window_*): this rank's window and every rank's as mapped here
    void* peer_win = nullptr;
    void* peer_base[HDP_PEER_MAX_RANKS] = {};
    int peer_nranks = 0, peer_rank = 0;
    unsigned long long peer_epoch = 0; // one per exchange, in step on all ranks (every exchange is collective)
    // MLP plan cache
    std::map<std::string, std::shared_ptr<hdp::MlpPlan>> plans;
    std::map<std::string, std::shared_ptr<hdp::MlpSharding>> shardings; // per (plan, nranks, rank)
    std::map<std::string, hdp::UserModule> user_modules;
    std::map<std::string, hdp::SdeUserModule> sde_user_modules;
    // hdp_set_user_exprs: μ/σ expressions for the K1 entry points (copies)
    std::string user_mu, user_sigma;
    bool has_user_mu = false, has_user_sigma = false;
    int launches = 0;
};

namespace hdp {

int set_err(hdp_context* h, int code, const char* fmt, ...);
// grow-only device buffer; returns nullptr (and sets h->err) on failure
void* ensure(hdp_context* h, int which, size_t bytes);

#define HDP_CUDA(h, call)                                                                      \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return hdp::set_err((h), HDP_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                                cudaGetErrorString(e__));                                      \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// NCCL through dlopen (no link-time dependency; reuses an already-loaded libnccl.so.2)
int nccl_allreduce_f64_device(hdp_context* h, double* dbuf, size_t n);

} // namespace hdp
