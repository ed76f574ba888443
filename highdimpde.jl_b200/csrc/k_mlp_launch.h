// k_mlp_launch.h — per-(element type, group size) launch table of the MLP kernels.
// Each (T, G) pair is compiled in its own translation unit (k_mlp_inst.cu with
// -DHDP_INST_T / -DHDP_INST_G) so the build parallelises.
#pragma once
#include <string.h>
#include <utility>
#include "k_mlp.cuh"

namespace hdp {

// Launch of one link of a solve's chain: programmatic stream serialisation, so that the link's blocks may be
// scheduled while the previous link still runs and wait at `griddepcontrol.wait` (chain_enter, k_mlp.cuh)
// instead of being launched only after it has drained.  (-DHDP_NO_CHAIN_PDL: plain launches, for A/B runs.)
inline void chained_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* at, dim3 grid, dim3 block, size_t smem, cudaStream_t st)
{
    memset(cfg, 0, sizeof(*cfg));
    cfg->gridDim = grid;
    cfg->blockDim = block;
    cfg->dynamicSmemBytes = smem;
    cfg->stream = st;
#ifndef HDP_NO_CHAIN_PDL
    at->id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at->val.programmaticStreamSerializationAllowed = 1;
    cfg->attrs = at;
    cfg->numAttrs = 1;
#endif
}
template <typename... KA, typename... A>
inline cudaError_t launch_chained(void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args)
{
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at;
    chained_config(&cfg, &at, grid, block, smem, st);
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);
}
// the same for a kernel of an NVRTC-built module
inline cudaError_t launch_chained_module(const void* kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, void** args)
{
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at;
    chained_config(&cfg, &at, grid, block, smem, st);
    return cudaLaunchKernelExC(&cfg, kernel, args);
}

template <typename T> struct MlpLaunchers {
    // [0]: one vec-group per pass (short rows), [1]: three per pass (long rows), [2]: four per pass (very long
    // rows: >= 48 vec-groups per lane, measured in profiles/r2c_*, r2j_*); second index: the mode of the
    // coordinate loop the kernel was compiled for (em_mode_of: 0 general, EM_NOBC, EM_SYM, EM_BOX)
    int (*max_blocks_per_sm[3][4])(int block, size_t dyn_smem);
    cudaError_t (*units[3][4])(const MlpParams<T>& P, int grid, int block, size_t dyn_smem, cudaStream_t st);
    cudaError_t (*expand)(const MlpParams<T>& P, const int32_t* list, int count, int count_hits, const int32_t* list2, int count2,
                          int count_hits2, int grid, int block, cudaStream_t st);
};

#define HDP_DECL_INST(T, TN, G) MlpLaunchers<T> mlp_launchers_##TN##_g##G();
#define HDP_FOR_ALL_G(X, T, TN) X(T, TN, 1) X(T, TN, 2) X(T, TN, 4) X(T, TN, 8) X(T, TN, 16) X(T, TN, 32)
HDP_FOR_ALL_G(HDP_DECL_INST, double, f64)
HDP_FOR_ALL_G(HDP_DECL_INST, float, f32)

} // namespace hdp
