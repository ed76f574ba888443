// k_mlp_inst.cu — one (T, G) instantiation of the MLP unit / expand kernels.
// Build: nvcc ... -DHDP_INST_T=double -DHDP_INST_TN=f64 -DHDP_INST_G=8 -c k_mlp_inst.cu
#include "k_mlp_launch.h"

namespace hdp {

#define HDP_CAT4_(a, b, c, d) a##b##c##d
#define HDP_CAT4(a, b, c, d) HDP_CAT4_(a, b, c, d)
#define HDP_INST_NAME HDP_CAT4(mlp_launchers_, HDP_INST_TN, _g, HDP_INST_G)

namespace {
typedef HDP_INST_T T_;
constexpr int G_ = HDP_INST_G;

template <int U, int FM> static int inst_max_blocks(int block, size_t dyn_smem)
{
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, mlp_units_kernel<T_, G_, U, FM>, block, dyn_smem) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return nb;
}
template <int U, int FM>
static cudaError_t inst_units(const MlpParams<T_>& P, int grid, int block, size_t dyn_smem, cudaStream_t st)
{
    return launch_chained(mlp_units_kernel<T_, G_, U, FM>, dim3(grid), dim3(block), dyn_smem, st, P);
}
static cudaError_t inst_expand(const MlpParams<T_>& P, const int32_t* list, int count, int count_hits, const int32_t* list2,
                               int count2, int count_hits2, int grid, int block, cudaStream_t st)
{
    return launch_chained(mlp_expand_kernel<T_, G_>, dim3(grid), dim3(block), 0, st, P, list, count, count_hits, list2, count2,
                          count_hits2);
}
} // namespace

MlpLaunchers<HDP_INST_T> HDP_INST_NAME()
{
    MlpLaunchers<HDP_INST_T> l;
    static_assert(EM_NOBC == 1 && EM_SYM == 2 && EM_BOX == 3, "mode index");
#define HDP_SET(LONG, U, FM)                           \
    l.max_blocks_per_sm[LONG][FM] = inst_max_blocks<U, FM>; \
    l.units[LONG][FM] = inst_units<U, FM>;
    HDP_SET(0, 1, 0) HDP_SET(0, 1, 1) HDP_SET(0, 1, 2) HDP_SET(0, 1, 3)
    // the general form always walks one vec-group per pass: long rows only change its launch bounds
    HDP_SET(1, 3, 0) HDP_SET(1, 3, 1) HDP_SET(1, 3, 2) HDP_SET(1, 3, 3)
    HDP_SET(2, 4, 0) HDP_SET(2, 4, 1) HDP_SET(2, 4, 2) HDP_SET(2, 4, 3)
#undef HDP_SET
    l.expand = inst_expand;
    return l;
}

} // namespace hdp
