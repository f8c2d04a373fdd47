// Instantiations and launch of the bucket form of K7 (plan_grid.cuh): RRTStandard and RRTStar, K = 8, 128 threads.
#include "plan_grid.cuh"

namespace rrtk {

template <int KIND, int K>
static int grid_launch_k(const PlanParams &P, int nplans, size_t smem, cudaStream_t st)
{
    auto kern = plan_grid_kernel<KIND, K, 128>;
    RRTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nplans, 128, smem, st>>>(P);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

template <int KIND, int K>
static int grid_occupancy_k(size_t smem)
{
    auto kern = plan_grid_kernel<KIND, K, 128>;
    int blocks = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, 128, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return blocks;
}

int grid_launch(int kind, int K, const PlanParams &P, int nplans, size_t smem, cudaStream_t st)
{
    if (K == 16) return kind == RRTK_STANDARD ? grid_launch_k<RRTK_STANDARD, 16>(P, nplans, smem, st) : grid_launch_k<RRTK_STAR, 16>(P, nplans, smem, st);
    return kind == RRTK_STANDARD ? grid_launch_k<RRTK_STANDARD, 8>(P, nplans, smem, st) : grid_launch_k<RRTK_STAR, 8>(P, nplans, smem, st);
}

int grid_occupancy(int kind, int K, size_t smem)
{
    if (K == 16) return kind == RRTK_STANDARD ? grid_occupancy_k<RRTK_STANDARD, 16>(smem) : grid_occupancy_k<RRTK_STAR, 16>(smem);
    return kind == RRTK_STANDARD ? grid_occupancy_k<RRTK_STANDARD, 8>(smem) : grid_occupancy_k<RRTK_STAR, 8>(smem);
}

}  // namespace rrtk
