// Instantiations and launch of the bucket form of K7 (plan_grid.cuh): RRTStandard and RRTStar, K = 8, 128 threads.
#include "plan_grid.cuh"

namespace rrtk {

template <int KIND>
static int grid_launch_k(const PlanParams &P, int nplans, size_t smem, cudaStream_t st)
{
    auto kern = plan_grid_kernel<KIND, 8, 128>;
    RRTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nplans, 128, smem, st>>>(P);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

template <int KIND>
static int grid_occupancy_k(size_t smem)
{
    auto kern = plan_grid_kernel<KIND, 8, 128>;
    int blocks = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, 128, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return blocks;
}

int grid_launch(int kind, const PlanParams &P, int nplans, size_t smem, cudaStream_t st)
{
    return kind == RRTK_STANDARD ? grid_launch_k<RRTK_STANDARD>(P, nplans, smem, st) : grid_launch_k<RRTK_STAR>(P, nplans, smem, st);
}

int grid_occupancy(int kind, size_t smem)
{
    return kind == RRTK_STANDARD ? grid_occupancy_k<RRTK_STANDARD>(smem) : grid_occupancy_k<RRTK_STAR>(smem);
}

}  // namespace rrtk
