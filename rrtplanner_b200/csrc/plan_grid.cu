// Instantiations and launch of the bucket form of K7 (plan_grid.cuh): RRTStandard and RRTStar, K = 8, 128 threads.
#include "plan_grid.cuh"

namespace rrtk {

template <int KIND, int K, bool KEY32>
static int grid_launch_k(const PlanParams &P, int nplans, size_t smem, cudaStream_t st)
{
    auto kern = plan_grid_kernel<KIND, K, 128, KEY32>;
    RRTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nplans, 128, smem, st>>>(P);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

template <int KIND, int K, bool KEY32>
static int grid_occupancy_k(size_t smem)
{
    auto kern = plan_grid_kernel<KIND, K, 128, KEY32>;
    int blocks = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, 128, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return blocks;
}

template <int KIND, int K>
static int grid_launch_kk(bool key32, const PlanParams &P, int nplans, size_t smem, cudaStream_t st)
{
    return key32 ? grid_launch_k<KIND, K, true>(P, nplans, smem, st) : grid_launch_k<KIND, K, false>(P, nplans, smem, st);
}

template <int KIND, int K>
static int grid_occupancy_kk(bool key32, size_t smem)
{
    return key32 ? grid_occupancy_k<KIND, K, true>(smem) : grid_occupancy_k<KIND, K, false>(smem);
}

int grid_launch(int kind, int K, const PlanParams &P, int nplans, size_t smem, cudaStream_t st)
{
    const bool k32 = P.g_kb > 0;
    if (K == 16) return kind == RRTK_STANDARD ? grid_launch_kk<RRTK_STANDARD, 16>(k32, P, nplans, smem, st) : grid_launch_kk<RRTK_STAR, 16>(k32, P, nplans, smem, st);
    return kind == RRTK_STANDARD ? grid_launch_kk<RRTK_STANDARD, 8>(k32, P, nplans, smem, st) : grid_launch_kk<RRTK_STAR, 8>(k32, P, nplans, smem, st);
}

int grid_occupancy(int kind, int K, bool key32, size_t smem)
{
    if (K == 16) return kind == RRTK_STANDARD ? grid_occupancy_kk<RRTK_STANDARD, 16>(key32, smem) : grid_occupancy_kk<RRTK_STAR, 16>(key32, smem);
    return kind == RRTK_STANDARD ? grid_occupancy_kk<RRTK_STANDARD, 8>(key32, smem) : grid_occupancy_kk<RRTK_STAR, 8>(key32, smem);
}

}  // namespace rrtk
