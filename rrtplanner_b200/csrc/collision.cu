// K1: RRT.collisionfree (rrt.py:183-229) for a batch of segments, one warp per segment.
//
// Lane l tests cell 32*c + l of the reference's walk through the closed form in common.cuh, so a
// chunk of 32 cells costs one grid-word load per lane and one ballot.  In the tiled bit layout a
// chunk touches at most a handful of 128-byte tiles whatever its direction, so grids that do not
// fit shared memory (2048^2 = 512 KB) are served from L1/L2 through the read-only path; grids that
// do fit are staged in shared memory once per block.
#include "common.cuh"

namespace rrtk {

template <class Grid>
__device__ __forceinline__ void walk_segments(const Grid &g, int TY, const int4 *__restrict__ segs, int64_t nseg,
                                              uint8_t *__restrict__ free_out, int *__restrict__ cells_out, int64_t first,
                                              int64_t stride, int lane)
{
    for (int64_t s = first; s < nseg; s += stride) {
        const int4 e = __ldg(segs + s);
        const int r = warp_first_hit(g, TY, e.x, e.y, e.z, e.w, lane);
        if (lane == 0) {
            free_out[s] = r < 0;
            if (cells_out) cells_out[s] = cells_tested(r);
        }
    }
}

// single world, grid read through L1/L2
__global__ void collision_global_kernel(const uint32_t *__restrict__ bits, int TY, const int4 *__restrict__ segs, int64_t nseg,
                                        uint8_t *__restrict__ free_out, int *__restrict__ cells_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    GlobalGrid g{bits};
    walk_segments(g, TY, segs, nseg, free_out, cells_out, warp, nwarps, lane);
}

// single world, grid staged in shared memory
__global__ void collision_shared_kernel(const uint32_t *__restrict__ bits, int words, int TY, const int4 *__restrict__ segs,
                                        int64_t nseg, uint8_t *__restrict__ free_out, int *__restrict__ cells_out)
{
    extern __shared__ __align__(16) uint32_t s_grid[];
    for (int i = threadIdx.x; i < words / 4; i += blockDim.x)
        reinterpret_cast<uint4 *>(s_grid)[i] = __ldg(reinterpret_cast<const uint4 *>(bits) + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    SharedGrid g{s_grid};
    walk_segments(g, TY, segs, nseg, free_out, cells_out, warp, nwarps, lane);
}

// per-segment world index (many small worlds): always through L1/L2
__global__ void collision_multi_kernel(const uint32_t *__restrict__ bits, size_t words_per, int TY, const int4 *__restrict__ segs,
                                       const int *__restrict__ world, int64_t nseg, uint8_t *__restrict__ free_out,
                                       int *__restrict__ cells_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = warp; s < nseg; s += nwarps) {
        const int4 e = __ldg(segs + s);
        GlobalGrid g{bits + (size_t)__ldg(world + s) * words_per};
        const int r = warp_first_hit(g, TY, e.x, e.y, e.z, e.w, lane);
        if (lane == 0) {
            free_out[s] = r < 0;
            if (cells_out) cells_out[s] = cells_tested(r);
        }
    }
}

int collision_launch(const uint32_t *d_bits, int W, int H, const int32_t *d_segs, const int32_t *d_world, int64_t nseg,
                     uint8_t *d_free, int32_t *d_cells, int sm_count, int optin, cudaStream_t st)
{
    if (nseg == 0) return RRTK_OK;
    const int TY = tiles_y(H);
    const size_t words = grid_words(W, H);
    const int threads = 256;
    const int64_t warps_needed = nseg;
    int64_t blocks = (warps_needed * 32 + threads - 1) / threads;
    const int4 *segs = reinterpret_cast<const int4 *>(d_segs);
    if (d_world) {
        const int64_t cap = (int64_t)sm_count * 8;
        if (blocks > cap) blocks = cap;
        collision_multi_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_bits, words, TY, segs, d_world, nseg, d_free, d_cells);
    } else if (words * 4 + 1024 <= (size_t)optin / 2 && nseg >= 4096) {
        // two blocks per SM keep 16 warps in flight per SM while the grid stays on chip
        const size_t smem = words * 4;
        int per_sm = (int)((size_t)(optin + 1024) / (smem + 1024));
        if (per_sm > 4) per_sm = 4;
        if (per_sm < 1) per_sm = 1;
        const int th = per_sm >= 4 ? 256 : (per_sm >= 2 ? 512 : 1024);
        int64_t cap = (int64_t)sm_count * per_sm;
        blocks = (warps_needed * 32 + th - 1) / th;
        if (blocks > cap) blocks = cap;
        RRTK_CUDA(cudaFuncSetAttribute(collision_shared_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        collision_shared_kernel<<<(unsigned)blocks, th, smem, st>>>(d_bits, (int)words, TY, segs, nseg, d_free, d_cells);
    } else {
        const int64_t cap = (int64_t)sm_count * 8;     // 8 x 256 threads = 64 resident warps per SM
        if (blocks > cap) blocks = cap;
        collision_global_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_bits, TY, segs, nseg, d_free, d_cells);
    }
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

}  // namespace rrtk
