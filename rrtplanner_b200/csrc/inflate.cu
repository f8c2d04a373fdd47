// Obstacle inflation for the replanning caller: the frame loop of
// DynamicEnvironmentAnimation.simulate_dynamic_goals (anim.py:69-94) dilates the obstacles of every
// frame before it calls set_og + plan:
//
//     og_dilated = binary_dilation(og, iterations=bsize)          anim.py:80   (scipy default structure:
//     buffer_reg = og_dilated - og                                 anim.py:81    the 4-connected cross,
//     buffer_reg[clamp(position + (i, j))] = 0, i, j < 2*bsize     anim.py:82-86  cells outside count as free)
//     og = og | buffer_reg                                         anim.py:87
//
// i.e.  out = og | (dilated & ~hole)  with hole the clamped square [px, px + 2 bsize) x [py, py + 2 bsize).
// Everything happens on the tiled bit grids in HBM, one thread per 32-cell word: a dilation pass ORs a
// word with its two y-shifts (carrying one bit from the words above / below) and the words of rows
// x - 1 and x + 1; bits outside (W, H) are kept clear during the passes and set (obstacle) at the end,
// as every bit grid has them (rrtk.h).
#include "common.cuh"

namespace rrtk {

// word of row x, y-tile ty; zero outside the grid
__device__ __forceinline__ uint32_t grid_word_or_zero(const uint32_t *g, int x, int ty, int W, int TY)
{
    if (x < 0 || x >= W || ty < 0 || ty >= TY) return 0u;
    return g[(((x >> 5) * TY + ty) << 5) | (x & 31)];
}

// bits of a word that lie inside the grid (rows beyond W have none)
__device__ __forceinline__ uint32_t inside_mask(int x, int ty, int W, int H)
{
    if (x >= W) return 0u;
    const int valid = H - ty * 32;
    return valid >= 32 ? 0xffffffffu : (valid <= 0 ? 0u : ((1u << valid) - 1u));
}

// one 4-connected dilation pass; `first` strips the padding bits of the source
__global__ void dilate_pass_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int nworlds, int W, int H, bool first)
{
    const int TX = tiles_x(W), TY = tiles_y(H);
    const size_t words_per = (size_t)TX * TY * 32;
    const size_t total = words_per * nworlds;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int world = (int)(t / words_per);
        const int local = (int)(t - (size_t)world * words_per);
        const int xl = local & 31, tile = local >> 5;
        const int ty = tile % TY, x = (tile / TY) * 32 + xl;
        const uint32_t *g = in + (size_t)world * words_per;
        auto ld = [&](int xx, int tty) {
            uint32_t w = grid_word_or_zero(g, xx, tty, W, TY);
            if (first) w &= inside_mask(xx, tty, W, H);
            return w;
        };
        const uint32_t c = ld(x, ty);
        uint32_t r = c | (c << 1) | (c >> 1) | (ld(x, ty - 1) >> 31) | (ld(x, ty + 1) << 31) | ld(x - 1, ty) | ld(x + 1, ty);
        out[(size_t)world * words_per + local] = r & inside_mask(x, ty, W, H);
    }
}

// out[o] = og | (dilated & ~hole[o]) of source world o % nworlds, padding bits set again; hole[o] = (px, py, size)
__global__ void inflate_merge_kernel(const uint32_t *__restrict__ og, const uint32_t *__restrict__ dil, const int *__restrict__ holes,
                                     uint32_t *__restrict__ out, int nworlds, int nout, int W, int H)
{
    const int TX = tiles_x(W), TY = tiles_y(H);
    const size_t words_per = (size_t)TX * TY * 32;
    const size_t total = words_per * nout;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int world = (int)(t / words_per);
        const int local = (int)(t - (size_t)world * words_per);
        const size_t src = (size_t)(world % nworlds) * words_per + local;
        const int xl = local & 31, tile = local >> 5;
        const int ty = tile % TY, x = (tile / TY) * 32 + xl;
        uint32_t buffer = dil[src];
        if (holes) {
            const int px = holes[3 * world], py = holes[3 * world + 1], sz = holes[3 * world + 2];
            if (sz > 0 && x >= px && x <= min(px + sz - 1, W - 1)) {
                // y range [py, min(py + sz - 1, H - 1)] intersected with this word's [32 ty, 32 ty + 31]
                const int lo = max(py, 32 * ty) - 32 * ty, hi = min(min(py + sz - 1, H - 1), 32 * ty + 31) - 32 * ty;
                if (lo <= hi) {
                    const uint32_t m = (hi - lo == 31) ? 0xffffffffu : (((1u << (hi - lo + 1)) - 1u) << lo);
                    buffer &= ~m;
                }
            }
        }
        out[t] = og[src] | buffer | ~inside_mask(x, ty, W, H);
    }
}

int inflate_launch(const uint32_t *d_bits, int nworlds, int W, int H, int iterations, const int32_t *d_holes, int nout,
                   uint32_t *d_out, uint32_t *d_scratch, cudaStream_t st)
{
    const size_t total = grid_words(W, H) * nworlds;
    if (total == 0 || nout == 0) return RRTK_OK;
    const int threads = 256;
    size_t blocks = (total + threads - 1) / threads;
    if (blocks > 148 * 64) blocks = 148 * 64;
    // ping-pong between the two halves of d_scratch (2 * nworlds grids): the merge reads the last pass there
    const uint32_t *src = d_bits;
    uint32_t *a = d_scratch, *b = d_scratch + total;
    for (int it = 0; it < iterations; ++it) {
        dilate_pass_kernel<<<(unsigned)blocks, threads, 0, st>>>(src, a, nworlds, W, H, it == 0);
        src = a;
        uint32_t *tmp = a; a = b; b = tmp;
    }
    if (iterations == 0) {
        // binary_dilation(iterations=0) in scipy repeats until nothing changes; the caller never asks for it
        // (bsize = int(movespeed / 2) >= 1 for any useful speed), so 0 means "no buffer": dilated = og
        src = d_bits;
    }
    const size_t total_out = grid_words(W, H) * nout;
    size_t mblocks = (total_out + threads - 1) / threads;
    if (mblocks > 148 * 64) mblocks = 148 * 64;
    inflate_merge_kernel<<<(unsigned)mblocks, threads, 0, st>>>(d_bits, src, iterations == 0 ? nullptr : d_holes, d_out, nworlds, nout, W, H);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

}  // namespace rrtk
