// K0 (bit-packing), the free-space index used by the sampler, and the synthetic world generator.
#include "common.cuh"

namespace rrtk {

// ---- K0: (W,H) uint8 grid -> tiled bit grid --------------------------------------------------
// One thread builds one 32-cell word.  Consecutive lanes take consecutive y-tiles of the same row,
// so a warp reads 1 KB of contiguous bytes; `og != 0` is the reference's obstacle test (rrt.py:218).
__global__ void pack_kernel(const uint8_t *__restrict__ og, int nworlds, int W, int H, uint32_t *__restrict__ bits)
{
    const int TX = tiles_x(W), TY = tiles_y(H);
    const size_t words_per = (size_t)TX * TY * 32;
    const size_t total = words_per * nworlds;
    const bool vec = (H % 16) == 0;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int world = (int)(t / words_per);
        const size_t local = t - (size_t)world * words_per;
        // thread order: ty fastest, then x (padded to 32*TX)
        const int ty = (int)(local % TY);
        const int x = (int)(local / TY);
        uint32_t word = 0xffffffffu;                   // outside the grid = obstacle
        if (x < W) {
            const uint8_t *row = og + ((size_t)world * W + x) * H + (size_t)ty * 32;
            const int valid = min(32, H - ty * 32);
            word = 0;
            if (vec && valid == 32) {
                const uint4 a = __ldg(reinterpret_cast<const uint4 *>(row));
                const uint4 b = __ldg(reinterpret_cast<const uint4 *>(row) + 1);
                const uint32_t w8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
#pragma unroll
                    for (int b4 = 0; b4 < 4; ++b4)
                        if ((w8[k] >> (8 * b4)) & 0xffu) word |= 1u << (4 * k + b4);
                }
            } else {
                for (int k = 0; k < valid; ++k)
                    if (__ldg(row + k)) word |= 1u << k;
                if (valid < 32) word |= 0xffffffffu << valid;
            }
        }
        bits[(size_t)world * words_per + ((((x >> 5) * TY + ty) << 5) | (x & 31))] = word;
    }
}

int pack_launch(const uint8_t *d_og, int nworlds, int W, int H, uint32_t *d_bits, cudaStream_t st)
{
    const size_t total = grid_words(W, H) * nworlds;
    const int threads = 256;
    size_t blocks = (total + threads - 1) / threads;
    if (blocks > 148 * 64) blocks = 148 * 64;
    pack_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_og, nworlds, W, H, d_bits);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

// inverse of K0: tiled bit grid -> (W,H) uint8 0/1 grid (for callers that want a device-made grid back)
__global__ void unpack_kernel(const uint32_t *__restrict__ bits, int nworlds, int W, int H, uint8_t *__restrict__ og)
{
    const int TY = tiles_y(H);
    const size_t cells = (size_t)W * H, total = cells * nworlds, words_per = grid_words(W, H);
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int world = (int)(t / cells);
        const size_t c = t - (size_t)world * cells;
        const int x = (int)(c / H), y = (int)(c - (size_t)x * H);
        og[t] = (bits[(size_t)world * words_per + word_index(x, y, TY)] >> (y & 31)) & 1u;
    }
}

int unpack_launch(const uint32_t *d_bits, int nworlds, int W, int H, uint8_t *d_og, cudaStream_t st)
{
    const size_t total = (size_t)W * H * nworlds;
    if (total == 0) return RRTK_OK;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    unpack_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_bits, nworlds, W, H, d_og);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

// ---- free-space index: rowcum[x] = number of free cells in rows < x ----------------------------
// The reference's sampler indexes free = argwhere(og == 0) (rrt.py:64), which lists free cells
// row-major (x, then y); rank -> cell is therefore "row by prefix count, then y by in-row rank".
__global__ void free_rows_kernel(const uint32_t *__restrict__ bits, int W, int H, int *__restrict__ rowcum)
{
    const int world = blockIdx.x;
    const int TY = tiles_y(H);
    const uint32_t *g = bits + (size_t)world * grid_words(W, H);
    int *out = rowcum + (size_t)world * (W + 1);
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        int c = 0;
        for (int ty = 0; ty < TY; ++ty) c += __popc(~g[word_index(x, ty << 5, TY)]);
        out[x + 1] = c;
    }
    __syncthreads();
    if (threadIdx.x < 32) {      // one warp turns counts into an exclusive prefix, 32 rows a step
        int carry = 0;
        for (int base = 0; base < W; base += 32) {
            const int x = base + threadIdx.x;
            int v = x < W ? out[x + 1] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(RRTK_FULL, v, o);
                if ((int)threadIdx.x >= o) v += t;
            }
            if (x < W) out[x + 1] = carry + v;
            carry += __shfl_sync(RRTK_FULL, v, 31);
        }
        if (threadIdx.x == 0) out[0] = 0;
    }
}

int free_rows_launch(const uint32_t *d_bits, int nworlds, int W, int H, int32_t *d_rowcum, cudaStream_t st)
{
    free_rows_kernel<<<nworlds, 256, 0, st>>>(d_bits, W, H, d_rowcum);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

// ---- synthetic worlds: integer value-noise fBm, bit-identical to rrtplanner_b200/worlds.py -------
__device__ __forceinline__ uint32_t lattice_hash(uint32_t ix, uint32_t iy, uint32_t seed)
{
    uint32_t u = ix * 0x9E3779B1u + iy * 0x85EBCA77u + seed * 0xC2B2AE3Du;
    u ^= u >> 15; u *= 0x2C1B3C6Du;
    u ^= u >> 12; u *= 0x297A2D39u;
    u ^= u >> 15;
    return u;
}
__device__ __forceinline__ long long smooth_q16(long long t)
{
    const long long t2 = (t * t) >> 16;
    return (t2 * (3 * 65536 - 2 * t)) >> 16;
}
__device__ __forceinline__ long long octave_q16(long long xq, long long yq, uint32_t seed)
{
    const uint32_t ix = (uint32_t)(xq >> 16), iy = (uint32_t)(yq >> 16);
    const long long fx = xq & 65535, fy = yq & 65535;
    const long long sx = smooth_q16(fx), sy = smooth_q16(fy);
    const long long v00 = lattice_hash(ix, iy, seed) >> 16, v10 = lattice_hash(ix + 1, iy, seed) >> 16;
    const long long v01 = lattice_hash(ix, iy + 1, seed) >> 16, v11 = lattice_hash(ix + 1, iy + 1, seed) >> 16;
    const long long a = (v00 * (65536 - sx) + v10 * sx) >> 16;
    const long long b = (v01 * (65536 - sx) + v11 * sx) >> 16;
    return (a * (65536 - sy) + b * sy) >> 16;
}

__global__ void world_field_kernel(const int *__restrict__ seeds, int W, int H, int *__restrict__ field, int *__restrict__ minmax)
{
    const int world = blockIdx.y;
    const int cells = W * H;
    const uint32_t seed = (uint32_t)seeds[world];
    int lo = 0x7fffffff, hi = -0x7fffffff;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += gridDim.x * blockDim.x) {
        const int x = c / H, y = c - x * H;
        const long long xq = (long long)x * 1311, yq = (long long)y * 1311;
        long long total = 0;
#pragma unroll
        for (int o = 0; o < 3; ++o)
            total += octave_q16(xq << o, yq << o, seed * 31u + (uint32_t)o * 7919u) << (2 - o);
        const int v = (int)total;
        field[(size_t)world * cells + c] = v;
        lo = min(lo, v);
        hi = max(hi, v);
    }
    lo = __reduce_min_sync(RRTK_FULL, lo);
    hi = __reduce_max_sync(RRTK_FULL, hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&minmax[2 * world], lo);
        atomicMax(&minmax[2 * world + 1], hi);
    }
}
__global__ void world_init_kernel(int nworlds, int *minmax)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < nworlds) { minmax[2 * w] = 0x7fffffff; minmax[2 * w + 1] = -0x7fffffff; }
}
__global__ void world_thresh_kernel(const int *__restrict__ field, const int *__restrict__ minmax, int cells, int permille,
                                    uint8_t *__restrict__ og)
{
    const int world = blockIdx.y;
    const long long lo = minmax[2 * world], hi = minmax[2 * world + 1];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += gridDim.x * blockDim.x) {
        const long long v = field[(size_t)world * cells + c];
        og[(size_t)world * cells + c] = ((v - lo) * 1000 < (long long)permille * (hi - lo)) ? 1 : 0;   // oggen.py:41-44
    }
}

int worlds_launch(const int32_t *d_seeds, int nworlds, int W, int H, int permille, int32_t *d_scratch, uint8_t *d_og,
                  cudaStream_t st)
{
    const int cells = W * H;
    int *field = d_scratch;
    int *minmax = d_scratch + (size_t)nworlds * cells;
    world_init_kernel<<<(nworlds + 127) / 128, 128, 0, st>>>(nworlds, minmax);
    dim3 grid((cells + 255) / 256 > 148 ? 148 : (cells + 255) / 256, nworlds);
    world_field_kernel<<<grid, 256, 0, st>>>(d_seeds, W, H, field, minmax);
    world_thresh_kernel<<<grid, 256, 0, st>>>(field, minmax, cells, permille, d_og);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

}  // namespace rrtk
