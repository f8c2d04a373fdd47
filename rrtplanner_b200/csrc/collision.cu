// K1: RRT.collisionfree (rrt.py:183-229) for a batch of segments, one warp per segment.
//
// The reference walks L + 1 = max(|dx|,|dy|) + 1 cells with an integer error accumulator; cell k of
// that walk has the closed form (tests/test_oracle.py::test_closed_form_cell_sequence)
//     major axis:  k steps          minor axis:  q(k) = floor((2*k*minor + major) / (2*major)) steps
// so lane l of a warp tests cell 32*c + l of chunk c and a ballot finds the first occupied one.
//
// The kernel is bound by instruction issue, not by bytes (ncu: profiles/r1_cc_*), so the work is
// arranged to keep the per-chunk instruction count low:
//   * a warp takes 32 consecutive segments: the per-segment constants (direction, major/minor,
//     the per-chunk increment divmod(64*minor, 2*major), a reciprocal) are computed lane-parallel,
//     one segment per lane, from one coalesced 512-byte load, and handed to the walk through shared
//     memory (two warp-uniform LDS.128); results go back with one coalesced store per 32 segments;
//   * the walk is specialised on the major axis.  In the tiled bit layout (rrtk.h) a step of 32 cells
//     along x leaves x & 31 unchanged and moves one tile column, along y it leaves the bit position
//     unchanged and moves one tile row, so the major-axis part of the word index advances by a
//     constant per chunk and only the minor coordinate is re-derived (incrementally, with a carry).
// Grids that fit shared memory are staged there once per block with one TMA bulk copy (SharedGrid); larger ones (2048^2 =
// 512 KB) are read through L1/L2 on the read-only path (GlobalGrid), where a chunk touches at most a
// handful of 128-byte tiles whatever its direction.
#include "common.cuh"

namespace rrtk {

struct __align__(16) SegPre {
    int ax, ay;          // first cell
    int major, minor;    // max / min of (|dx|, |dy|)
    int flags;           // 1: x is the major axis, 2: sx < 0, 4: sy < 0
    int dq, dr;          // (q, r) advance per 32 cells: divmod(64 * minor, 2 * major)
    float inv;           // ~ 1 / (2 * major)
};

__device__ __forceinline__ SegPre seg_prepare(int4 e)
{
    SegPre p;
    const int dx = e.z - e.x, dy = e.w - e.y;
    const int adx = abs(dx), ady = abs(dy);
    const bool xmajor = adx >= ady;
    p.ax = e.x; p.ay = e.y;
    p.major = xmajor ? adx : ady;
    p.minor = xmajor ? ady : adx;
    p.flags = (xmajor ? 1 : 0) | (dx > 0 ? 0 : 2) | (dy > 0 ? 0 : 4);      // reference: sx = +1 iff x0 < x1 (rrt.py:207-215)
    const int den = 2 * p.major;
    p.dq = 0; p.dr = 0;
    if (p.major >= 32) p.dq = small_div(64 * p.minor, den, p.dr);
    float inv = 0.f;
    if (p.major > 0) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(__int2float_rn(den)));
    p.inv = inv;
    return p;
}

// All lanes call this with the same (warp-uniform) SegPre.  Returns k >= 0 = index of the first
// occupied cell, or -(L + 1) when the walk is free; |ret| or ret + 1 = cells the reference reads.
template <bool XMAJOR, class Grid>
__device__ __forceinline__ int walk_axis(const Grid &g, int TY, const SegPre &p, int lane)
{
    const int major = p.major, den = 2 * major;
    const int smaj = XMAJOR ? ((p.flags & 2) ? -1 : 1) : ((p.flags & 4) ? -1 : 1);
    const int smin = XMAJOR ? ((p.flags & 4) ? -1 : 1) : ((p.flags & 2) ? -1 : 1);
    // this lane's first cell: k = lane, q = floor((2*lane*minor + major) / den) via the reciprocal + fix-up
    int q = 0, r = 0;
    if (major > 0) {
        const int num = 2 * lane * p.minor + major;                        // < 63 * 32768
        q = __float2int_rz(__int2float_rn(num) * p.inv);
        r = num - q * den;
        if (r < 0) { r += den; --q; }
        if (r >= den) { r -= den; ++q; }
    }
    int cmaj = (XMAJOR ? p.ax : p.ay) + smaj * lane;
    int cmin = (XMAJOR ? p.ay : p.ax) + smin * q;
    const int dmin = smin * p.dq, dr = p.dr;
    // `left` = major - base (warp-uniform): this lane's cell is on the segment while lane <= left.  Lanes past
    // the end read word 0 (always inside the grid) and ignore it, so the load needs no branch.
    int left = major;
    if (XMAJOR) {
        // word = ((x >> 5) * TY + (y >> 5)) * 32 + (x & 31); x advances by 32 per chunk
        int xw = (((cmaj >> 5) * TY) << 5) | (cmaj & 31);
        const int dxw = smaj * (TY << 5);
        for (;;) {
            const bool on = lane <= left;
            const uint32_t w = g.load(on ? (uint32_t)(xw + (cmin & ~31)) : 0u);
            const unsigned m = __ballot_sync(RRTK_FULL, on && ((w >> (cmin & 31)) & 1u));
            if (m) return major - left + __ffs(m) - 1;
            if (left < 32) break;
            left -= 32; xw += dxw; cmin += dmin; r += dr;
            if (r >= den) { r -= den; cmin += smin; }
        }
    } else {
        // y advances by 32 per chunk: the bit position is fixed, the tile row moves
        const uint32_t bit = 1u << (cmaj & 31);
        int yw = cmaj & ~31;
        const int dyw = smaj * 32;
        for (;;) {
            const bool on = lane <= left;
            const uint32_t w = g.load(on ? (uint32_t)(((((cmin >> 5) * TY) << 5) | (cmin & 31)) + yw) : 0u);
            const unsigned m = __ballot_sync(RRTK_FULL, on && (w & bit));
            if (m) return major - left + __ffs(m) - 1;
            if (left < 32) break;
            left -= 32; yw += dyw; cmin += dmin; r += dr;
            if (r >= den) { r -= den; cmin += smin; }
        }
    }
    return -(major + 1);
}

// 32 consecutive segments starting at `first`: lane-parallel set-up, warp-cooperative walks, coalesced results.
// s_pre: 32 SegPre of this warp.
template <class Grid, class GridOf>
__device__ __forceinline__ void walk_group(GridOf grid_of, int TY, const int4 *__restrict__ segs, int64_t nseg, int64_t first,
                                           uint8_t *__restrict__ free_out, int *__restrict__ cells_out, SegPre *s_pre, int lane)
{
    const int64_t mine = first + lane;
    int4 e = make_int4(0, 0, 0, 0);
    if (mine < nseg) e = __ldg(segs + mine);
    s_pre[lane] = seg_prepare(e);
    __syncwarp();
    const int cnt = (int)min((int64_t)32, nseg - first);
    int result = 0;
    for (int i = 0; i < cnt; ++i) {
        const SegPre p = s_pre[i];                                          // warp-uniform address: broadcast
        const Grid g = grid_of(first + i);
        const int r = (p.flags & 1) ? walk_axis<true>(g, TY, p, lane) : walk_axis<false>(g, TY, p, lane);
        if (lane == i) result = r;
    }
    __syncwarp();
    if (mine < nseg) {
        free_out[mine] = result < 0;
        if (cells_out) cells_out[mine] = cells_tested(result);
    }
}

constexpr int kCcThreads = 256;          // 8 warps per block, one group of 32 segments per warp at a time

// single world, grid read through L1/L2: one warp per group, blocks scheduled by the hardware
__global__ void __launch_bounds__(kCcThreads) collision_global_kernel(const uint32_t *__restrict__ bits, int TY, const int4 *__restrict__ segs,
                                                                      int64_t nseg, uint8_t *__restrict__ free_out, int *__restrict__ cells_out)
{
    __shared__ SegPre s_pre[kCcThreads];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kCcThreads) >> 5;
    const GlobalGrid g{bits};
    for (int64_t grp = ((int64_t)blockIdx.x * kCcThreads >> 5) + warp; grp * 32 < nseg; grp += nwarps)
        walk_group<GlobalGrid>([&](int64_t) { return g; }, TY, segs, nseg, grp * 32, free_out, cells_out, s_pre + warp * 32, lane);
}

// single world, grid staged in shared memory; persistent blocks
__global__ void collision_shared_kernel(const uint32_t *__restrict__ bits, int words, int TY, const int4 *__restrict__ segs,
                                        int64_t nseg, uint8_t *__restrict__ free_out, int *__restrict__ cells_out)
{
    extern __shared__ __align__(16) uint32_t s_dyn[];
    SegPre *s_pre = reinterpret_cast<SegPre *>(s_dyn);                      // blockDim.x entries
    uint32_t *s_grid = s_dyn + blockDim.x * (sizeof(SegPre) / 4);
    // The grid comes in with one bulk asynchronous copy (TMA, cp.async.bulk: no registers, no per-thread loop): thread 0 arms
    // an mbarrier with the byte count and issues the copy, everybody waits on the barrier's phase.  Both addresses and the size
    // are multiples of 128 bytes (whole tiles of a cudaMalloc'd grid).
    __shared__ __align__(8) unsigned long long s_bar;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)words * 4u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(s_grid)), "l"(bits), "r"(bytes), "r"(bar) : "memory");
    }
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "GRID_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
                 "@p bra GRID_DONE;\n"
                 "bra GRID_WAIT;\n"
                 "GRID_DONE:\n"
                 "}" ::"r"(bar) : "memory");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const SharedGrid g{s_grid};
    for (int64_t grp = ((int64_t)blockIdx.x * blockDim.x >> 5) + warp; grp * 32 < nseg; grp += nwarps)
        walk_group<SharedGrid>([&](int64_t) { return g; }, TY, segs, nseg, grp * 32, free_out, cells_out, s_pre + warp * 32, lane);
}

// per-segment world index (many small worlds): always through L1/L2
__global__ void __launch_bounds__(kCcThreads) collision_multi_kernel(const uint32_t *__restrict__ bits, size_t words_per, int TY,
                                                                     const int4 *__restrict__ segs, const int *__restrict__ world, int64_t nseg,
                                                                     uint8_t *__restrict__ free_out, int *__restrict__ cells_out)
{
    __shared__ SegPre s_pre[kCcThreads];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kCcThreads) >> 5;
    for (int64_t grp = ((int64_t)blockIdx.x * kCcThreads >> 5) + warp; grp * 32 < nseg; grp += nwarps)
        walk_group<GlobalGrid>([&](int64_t s) { return GlobalGrid{bits + (size_t)__ldg(world + s) * words_per}; }, TY, segs, nseg, grp * 32,
                               free_out, cells_out, s_pre + warp * 32, lane);
}

int collision_launch(const uint32_t *d_bits, int W, int H, const int32_t *d_segs, const int32_t *d_world, int64_t nseg,
                     uint8_t *d_free, int32_t *d_cells, int sm_count, int optin, cudaStream_t st)
{
    if (nseg == 0) return RRTK_OK;
    const int TY = tiles_y(H);
    const size_t words = grid_words(W, H);
    const int64_t groups = (nseg + 31) / 32;
    int64_t blocks = (groups * 32 + kCcThreads - 1) / kCcThreads;           // one warp per group of 32 segments
    if (blocks > 0x7fffffff) blocks = 0x7fffffff;
    const int4 *segs = reinterpret_cast<const int4 *>(d_segs);
    if (d_world) {
        collision_multi_kernel<<<(unsigned)blocks, kCcThreads, 0, st>>>(d_bits, words, TY, segs, d_world, nseg, d_free, d_cells);
    } else if (words * 4 + 1024 <= (size_t)optin / 2 && nseg >= 4096) {
        // grid on chip; as many blocks per SM as the copies allow (up to 4), persistent
        const size_t grid_b = words * 4;
        int per_sm = (int)((size_t)(optin + 1024) / (grid_b + 8192 + 1024));
        if (per_sm > 4) per_sm = 4;
        if (per_sm < 1) per_sm = 1;
        const int th = per_sm >= 4 ? 256 : (per_sm >= 2 ? 512 : 1024);
        const size_t smem = grid_b + (size_t)th * sizeof(SegPre);
        const int64_t cap = (int64_t)sm_count * per_sm;
        blocks = (groups * 32 + th - 1) / th;
        if (blocks > cap) blocks = cap;
        RRTK_CUDA(cudaFuncSetAttribute(collision_shared_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        collision_shared_kernel<<<(unsigned)blocks, th, smem, st>>>(d_bits, (int)words, TY, segs, nseg, d_free, d_cells);
    } else {
        collision_global_kernel<<<(unsigned)blocks, kCcThreads, 0, st>>>(d_bits, TY, segs, nseg, d_free, d_cells);
    }
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

}  // namespace rrtk
