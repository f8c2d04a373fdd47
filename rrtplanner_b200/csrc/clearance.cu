// K1b: RRT.collisionfree (rrt.py:183-229) on a clearance field -- same verdicts and the same
// "cells the reference reads" as K1 (collision.cu), far fewer grid reads for long segments.
//
// clear[x*H + y] = min(cap, Chebyshev distance from cell (x, y) to the nearest obstacle cell or to the
// outside of the grid); 0 on obstacles.  Consecutive cells of the reference's walk differ by one step
// on the major axis and at most one on the minor axis, so cell k + i lies within Chebyshev distance i
// of cell k: if clear(cell k) = d > 0 the cells k+1 .. k+d-1 are free and the walk may continue at
// k + d.  It stops at the first visited cell with clear = 0, which is the first occupied cell of the
// walk because every skipped cell was proven free -- the returned index (and with it the number of
// cells the reference's loop would have read) is identical.
//
// The field is built from the tiled bit grid by cap-1 passes of 8-connected dilation (one thread per
// 32-cell word, 9 word loads), each pass writing its distance into the cells it newly reaches.
// The walk is one thread per segment: every step needs cell k in closed form,
//     q(k) = floor((2 k minor + major) / (2 major)),
// taken from an fp32 reciprocal estimate with an exact +-1 integer fix-up (num < 2^31 for grids up to
// 32768 cells a side), then one byte load.  A warp finishes with its slowest lane, so lanes whose
// segment is done take the next one from the warp's block of segments while the others keep stepping.
#include <cstdlib>

#include "common.cuh"

namespace rrtk {

// ---- clearance field ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t word_or_ones(const uint32_t *g, int x, int ty, int W32, int TY)
{
    if (x < 0 || x >= W32 || ty < 0 || ty >= TY) return 0xffffffffu;          // outside the grid counts as obstacle
    return g[(((x >> 5) * TY + ty) << 5) | (x & 31)];
}

__device__ __forceinline__ uint32_t ydilate(const uint32_t *g, int x, int ty, int W32, int TY)
{
    const uint32_t c = word_or_ones(g, x, ty, W32, TY);
    return c | (c << 1) | (c >> 1) | (word_or_ones(g, x, ty - 1, W32, TY) >> 31) | (word_or_ones(g, x, ty + 1, W32, TY) << 31);
}

__global__ void clearance_init_kernel(const uint32_t *__restrict__ bits, int nworlds, int W, int H, int cap, uint8_t *__restrict__ clear)
{
    const int TY = tiles_y(H);
    const size_t cells = (size_t)W * H, total = cells * nworlds, words_per = grid_words(W, H);
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int world = (int)(t / cells);
        const size_t c = t - (size_t)world * cells;
        const int x = (int)(c / H), y = (int)(c - (size_t)x * H);
        const bool occ = (bits[(size_t)world * words_per + word_index(x, y, TY)] >> (y & 31)) & 1u;
        clear[t] = occ ? 0 : (uint8_t)cap;
    }
}

// one 8-connected dilation pass in -> out; cells reached by this pass get distance `it`
__global__ void clearance_pass_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int nworlds, int W, int H, int it,
                                      uint8_t *__restrict__ clear)
{
    const int TX = tiles_x(W), TY = tiles_y(H), W32 = TX * 32;
    const size_t words_per = (size_t)TX * TY * 32;
    const size_t total = words_per * nworlds;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int world = (int)(t / words_per);
        const int local = (int)(t - (size_t)world * words_per);
        const int xl = local & 31, tile = local >> 5;
        const int ty = tile % TY, x = (tile / TY) * 32 + xl;
        const uint32_t *g = in + (size_t)world * words_per;
        const uint32_t prev = g[local];
        const uint32_t next = ydilate(g, x, ty, W32, TY) | ydilate(g, x - 1, ty, W32, TY) | ydilate(g, x + 1, ty, W32, TY);
        out[t] = next;
        uint32_t fresh = next & ~prev;
        if (x < W) {
            uint8_t *row = clear + ((size_t)world * W + x) * H + ty * 32;
            const int valid = min(32, H - ty * 32);
            while (fresh) {
                const int b = __ffs(fresh) - 1;
                fresh &= fresh - 1;
                if (b < valid) row[b] = (uint8_t)it;
            }
        }
    }
}

int clearance_launch(const uint32_t *d_bits, int nworlds, int W, int H, int cap, uint8_t *d_clear, uint32_t *d_scratch, cudaStream_t st)
{
    const size_t words = grid_words(W, H) * nworlds, cells = (size_t)W * H * nworlds;
    if (cells == 0) return RRTK_OK;
    const int threads = 256;
    size_t cb = (cells + threads - 1) / threads, wb = (words + threads - 1) / threads;
    if (cb > 148 * 64) cb = 148 * 64;
    if (wb > 148 * 64) wb = 148 * 64;
    clearance_init_kernel<<<(unsigned)cb, threads, 0, st>>>(d_bits, nworlds, W, H, cap, d_clear);
    const uint32_t *src = d_bits;
    uint32_t *a = d_scratch, *b = d_scratch + words;
    for (int it = 1; it < cap; ++it) {
        clearance_pass_kernel<<<(unsigned)wb, threads, 0, st>>>(src, a, nworlds, W, H, it, d_clear);
        src = a;
        uint32_t *tmp = a; a = b; b = tmp;
    }
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

// ---- directional clearance: one field per octant of the walk -------------------------------------------
// The walk of a segment never turns: with (major axis, sign of dx, sign of dy) fixed, cell k + i lies i steps ahead on the
// major axis and between 0 and i steps ahead on the minor axis.  So instead of the Chebyshev ball the field of an octant
// keeps, per cell, the depth of the free *cone* ahead of it:
//     D(c) = 0 on obstacles, else min(cap, 1 + min(D(c + major step), D(c + major step + minor step)))
// (cells outside the grid count as free: the walk stays inside the bounding box of its end points).  D(cell k) = d > 0
// proves cells k+1 .. k+d-1 free exactly like the isotropic field, but an obstacle beside or behind the walk no longer
// shortens the step: on cfg2 a segment takes 4.9 reads instead of 10.1 (p99: 17 instead of 44), at 8 bytes per cell.
// Octant o = 4 * (x is the major axis) + 2 * (dx > 0) + (dy > 0); layout clear8[(world * 8 + o) * W * H + x * H + y].
//
// Built by one sweep against the direction of the major axis: a thread keeps one minor coordinate and the block exchanges
// the previous column through shared memory.  D saturates at cap, so a value depends on at most cap - 1 further minor
// rows: stripes of the minor axis overlap by that much and need no exchange between blocks.
constexpr int kDirThreadsMax = 1024;

__global__ void __launch_bounds__(kDirThreadsMax) clearance_dir_kernel(const uint32_t *__restrict__ bits, int W, int H, int cap, int stripes, int rows_per,
                                                                      int wide, uint8_t *__restrict__ clear8)
{
    __shared__ uint8_t s_col[2][kDirThreadsMax + 4];
    const int tid = threadIdx.x;
    const int stripe = blockIdx.x % stripes, oct = (blockIdx.x / stripes) & 7, world = blockIdx.x / (stripes * 8);
    const bool xmajor = (oct & 4) != 0;
    const bool maj_pos = xmajor ? (oct & 2) != 0 : (oct & 1) != 0, min_pos = xmajor ? (oct & 1) != 0 : (oct & 2) != 0;
    const int nmaj = xmajor ? W : H, nmin = xmajor ? H : W;
    const int TY = tiles_y(H);
    if (stripe * rows_per >= nmin) return;                                     // block-uniform
    const int u = stripe * rows_per + tid;                                     // minor coordinate, counted along the walk's minor direction
    const bool inside = u < nmin;
    const bool writes = inside && tid < rows_per;
    const int m = min_pos ? u : nmin - 1 - u;
    const uint32_t *g = bits + (size_t)world * grid_words(W, H);
    uint8_t *field = clear8 + ((size_t)world * 8 + oct) * ((size_t)W * H);
    s_col[0][tid] = (uint8_t)cap;
    if (tid == 0) { s_col[0][blockDim.x] = (uint8_t)cap; s_col[1][blockDim.x] = (uint8_t)cap; }
    __syncthreads();
    int own = cap, cur = 0;
    uint32_t acc = 0;
    for (int v = nmaj - 1; v >= 0; --v) {
        const int M = maj_pos ? v : nmaj - 1 - v;
        int d = cap;
        if (inside) {
            const int x = xmajor ? M : m, y = xmajor ? m : M;
            const bool occ = (g[word_index(x, y, TY)] >> (y & 31)) & 1u;
            d = occ ? 0 : min(cap, 1 + min(own, (int)s_col[cur][tid + 1]));
            if (writes) {
                if (xmajor) field[(size_t)M * H + m] = (uint8_t)d;                  // threads <-> consecutive y: whole sectors
                else if (!wide) field[(size_t)m * H + M] = (uint8_t)d;
                else {                                                              // threads <-> x (stride H): four steps per store
                    acc |= (uint32_t)d << (8 * (M & 3));
                    if ((M & 3) == (maj_pos ? 0 : 3)) {
                        *reinterpret_cast<uint32_t *>(field + (size_t)m * H + (M & ~3)) = acc;
                        acc = 0;
                    }
                }
            }
        }
        own = d;
        s_col[cur ^ 1][tid] = (uint8_t)d;
        cur ^= 1;
        __syncthreads();
    }
}

int clearance_dir_launch(const uint32_t *d_bits, int nworlds, int W, int H, int cap, uint8_t *d_clear8, cudaStream_t st)
{
    if ((size_t)W * H * nworlds == 0) return RRTK_OK;
    const int nmin_max = W > H ? W : H;
    const int threads = nmin_max >= kDirThreadsMax ? kDirThreadsMax : (nmin_max + 31) & ~31;
    const int rows_per = nmin_max <= threads ? threads : threads - (cap - 1);     // one stripe sees the whole axis: no overlap needed
    const int stripes = (nmin_max + rows_per - 1) / rows_per;
    const int wide = (H % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_clear8) & 3) == 0);
    clearance_dir_kernel<<<(unsigned)((size_t)nworlds * 8 * stripes), threads, 0, st>>>(d_bits, W, H, cap, stripes, rows_per, wide, d_clear8);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

// ---- sixteen fields: every octant split at slope 1/2 ----------------------------------------------------------------------
// Cell k + i of a walk with minor / major <= 1/2 lies between 0 and ceil(i / 2) minor steps ahead, with minor / major > 1/2
// between floor(i / 2) and i: two narrower cones per octant, i.e. longer steps again (cfg2: 3.9 reads per segment instead of
// 4.9, p99 12 instead of 17) at 16 bytes per cell.  The cones are not self-similar under one major step but under two, so each
// half keeps a helper field for the odd phase (a = major step, b = minor step, all capped, 0 on obstacles):
//     low  half  A(c) = 1 + min(B(c + a), B(c + a + b)),   B(c) = 1 + A(c + a)
//     high half  C(c) = 1 + min(D(c + a), D(c + a + b)),   D(c) = 1 + C(c + a + b)
// A and C are stored (fields 2 o and 2 o + 1 of octant o); B and D live only in the sweep.
__global__ void __launch_bounds__(kDirThreadsMax) clearance_dir16_kernel(const uint32_t *__restrict__ bits, int W, int H, int cap, int stripes,
                                                                        int rows_per, int wide, uint8_t *__restrict__ clear16)
{
    __shared__ uint8_t s_b[2][kDirThreadsMax + 4], s_c[2][kDirThreadsMax + 4], s_d[2][kDirThreadsMax + 4];
    const int tid = threadIdx.x;
    const int stripe = blockIdx.x % stripes, oct = (blockIdx.x / stripes) & 7, world = blockIdx.x / (stripes * 8);
    const bool xmajor = (oct & 4) != 0;
    const bool maj_pos = xmajor ? (oct & 2) != 0 : (oct & 1) != 0, min_pos = xmajor ? (oct & 1) != 0 : (oct & 2) != 0;
    const int nmaj = xmajor ? W : H, nmin = xmajor ? H : W;
    const int TY = tiles_y(H);
    if (stripe * rows_per >= nmin) return;                                     // block-uniform
    const int u = stripe * rows_per + tid;
    const bool inside = u < nmin;
    const bool writes = inside && tid < rows_per;
    const int m = min_pos ? u : nmin - 1 - u;
    const uint32_t *g = bits + (size_t)world * grid_words(W, H);
    uint8_t *fa = clear16 + ((size_t)world * 16 + 2 * oct) * ((size_t)W * H), *fc = fa + (size_t)W * H;
    for (int k = 0; k < 2; ++k) {
        s_b[k][tid] = s_c[k][tid] = s_d[k][tid] = (uint8_t)cap;
        if (tid == 0) s_b[k][blockDim.x] = s_c[k][blockDim.x] = s_d[k][blockDim.x] = (uint8_t)cap;
    }
    __syncthreads();
    int pa = cap, pb = cap, pd = cap, cur = 0;                                  // this thread's values one major step ahead
    uint32_t acca = 0, accc = 0;
    for (int v = nmaj - 1; v >= 0; --v) {
        const int M = maj_pos ? v : nmaj - 1 - v;
        int a = cap, b = cap, c = cap, d = cap;
        if (inside) {
            const int x = xmajor ? M : m, y = xmajor ? m : M;
            const bool occ = (g[word_index(x, y, TY)] >> (y & 31)) & 1u;
            a = occ ? 0 : min(cap, 1 + min(pb, (int)s_b[cur][tid + 1]));
            b = occ ? 0 : min(cap, 1 + pa);
            c = occ ? 0 : min(cap, 1 + min(pd, (int)s_d[cur][tid + 1]));
            d = occ ? 0 : min(cap, 1 + (int)s_c[cur][tid + 1]);
            if (writes) {
                if (xmajor) { fa[(size_t)M * H + m] = (uint8_t)a; fc[(size_t)M * H + m] = (uint8_t)c; }
                else if (!wide) { fa[(size_t)m * H + M] = (uint8_t)a; fc[(size_t)m * H + M] = (uint8_t)c; }
                else {
                    acca |= (uint32_t)a << (8 * (M & 3)); accc |= (uint32_t)c << (8 * (M & 3));
                    if ((M & 3) == (maj_pos ? 0 : 3)) {
                        *reinterpret_cast<uint32_t *>(fa + (size_t)m * H + (M & ~3)) = acca;
                        *reinterpret_cast<uint32_t *>(fc + (size_t)m * H + (M & ~3)) = accc;
                        acca = accc = 0;
                    }
                }
            }
        }
        pa = a; pb = b; pd = d;
        s_b[cur ^ 1][tid] = (uint8_t)b; s_c[cur ^ 1][tid] = (uint8_t)c; s_d[cur ^ 1][tid] = (uint8_t)d;
        cur ^= 1;
        __syncthreads();
    }
}

int clearance_dir16_launch(const uint32_t *d_bits, int nworlds, int W, int H, int cap, uint8_t *d_clear16, cudaStream_t st)
{
    if ((size_t)W * H * nworlds == 0) return RRTK_OK;
    const int nmin_max = W > H ? W : H;
    const int threads = nmin_max >= kDirThreadsMax ? kDirThreadsMax : (nmin_max + 31) & ~31;
    const int rows_per = nmin_max <= threads ? threads : threads - (cap - 1);
    const int stripes = (nmin_max + rows_per - 1) / rows_per;
    const int wide = (H % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_clear16) & 3) == 0);
    clearance_dir16_kernel<<<(unsigned)((size_t)nworlds * 8 * stripes), threads, 0, st>>>(d_bits, W, H, cap, stripes, rows_per, wide, d_clear16);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

// ---- the walk ---------------------------------------------------------------------------------------
// A segment as the walk wants it, made lane-parallel (all 32 lanes busy) when the warp stages its pool and kept in shared
// memory as two 16-byte words, so that a lane that draws a new segment only loads them:
//   a = (base, step_k, step_q, major)      cell k of the walk lives at clear[base + k * step_k + q(k) * step_q]
//   b = (minor, fp32 bits of ~1 / (2 major), field, result)        field: which W x H field to read; result: written when the walk is over
struct CfRec { int4 a, b; };

__device__ __forceinline__ CfRec cf_prepare(int4 e, int world, int H, int nf)
{
    const int dx = e.z - e.x, dy = e.w - e.y;
    const int adx = abs(dx), ady = abs(dy);
    const bool xmajor = adx >= ady;
    const int major = xmajor ? adx : ady, minor = xmajor ? ady : adx;
    float inv = 0.f;
    if (major > 0) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(__int2float_rn(2 * major)));
    const int sxH = dx > 0 ? H : -H, sy = dy > 0 ? 1 : -1;                            // sx = +1 iff x0 < x1 (rrt.py:207-215)
    CfRec r;
    r.a = make_int4(e.x * H + e.y, xmajor ? sxH : sy, xmajor ? sy : sxH, major);
    // which field of the world: nf = 1 (isotropic), 8 (one per octant) or 16 (octants split at minor / major = 1/2)
    const int oct = (xmajor ? 4 : 0) | (dx > 0 ? 2 : 0) | (dy > 0 ? 1 : 0);
    const int sector = nf == 16 ? 2 * oct + (2 * minor > major ? 1 : 0) : (oct & (nf - 1));
    r.b = make_int4(minor, __float_as_int(inv), world * nf + sector, 0);
    return r;
}

// Scheduling.  A segment needs anything between one and ~60 dependent reads (how far it gets before it hits something is
// close to exponential, whatever its length), and a warp finishes with its slowest lane: 17 of 32 threads are active on
// average (profiles/r2_v2_cf_ncu.txt).  Each warp owns a pool of kCfPerWarp segments staged in shared memory and its lanes
// draw the next one as they finish.  The kernel is bound by instruction issue (68 % of the issue slots, 48 warp
// instructions per segment before the records were staged ready-made), so what pays is fewer instructions per step and
// per draw.  Tried on cfg2 and measured without gain: drawing the pool longest first (length does not predict the work),
// time-slicing the segments between the lanes (0.130 ms instead of 0.089: the swaps cost more than the idle lanes), several
// segments per lane at once (2 / 3 / 4: 0.112 / 0.152 / 0.179 ms), larger pools with fewer warps (256: 0.114 ms).
constexpr int kCfThreads = 128;
#ifndef RRTK_CF_MINB
#define RRTK_CF_MINB 12
#endif

template <int kCfPerWarp>
__global__ void __launch_bounds__(kCfThreads, RRTK_CF_MINB) collision_cf_kernel(const uint8_t *__restrict__ clear, size_t cells_per, int W, int H,
                                                                  const int4 *__restrict__ segs, const int *__restrict__ world, int nf,
                                                                  int64_t nseg, uint8_t *__restrict__ free_out, int *__restrict__ cells_out)
{
    __shared__ int4 s_a[kCfThreads / 32][kCfPerWarp];
    __shared__ int4 s_b[kCfThreads / 32][kCfPerWarp];
    static_assert(kCfPerWarp % 32 == 0, "pool shape");
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * kCfThreads + threadIdx.x) >> 5;
    const int64_t first = warp * kCfPerWarp;
    if (first >= nseg) return;
    const int cnt = (int)min((int64_t)kCfPerWarp, nseg - first);
    int4 *ra = s_a[wib], *rb = s_b[wib];
    for (int i = lane; i < cnt; i += 32) {                                           // coalesced: 512 bytes per step
        const CfRec r = cf_prepare(__ldg(segs + first + i), world ? __ldg(world + first + i) : 0, H, nf);
        ra[i] = r.a; rb[i] = r.b;
    }
    __syncwarp();
    int next = 32;                                   // next undistributed slot (warp-uniform)
    int slot = lane, k = 0;
    bool active = lane < cnt;
    int4 a = make_int4(0, 0, 0, 0), b = a;
    const uint8_t *field = clear;
    if (active) {
        a = ra[slot]; b = rb[slot];
        field = clear + (size_t)b.z * cells_per;
    }
    while (__any_sync(RRTK_FULL, active)) {
        bool done = false;
        if (active) {
            // cell k of the walk: q(k) = floor((2 k minor + major) / (2 major)); a = (base, step_k, step_q, major), b.x = minor
            int q = 0;
            if (a.w > 0) {
                const unsigned den = 2u * (unsigned)a.w;
                const unsigned num = 2u * (unsigned)k * (unsigned)b.x + (unsigned)a.w;       // < 2^31
                q = __float2int_rz(__uint2float_rn(num) * __int_as_float(b.y));
                int r = (int)(num - (unsigned)q * den);
                if (r < 0) { --q; r += (int)den; }
                if (r >= (int)den) ++q;
            }
            const int d = __ldg(field + (a.x + k * a.y + q * a.z));
            int result = k;                                                          // d == 0: first occupied cell
            done = d == 0;
            k += d;                                                                  // cells k+1 .. k+d-1 are free
            if (k > a.w) { done = true; result = -(a.w + 1); }
            if (done) rb[slot].w = result;
        }
        // lanes that finished take the next slots of the pool, in lane order
        const unsigned fin = __ballot_sync(RRTK_FULL, active && done);
        if (fin) {
            if (active && done) {
                slot = next + __popc(fin & ((1u << lane) - 1u));
                active = slot < cnt;
                if (active) {
                    a = ra[slot]; b = rb[slot];
                    k = 0;
                    field = clear + (size_t)b.z * cells_per;
                }
            }
            next += __popc(fin);
        }
    }
    __syncwarp();
    for (int i = lane; i < cnt; i += 32) {                                            // coalesced results
        const int r = rb[i].w;
        free_out[first + i] = r < 0;
        if (cells_out) cells_out[first + i] = cells_tested(r);
    }
}

template <int kPool>
static void cf_launch_pool(const uint8_t *d_clear, int W, int H, const int32_t *d_segs, const int32_t *d_world, int nf, int64_t nseg,
                           uint8_t *d_free, int32_t *d_cells, cudaStream_t st)
{
    const int64_t warps = (nseg + kPool - 1) / kPool;
    const int64_t blocks = (warps * 32 + kCfThreads - 1) / kCfThreads;
    collision_cf_kernel<kPool><<<(unsigned)blocks, kCfThreads, 0, st>>>(d_clear, (size_t)W * H, W, H, reinterpret_cast<const int4 *>(d_segs),
                                                                      d_world, nf, nseg, d_free, d_cells);
}

int collision_cf_launch(const uint8_t *d_clear, int W, int H, const int32_t *d_segs, const int32_t *d_world, int nf, int64_t nseg,
                        uint8_t *d_free, int32_t *d_cells, int sm_count, cudaStream_t st)
{
    if (nseg == 0) return RRTK_OK;
    // segments per warp: enough for several draws per lane, few enough for the SMs to be full of warps on a big launch
    // (cfg2, 1 Mi segments, one segment per lane: pools of 64 / 128 / 256 -> 0.094 / 0.089 / 0.114 ms); RRTK_CF_POOL overrides
    int64_t per = nseg / ((int64_t)sm_count * 128) + 1;
    const char *env = getenv("RRTK_CF_POOL");
    if (env && *env) per = atoi(env);
    if (per > 128) cf_launch_pool<256>(d_clear, W, H, d_segs, d_world, nf, nseg, d_free, d_cells, st);
    else if (per > 64) cf_launch_pool<128>(d_clear, W, H, d_segs, d_world, nf, nseg, d_free, d_cells, st);
    else if (per > 32) cf_launch_pool<64>(d_clear, W, H, d_segs, d_world, nf, nseg, d_free, d_cells, st);
    else cf_launch_pool<32>(d_clear, W, H, d_segs, d_world, nf, nseg, d_free, d_cells, st);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

}  // namespace rrtk
