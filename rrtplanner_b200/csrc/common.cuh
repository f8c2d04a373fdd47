// Shared device helpers: tiled bit-grid addressing and the warp-cooperative line walk.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtk.h"

#define RRTK_FULL 0xffffffffu

namespace rrtk {

// ---- error plumbing (api.cu owns the buffer) ------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
#define RRTK_CUDA(call)                                              \
    do {                                                             \
        cudaError_t e_ = (call);                                     \
        if (e_ != cudaSuccess) return rrtk::cuda_fail(e_, #call);    \
    } while (0)

// ---- tiled bit grid (layout documented in include/rrtk.h) ------------------------------------
__host__ __device__ __forceinline__ int tiles_y(int H) { return (H + 31) >> 5; }
__host__ __device__ __forceinline__ int tiles_x(int W) { return (W + 31) >> 5; }
__host__ __device__ __forceinline__ size_t grid_words(int W, int H)
{
    return (size_t)tiles_x(W) * tiles_y(H) * 32;
}
__device__ __forceinline__ uint32_t word_index(int x, int y, int TY)
{
    return (uint32_t)((((x >> 5) * TY + (y >> 5)) << 5) | (x & 31));
}

// Grid accessors: the walk is templated on where the words live.
struct SharedGrid {
    const uint32_t *w;
    __device__ __forceinline__ uint32_t load(uint32_t i) const { return w[i]; }
};
struct GlobalGrid {
    const uint32_t *w;
    __device__ __forceinline__ uint32_t load(uint32_t i) const { return __ldg(w + i); }
};

// floor(num / den) for 0 <= num < 2^23, 0 < den <= 2^15 and a quotient <= 64, via one approximate fp32 reciprocal.
// (num + 1/2) / den lies at least 1 / (2 den) >= 1.5e-5 away from every integer, and the computed value is within
// 64 * 3 * 2^-24 = 1.2e-5 of it (one ulp for the reciprocal estimate, half an ulp each for the two roundings), so
// truncation gives the exact quotient without a fix-up.  -DRRTK_DIV_FIXUP: the earlier estimate with a +-1 correction.
__device__ __forceinline__ int small_div(int num, int den, int &rem)
{
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(__int2float_rn(den)));      // one MUFU
#ifdef RRTK_DIV_FIXUP
    int q = __float2int_rz(__int2float_rn(num) * inv);
    int r = num - q * den;
    if (r < 0) { r += den; --q; }
    if (r >= den) { r -= den; ++q; }
#else
    const int q = __float2int_rz((__int2float_rn(num) + 0.5f) * inv);
    const int r = num - q * den;
#endif
    rem = r;
    return q;
}

// Warp-cooperative form of RRT.collisionfree (rrt.py:183-229).  The reference walks
// L + 1 = max(|dx|,|dy|) + 1 cells with an integer error accumulator; cell k of that walk has the
// closed form
//     major axis:  k steps          minor axis:  floor((2*k*minor + major) / (2*major)) steps
// (checked exhaustively in tests/test_oracle.py::test_closed_form_cell_sequence), so lane l of the
// warp tests cell 32*c + l of chunk c, and a ballot finds the first occupied one.  All 32 lanes
// must call this with the same segment.  Returns k >= 0 = index of the first occupied cell, or
// -(L + 1) when the walk is free -- i.e. |ret| or ret + 1 is the number of cells the reference reads.
// The walk comes in two halves so that a caller can put independent work between the issue of the
// first chunk's grid load and the first ballot (walk_begin ... walk_finish); warp_first_hit is the
// two back to back.
struct WalkState {
    int ax, ay, sx, sy;
    int major, den;
    int k, q, r, dq, dr;
    bool xmajor;
    uint32_t word;        // grid word of this lane's cell in the first chunk (0 beyond the segment's end)
    int bitpos;
};

template <class Grid>
__device__ __forceinline__ WalkState walk_begin(const Grid &g, int TY, int ax, int ay, int bx, int by, int lane)
{
    WalkState s;
    const int dx = bx - ax, dy = by - ay;
    const int adx = abs(dx), ady = abs(dy);
    s.ax = ax; s.ay = ay;
    s.sx = dx > 0 ? 1 : -1; s.sy = dy > 0 ? 1 : -1;
    s.xmajor = adx >= ady;
    s.major = s.xmajor ? adx : ady;
    const int minor = s.xmajor ? ady : adx;
    s.den = 2 * s.major;
    s.k = lane;
    s.q = 0; s.r = 0;
    if (s.major > 0) s.q = small_div(2 * lane * minor + s.major, s.den, s.r);   // num < 63 * 16384 + ... < 2^24
    s.dq = 0; s.dr = 0;
    if (s.major >= 32) s.dq = small_div(64 * minor, s.den, s.dr);             // per-chunk increment of (q, r)
    const int cx = s.xmajor ? ax + s.sx * s.k : ax + s.sx * s.q;
    const int cy = s.xmajor ? ay + s.sy * s.q : ay + s.sy * s.k;
    s.word = 0;
    s.bitpos = cy & 31;
    if (s.k <= s.major) s.word = g.load(word_index(cx, cy, TY));
    return s;
}

template <class Grid>
__device__ __forceinline__ int walk_finish(const Grid &g, int TY, WalkState &s, int lane)
{
    unsigned m = __ballot_sync(RRTK_FULL, (s.word >> s.bitpos) & 1u);
    if (m) return __ffs(m) - 1;
    for (int base = 32; base <= s.major; base += 32) {
        s.k += 32;
        s.q += s.dq;
        s.r += s.dr;
        if (s.r >= s.den) { s.r -= s.den; ++s.q; }
        const int cx = s.xmajor ? s.ax + s.sx * s.k : s.ax + s.sx * s.q;
        const int cy = s.xmajor ? s.ay + s.sy * s.q : s.ay + s.sy * s.k;
        bool hit = false;
        if (s.k <= s.major) hit = (g.load(word_index(cx, cy, TY)) >> (cy & 31)) & 1u;
        m = __ballot_sync(RRTK_FULL, hit);
        if (m) return base + __ffs(m) - 1;
    }
    return -(s.major + 1);
}

template <class Grid>
__device__ __forceinline__ int warp_first_hit(const Grid &g, int TY, int ax, int ay, int bx, int by, int lane)
{
    WalkState s = walk_begin(g, TY, ax, ay, bx, by, lane);
    return walk_finish(g, TY, s, lane);
}

__device__ __forceinline__ int cells_tested(int first_hit_ret) { return first_hit_ret < 0 ? -first_hit_ret : first_hit_ret + 1; }

// packed tree vertex: x | y << 16 (coordinates < 16384)
__device__ __forceinline__ uint32_t pack_xy(int x, int y) { return (uint32_t)x | ((uint32_t)y << 16); }
__device__ __forceinline__ int px(uint32_t p) { return (int)(p & 0xffffu); }
__device__ __forceinline__ int py(uint32_t p) { return (int)(p >> 16); }
// vertex slot that is never nearest and never inside a radius (x = 49151): for any real point
// (coordinates < 16384) its squared distance is >= 2^30 > 2 * 16383^2 and < 2^32.
#define RRTK_FAR_VERTEX 0x0000BFFFu

__device__ __forceinline__ uint32_t dist2(uint32_t p, int qx, int qy)
{
    const int dx = px(p) - qx, dy = py(p) - qy;
    return (uint32_t)(dx * dx) + (uint32_t)(dy * dy);
}

// cost-to-come of v plus straight-line length: rrt.py:70-78 with r2norm of rrt.py:10-24 --
// exact integer d^2, correctly rounded f64 sqrt, one IEEE add (no contraction possible).
__device__ __forceinline__ double reach_cost(double cv, uint32_t d2) { return __dadd_rn(cv, __dsqrt_rn((double)d2)); }

}  // namespace rrtk
