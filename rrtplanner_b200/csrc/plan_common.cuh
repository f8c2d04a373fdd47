// Declarations shared by the two plan kernels (plan.cu: packed-key scan, the default;
// plan_wide.cu: 32-bit distances, any grid up to 16384 x 16384).
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace rrtk {

struct PlanParams {
    const uint32_t *bits;
    size_t words_per_grid;
    int W, H, TY;
    const rrtk_plan_desc *plans;
    int n;
    uint32_t r2_excl;     // radius test  d2 < r2_excl   (= ceil(r_rewire^2), capped at 2^30)
    double r_goal;
    const short2 *samples;
    const double2 *balls;
    short2 *pts;
    double *cost;
    int *parent;
    long long *stats;
    double *ell_c;
    // packed-key kernel only (plan.cu)
    int sbits;            // key = (d2 - |q|^2) << sbits | row, row = vertex index / T
    int list_cap;         // uint16 entries per sample in the radius-set lists
    int hit_words;        // membership words per thread per sample (32 rows each)
    int tail_bytes;       // storage of the last of them: 1, 2 or 4 bytes (a full tree fills only its first 8 / 16 / 32 rows)
    int steps_max;        // quads per thread that hold tree vertices: ceil(ceil((n + 1) / T) / 4)
    // bucket kernel only (plan_grid.cuh)
    int g_xb, g_yb;       // bits of x and y in a tree entry  id << (xb + yb) | y << xb | x
    int g_bshift, g_bshy; // buckets of 2^bshift x 2^bshy cells (x-major order: the buckets of one x form a run of slots)
    int g_nbx, g_nby;     // buckets per axis
    int g_rad;            // a sample reads the buckets that hold every cell within rad of it on both axes, rad >= r_rewire
    uint32_t g_near_ok2;  // rad^2: a nearest vertex at most this far away is the nearest of the whole tree
    int g_ent_words, g_off_list, g_off_bstart;   // shared-memory layout: entries (words), byte offsets of the lists and of the bucket starts
    int g_kb;             // id bits of a 32-bit (distance, id) key, 0 = distances and ids do not fit one word together
};

// what the owner warp of a sample hands to the commit phase
struct SampleRec {
    uint32_t pnew;        // packed sample
    uint32_t bd;          // squared distance to the nearest vertex of the round-start tree
    int vnear;
    int flags;            // 1 valid (iteration < n), 2 duplicate of a vertex >= 1, 4 nearest -> sample is free
    int bv;               // best parent found among the radius set (0x7fffffff: none beats the nearest vertex)
    int ring;             // |within(points, xnew, r_rewire)| on the round-start tree
    double c0;            // cost through the nearest vertex
    double bc;            // cost through bv
    double ell;           // informed: cbest the ellipse sample was drawn with
};

struct RoundSummary {
    int j, consumed, flags;       // flags: 1 have_sol, 2 finished
    int vsol;
    int kwant, pad;               // packed-key kernel: samples the next round should take (adaptive, see plan_scan.cuh)
    double csol;
    long long first_sol;
};

__device__ __forceinline__ unsigned warp_min_u32(unsigned v) { return __reduce_min_sync(RRTK_FULL, v); }

// informed ellipse sample, rrt.py:589-599 + 615-625 (rotation computed on the host, rrt.py:601-613)
__device__ __forceinline__ void ellipse_sample(int W, int H, const double *rot, int sx, int sy, int gx, int gy,
                                               double c, double2 ball, int &ox, int &oy)
{
    const double cx = __ddiv_rn((double)(sx + gx), 2.0), cy = __ddiv_rn((double)(sy + gy), 2.0);
    const double r1 = __ddiv_rn(c, 2.0);
    const long long ddx = sx - gx, ddy = sy - gy;
    const double d2 = (double)(ddx * ddx + ddy * ddy);
    const double r2 = __ddiv_rn(__dsqrt_rn(fabs(__dsub_rn(__dmul_rn(c, c), d2))), 2.0);
    const double m00 = __dmul_rn(rot[0], r1), m01 = __dmul_rn(rot[1], r2);
    const double m10 = __dmul_rn(rot[2], r1), m11 = __dmul_rn(rot[3], r2);
    const double x = __dadd_rn(__dadd_rn(__dmul_rn(m00, ball.x), __dmul_rn(m01, ball.y)), cx);
    const double y = __dadd_rn(__dadd_rn(__dmul_rn(m10, ball.x), __dmul_rn(m11, ball.y)), cy);
    double lx = (x < (double)(W - 1)) ? x : (double)(W - 1);     // NaN falls to W-1 like Python's min()
    double ly = (y < (double)(H - 1)) ? y : (double)(H - 1);
    lx = (lx > 0.0) ? lx : 0.0;
    ly = (ly > 0.0) ? ly : 0.0;
    ox = (int)lx;
    oy = (int)ly;
}


}  // namespace rrtk
