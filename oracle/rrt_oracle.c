/*
 * CPU oracle (plain C) for the rrtplanner tree-expansion hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * A from-scratch restatement of the algorithms of the reference's rrtplanner/rrt.py, fast enough
 * to check the CUDA path at BASELINE.json's full sizes (n = 5000 .. 20000).  It is never linked or
 * loaded by the product package; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg use it (through oracle/c_oracle.py).
 *
 * Parity status: PINNED -- tests/test_oracle.py checks every entry point against the golden
 * vectors under tests/golden/ that tests/golden/make_golden.py produced by running the real
 * reference, and against oracle/rrt_oracle.py.
 *
 * Pinned conventions (same as oracle/rrt_oracle.py):
 *   nearest vertex  : lowest index among equals           (rrt.py:150-155 is an unstable argsort)
 *   goal connection : ascending (cost, index), filled vertices only; none visible -> vgoal = 0,
 *                     no goal row                        (rrt.py:317-331; the reference reads out
 *                                                         of bounds when it reaches unfilled slots)
 *
 * Build:  gcc -O2 -fPIC -shared -ffp-contract=off -o oracle/_build/liboracle.so oracle/rrt_oracle.c -lm
 * (-ffp-contract=off: costs must be sqrt-then-add in IEEE double, no fused multiply-add.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KIND_STANDARD 0
#define KIND_STAR 1
#define KIND_INFORMED 2

/* stats slots written by orc_plan */
enum {
    ST_J = 0, ST_VGOAL, ST_FOUND, ST_CHECKS, ST_CELLS, ST_FIRST_SOL_ITER, ST_ELL_ITERS,
    ST_NN_PAIRS, ST_RING_MEMBERS, ST_ACCEPTED, ST_REWIRE_FIRED, ST_COUNT
};

/* ---- rrt.py:183-229 : integer line walk, start and end cells included -------------------- */
/* returns k >= 0 = index of first occupied cell, or -(cells tested) when the walk is free */
int orc_first_hit(const uint8_t *og, int W, int H, int ax, int ay, int bx, int by)
{
    (void)W;
    int adx = abs(bx - ax), ady = abs(by - ay);
    int stepx = ax < bx ? 1 : -1, stepy = ay < by ? 1 : -1;
    int acc = adx - ady, x = ax, y = ay, k = 0;
    for (;;) {
        if (og[(size_t)x * H + y]) return k;
        if (x == bx && y == by) return -(k + 1);
        int twice = 2 * acc;
        if (twice >= -ady) { acc -= ady; x += stepx; }
        if (twice <= adx)  { acc += adx; y += stepy; }
        ++k;
    }
}

void orc_collision_batch(const uint8_t *og, int W, int H, const int32_t *segs, long nseg,
                         uint8_t *free_out, int32_t *cells_out)
{
    for (long s = 0; s < nseg; ++s) {
        int r = orc_first_hit(og, W, H, segs[4 * s], segs[4 * s + 1], segs[4 * s + 2], segs[4 * s + 3]);
        free_out[s] = r < 0;
        if (cells_out) cells_out[s] = r < 0 ? -r : r + 1;
    }
}

/* ---- rrt.py:131-155 (element 0 only), pinned tie rule ------------------------------------- */
int orc_nearest(const int32_t *pts, int j, int x, int y, int64_t *d2_out)
{
    int best = 0;
    int64_t bd = INT64_MAX;
    for (int v = 0; v < j; ++v) {
        int64_t dx = (int64_t)pts[2 * v] - x, dy = (int64_t)pts[2 * v + 1] - y;
        int64_t d = dx * dx + dy * dy;
        if (d < bd) { bd = d; best = v; }
    }
    if (d2_out) *d2_out = bd;
    return best;
}

/* ---- rrt.py:157-181 : strict d^2 < r*r, ascending index ------------------------------------ */
int orc_within(const int32_t *pts, int j, int x, int y, double r, int32_t *out)
{
    int m = 0;
    double rr = r * r;
    for (int v = 0; v < j; ++v) {
        int64_t dx = (int64_t)pts[2 * v] - x, dy = (int64_t)pts[2 * v + 1] - y;
        if ((double)(dx * dx + dy * dy) < rr) out[m++] = v;
    }
    return m;
}

/* ---- rrt.py:70-78 + 10-24 ------------------------------------------------------------------ */
static inline double reach(const double *cost, const int32_t *pts, int v, int x, int y)
{
    int64_t dx = (int64_t)pts[2 * v] - x, dy = (int64_t)pts[2 * v + 1] - y;
    return cost[v] + sqrt((double)(dx * dx + dy * dy));
}

typedef struct { double c; int v; } cand_t;
static int cand_cmp(const void *a, const void *b)
{
    const cand_t *p = a, *q = b;
    if (p->c < q->c) return -1;
    if (p->c > q->c) return 1;
    return (p->v > q->v) - (p->v < q->v);
}

/* ---- rrt.py:589-599, 615-625 : informed ellipse sample ------------------------------------ */
static void ellipse_point(int W, int H, const double rot[4], const int32_t *s, const int32_t *g,
                          double c, const double ball[2], int *ox, int *oy)
{
    double cx = (s[0] + g[0]) / 2.0, cy = (s[1] + g[1]) / 2.0;
    double r1 = c / 2.0;
    int64_t gx = (int64_t)s[0] - g[0], gy = (int64_t)s[1] - g[1];
    double d2 = (double)(gx * gx + gy * gy);
    double r2 = sqrt(fabs(c * c - d2)) / 2.0;
    double m00 = rot[0] * r1, m01 = rot[1] * r2, m10 = rot[2] * r1, m11 = rot[3] * r2;
    double x = m00 * ball[0] + m01 * ball[1] + cx;
    double y = m10 * ball[0] + m11 * ball[1] + cy;
    /* int(max(0, min(dim-1, v))) with Python's comparison semantics (NaN falls to dim-1) */
    double lx = (x < W - 1) ? x : (double)(W - 1);
    double ly = (y < H - 1) ? y : (double)(H - 1);
    lx = (lx > 0) ? lx : 0.0;
    ly = (ly > 0) ? ly : 0.0;
    *ox = (int)lx;
    *oy = (int)ly;
}

/*
 * One plan on explicit streams.
 *   kind 0: rrt.py:386-447   kind 1: rrt.py:466-556   kind 2: rrt.py:653-758
 * samples[i] (int32 x,y) feeds iteration i while no solution vertex exists; balls[i] (unit-ball
 * point, rrt.py:579-587) feeds it afterwards (kind 2 only).  rot = row-major 2x2 of rrt.py:601-613.
 * Outputs have n + 1 rows; rows that the reference leaves unfilled hold pts = INT32_MIN,
 * cost = +inf, parent = -1.  Row layout after goal connection follows rrt.py:320-323 (the caller
 * rebuilds the duplicate goal row n when j < n).  ell_c[j] = cbest of the last ellipse sample
 * drawn while the tree had j vertices (rrt.py:701), NaN if none.
 */
int orc_plan(int kind, const uint8_t *og, int W, int H, int n, double r_rewire, double r_goal,
             const int32_t *start, const int32_t *goal, const int32_t *samples, const double *balls,
             const double *rot, int32_t *pts, double *cost, int32_t *parent, int64_t *stats,
             double *ell_c)
{
    uint8_t *seen = calloc((size_t)W * H, 1);
    int32_t *ring = malloc(sizeof(int32_t) * (size_t)(n + 1));
    cand_t *cands = malloc(sizeof(cand_t) * (size_t)(n + 1));
    if (!seen || !ring || !cands) { free(seen); free(ring); free(cands); return -1; }
    for (int i = 0; i <= n; ++i) {
        pts[2 * i] = pts[2 * i + 1] = INT32_MIN;
        cost[i] = INFINITY;
        parent[i] = -1;
        if (ell_c) ell_c[i] = NAN;
    }
    memset(stats, 0, sizeof(int64_t) * ST_COUNT);
    stats[ST_FIRST_SOL_ITER] = -1;
    pts[0] = start[0]; pts[1] = start[1]; cost[0] = 0.0;
    int j = 1;
    int have_sol = 0, vsol = -1;       /* running least_cost over vsoln, rrt.py:627-633 */
    double csol = INFINITY;

    for (int i = 0; i < n; ++i) {
        int x, y;
        if (kind == KIND_INFORMED && have_sol) {
            int64_t dx = (int64_t)goal[0] - pts[2 * vsol], dy = (int64_t)goal[1] - pts[2 * vsol + 1];
            double c = csol + sqrt((double)(dx * dx + dy * dy));           /* rrt.py:698-699 */
            ellipse_point(W, H, rot, start, goal, c, balls + 2 * i, &x, &y);
            if (ell_c) ell_c[j] = c;
            stats[ST_ELL_ITERS]++;
        } else {
            x = samples[2 * i]; y = samples[2 * i + 1];
        }
        int vnear = orc_nearest(pts, j, x, y, NULL);
        stats[ST_NN_PAIRS] += j;
        int r = orc_first_hit(og, W, H, pts[2 * vnear], pts[2 * vnear + 1], x, y);
        stats[ST_CHECKS]++; stats[ST_CELLS] += r < 0 ? -r : r + 1;
        if (r >= 0 || seen[(size_t)x * H + y] || j == n) continue;        /* rrt.py:425,507,707 */
        seen[(size_t)x * H + y] = 1;
        int vbest = vnear;
        double cbest = reach(cost, pts, vnear, x, y);
        if (kind != KIND_STANDARD) {
            int m = orc_within(pts, j, x, y, r_rewire, ring);
            stats[ST_RING_MEMBERS] += m;
            for (int k = 0; k < m; ++k) {                                  /* rrt.py:515-521 */
                int vn = ring[k];
                double cn = reach(cost, pts, vn, x, y);
                if (cn < cbest) {
                    int rr = orc_first_hit(og, W, H, pts[2 * vn], pts[2 * vn + 1], x, y);
                    stats[ST_CHECKS]++; stats[ST_CELLS] += rr < 0 ? -rr : rr + 1;
                    if (rr < 0) { vbest = vn; cbest = cn; }
                }
            }
            for (int k = 0; k < m; ++k)                                    /* rrt.py:532-536 */
                if (reach(cost, pts, ring[k], x, y) < cost[ring[k]]) stats[ST_REWIRE_FIRED]++;
        }
        pts[2 * j] = x; pts[2 * j + 1] = y; cost[j] = cbest; parent[j] = vbest;
        if (kind == KIND_INFORMED) {
            int64_t dx = (int64_t)x - goal[0], dy = (int64_t)y - goal[1];
            if (sqrt((double)(dx * dx + dy * dy)) < r_goal) {              /* rrt.py:744-745 */
                if (!have_sol) stats[ST_FIRST_SOL_ITER] = i;
                if (!have_sol || cbest < csol) { csol = cbest; vsol = j; }
                have_sol = 1;
            }
        }
        stats[ST_ACCEPTED]++;
        ++j;
    }

    /* goal connection, rrt.py:284-332 */
    for (int v = 0; v < j; ++v) { cands[v].c = reach(cost, pts, v, goal[0], goal[1]); cands[v].v = v; }
    qsort(cands, (size_t)j, sizeof(cand_t), cand_cmp);
    int vgoal = 0, found = 0;
    for (int k = 0; k < j; ++k) {
        int v = cands[k].v;
        int rr = orc_first_hit(og, W, H, pts[2 * v], pts[2 * v + 1], goal[0], goal[1]);
        stats[ST_CHECKS]++; stats[ST_CELLS] += rr < 0 ? -rr : rr + 1;
        if (rr < 0) {
            vgoal = j; found = 1;
            pts[2 * j] = goal[0]; pts[2 * j + 1] = goal[1];
            cost[j] = cands[k].c; parent[j] = v;
            break;
        }
    }
    stats[ST_J] = j; stats[ST_VGOAL] = vgoal; stats[ST_FOUND] = found;
    free(seen); free(ring); free(cands);
    return 0;
}

/* many independent plans on one thread each is the caller's business (bench.py uses processes);
 * this helper runs a list sequentially so that one ctypes call amortises the FFI cost. */
int orc_plan_many(int kind, int nplans, const uint8_t *ogs, const int64_t *og_offset, int W, int H,
                  int n, double r_rewire, double r_goal, const int32_t *starts, const int32_t *goals,
                  const int32_t *samples, int32_t *pts, double *cost, int32_t *parent, int64_t *stats)
{
    for (int p = 0; p < nplans; ++p) {
        int rc = orc_plan(kind, ogs + og_offset[p], W, H, n, r_rewire, r_goal, starts + 2 * p,
                          goals + 2 * p, samples + (size_t)2 * n * p, NULL, NULL,
                          pts + (size_t)2 * (n + 1) * p, cost + (size_t)(n + 1) * p,
                          parent + (size_t)(n + 1) * p, stats + (size_t)ST_COUNT * p, NULL);
        if (rc) return rc;
    }
    return 0;
}
