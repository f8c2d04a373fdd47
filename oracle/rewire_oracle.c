/*
 * CPU oracle (plain C) for the planners the reference ADVERTISES but does not contain -- TEST
 * INFRASTRUCTURE ONLY (never linked or loaded by the product package; used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline legs through oracle/rewire_oracle.py).
 *
 *   model EUCLID + rewire : RRT* with a rewire step that can fire.  The reference's own rewire
 *                           (rrt.py:532-546) compares cost(vn -> xnew) with vcosts[vn] and therefore
 *                           never fires; this is the textbook
 *                           predicate  vcosts[vnew] + |xnew - xn| < vcosts[vn]  with the costs of the
 *                           rewired subtree updated.
 *   model DUBINS          : "Dubins Vehicle RRT Planner" / "Dubins Vehicle RRT(star) Planner"
 *                           (README.md:18-19) on the "Dubins Primitive Module" (README.md:12).
 *
 * Parity status: UNPINNED.  /root/reference holds no Dubins code and no rewire that fires, so there
 * is nothing of the reference to check these against; this file is the specification the CUDA path
 * (csrc/plan_rewire.cu, csrc/dubins.cuh) is tested against bit for bit.  What IS pinned: with model
 * EUCLID and rewire off, orc2_plan must reproduce the pinned reference trees of tests/golden/
 * (tests/test_rewire_oracle.py), which covers everything the two models share (nearest, radius set,
 * choose-parent order, duplicate gate, goal connection).
 *
 * Bit-exactness across CPU and GPU needs every floating-point step to be an IEEE-754 double
 * add / sub / mul / div / sqrt / floor in a fixed order, so the transcendental functions the Dubins
 * construction needs (atan2, sin, cos) are fixed polynomials defined HERE as part of the
 * specification (dm_* below) instead of libm calls.  Build with -ffp-contract=off.
 *
 * ---- specification --------------------------------------------------------------------------------
 * Vertex = (x, y, h): integer cell and heading index h in [0, NH), heading angle h * (2 pi / NH).
 * Edge length  len(a -> b):  EUCLID  sqrt(dx^2 + dy^2);  DUBINS  shortest of the six Dubins words
 *   LSL RSR LSR RSL RLR LRL (Shkel & Lumelsky's closed forms), ties to the first word in that order,
 *   length ((t + p) + q) * rho.
 * Edge test  free(a -> b):  EUCLID  the reference's integer line walk (rrt.py:183-229) from a to b;
 *   DUBINS  the points of the path at arc length k * ds, k = 0 .. floor(len / ds), rounded to the
 *   nearest cell (floor(v + 0.5)), plus b's own cell; a point outside the grid blocks.
 * Loop (the reference's RRT* loop rrt.py:498-548 with len / free substituted):
 *   xnew = sample i;  vnearest = Euclidean nearest (lowest index);  reject unless free(vnearest -> xnew),
 *   (x, y) not sampled before, j != n.  c0 = cost[vnearest] + len(vnearest -> xnew).
 *   choose parent: over vn in within(r) ascending, vn != vnearest, with cost[vn] + |xn - xnew| < c0
 *   (Euclidean prefilter), cn = cost[vn] + len(vn -> xnew) < running best (strict) and free(vn -> xnew).
 *   rewire (if enabled): over vn in within(r) ascending, vn != parent, with cbest + |xn - xnew| < cost[vn]
 *   and cm = cbest + len(xnew -> vn) < cost[vn] and free(xnew -> vn):  parent[vn] = vnew,
 *   elen[vn] = len, cost[vn] = cm, and every descendant d of vn gets cost[d] = cost[parent[d]] + elen[d].
 * Goal connection: ascending (cost[v] + len(v -> goal), v), first free one; none -> vgoal = 0.
 * Informed sampling (orc2_plan_informed; the sampling rule of rrt.py:690-701,744-745 on top of the loop above):
 *   a vertex accepted within Euclidean distance r_goal of the goal (strict) joins vsoln.  While vsoln is
 *   empty iteration i takes sample i.  Otherwise vb = the first minimum of cost over vsoln in the order the
 *   vertices joined (rrt.py:627-633; costs as they stand NOW, i.e. after every rewire so far),
 *   c = cost[vb] + |x_vb - x_goal|, and the (x, y) of iteration i is the ellipse point of rrt.py:589-625
 *   for c and the unit-disc draw balls[i] (rotation given by the caller), clamped to the grid and truncated;
 *   the heading stays that of sample i.  ell_c[j] = c for the running j (rrt.py:701).  balls == NULL is the
 *   probe: the plan stops after the iteration that puts the first vertex into vsoln.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MODEL_EUCLID 0
#define MODEL_DUBINS 1

enum { S2_J = 0, S2_VGOAL, S2_FOUND, S2_CHECKS, S2_ACCEPTED, S2_REWIRES, S2_PROPAGATED, S2_RING, S2_LEN_EVALS, S2_OVERFLOW, S2_ELL_ITERS,
       S2_FIRST_SOL, S2_COUNT = 12 };

/* ---- deterministic elementary functions (part of the specification) ------------------------------ */
#define DM_PI 3.14159265358979323846
#define DM_TWO_PI 6.28318530717958647692
#define DM_INV_TWO_PI 0.15915494309189533577
#define DM_HALF_PI 1.57079632679489661923
#define DM_QUARTER_PI 0.78539816339744830962
#define DM_TWO_OVER_PI 0.63661977236758134308
#define DM_TAN_PI_8 0.41421356237309504880

/* atan(z) for |z| <= tan(pi/8): z * sum_{k<16} (-1)^k z^(2k) / (2k+1)  (truncation < 2e-14), evaluated as
 * E(s^2) + s * O(s^2) with s = z^2, both halves by Horner: E holds the even k, O the odd k */
static double dm_atan_small(double z)
{
    const double s = z * z, s2 = s * s;
    double e = 1.0 / 29.0, o = -1.0 / 31.0;
    e = e * s2 + 1.0 / 25.0;  o = o * s2 - 1.0 / 27.0;
    e = e * s2 + 1.0 / 21.0;  o = o * s2 - 1.0 / 23.0;
    e = e * s2 + 1.0 / 17.0;  o = o * s2 - 1.0 / 19.0;
    e = e * s2 + 1.0 / 13.0;  o = o * s2 - 1.0 / 15.0;
    e = e * s2 + 1.0 / 9.0;   o = o * s2 - 1.0 / 11.0;
    e = e * s2 + 1.0 / 5.0;   o = o * s2 - 1.0 / 7.0;
    e = e * s2 + 1.0;         o = o * s2 - 1.0 / 3.0;
    return z * (e + s * o);
}

/* one division: atan(num/den) = pi/4 + atan((num - den) / (num + den)) above tan(pi/8) */
double dm_atan2(double y, double x)
{
    if (x == 0.0 && y == 0.0) return 0.0;
    const double ax = fabs(x), ay = fabs(y);
    const int swap = ay > ax;
    const double num = swap ? ax : ay, den = swap ? ay : ax;
    double r;
    if (num > DM_TAN_PI_8 * den) r = DM_QUARTER_PI + dm_atan_small((num - den) / (num + den));
    else r = dm_atan_small(num / den);
    if (swap) r = DM_HALF_PI - r;
    if (x < 0.0) r = DM_PI - r;
    if (y < 0.0) r = -r;
    return r;
}

/* sin and cos: quadrant k = floor(a * 2/pi + 1/2), r = a - k * pi/2, Taylor polynomials on |r| <= pi/4 */
void dm_sincos(double a, double *sn, double *cs)
{
    const double kf = floor(a * DM_TWO_OVER_PI + 0.5);
    const double r = a - kf * DM_HALF_PI;
    const double s = r * r;
    /* sin r = r * (1 - s/3! + s^2/5! - ... + s^7/15!) */
    double ps = -1.0 / 1307674368000.0;
    ps = ps * s + 1.0 / 6227020800.0;
    ps = ps * s - 1.0 / 39916800.0;
    ps = ps * s + 1.0 / 362880.0;
    ps = ps * s - 1.0 / 5040.0;
    ps = ps * s + 1.0 / 120.0;
    ps = ps * s - 1.0 / 6.0;
    ps = ps * s + 1.0;
    ps = ps * r;
    /* cos r = 1 - s/2! + s^2/4! - ... + s^8/16! */
    double pc = 1.0 / 20922789888000.0;
    pc = pc * s - 1.0 / 87178291200.0;
    pc = pc * s + 1.0 / 479001600.0;
    pc = pc * s - 1.0 / 3628800.0;
    pc = pc * s + 1.0 / 40320.0;
    pc = pc * s - 1.0 / 720.0;
    pc = pc * s + 1.0 / 24.0;
    pc = pc * s - 1.0 / 2.0;
    pc = pc * s + 1.0;
    const long long k = (long long)kf;
    switch ((int)(k & 3)) {
        case 0: *sn = ps; *cs = pc; break;
        case 1: *sn = pc; *cs = -ps; break;
        case 2: *sn = -ps; *cs = -pc; break;
        default: *sn = -pc; *cs = ps; break;
    }
}

/* angle into [0, 2 pi); results within 1e-9 of a full turn (rounding of an exact 0) snap to 0 so that no path
 * carries a spurious extra circle */
static double dm_mod2pi(double x)
{
    const double r = x - DM_TWO_PI * floor(x * DM_INV_TWO_PI);
    return (r < 0.0 || r > DM_TWO_PI - 1e-9) ? 0.0 : r;
}
/* acos through atan2: acos(v) = atan2(sqrt(1 - v^2), v), |v| <= 1 */
static double dm_acos(double v) { return dm_atan2(sqrt(1.0 - v * v), v); }

/* ---- Dubins primitive ------------------------------------------------------------------------------ */
typedef struct { int word; double t, p, q, len; } dubins_t;          /* word 0..5 = LSL RSR LSR RSL RLR LRL */

/* quantities every word shares: normalised distance and the sines / cosines of alpha, beta */
typedef struct { double d, dd, alpha, beta, sa, ca, sb, cb, cab; } dubins_in_t;

static void dubins_setup(int dx, int dy, int h0, int h1, int NH, double rho, dubins_in_t *g)
{
    const double dth = DM_TWO_PI / (double)NH;
    const double th0 = (double)h0 * dth, th1 = (double)h1 * dth;
    double s0, c0, s1, c1, sd, cd;
    dm_sincos(th0, &s0, &c0);
    dm_sincos(th1, &s1, &c1);
    dm_sincos((double)(((h0 - h1) % NH + NH) % NH) * dth, &sd, &cd);
    const double D = sqrt((double)((int64_t)dx * dx + (int64_t)dy * dy));
    g->d = D * (1.0 / rho);
    g->dd = g->d * g->d;
    double cphi = 1.0, sphi = 0.0;
    if (D > 0.0) { const double inv = 1.0 / D; cphi = (double)dx * inv; sphi = (double)dy * inv; }
    const double phi = dm_atan2((double)dy, (double)dx);
    g->alpha = dm_mod2pi(th0 - phi);
    g->beta = dm_mod2pi(th1 - phi);
    g->sa = s0 * cphi - c0 * sphi; g->ca = c0 * cphi + s0 * sphi;
    g->sb = s1 * cphi - c1 * sphi; g->cb = c1 * cphi + s1 * sphi;
    g->cab = cd;
}

/* (t, p, q) of word w (0..5 = LSL RSR LSR RSL RLR LRL); 0 when the word does not exist */
static int dubins_word(const dubins_in_t *g, int w, double *ot, double *op, double *oq)
{
    const double d = g->d, dd = g->dd, alpha = g->alpha, beta = g->beta;
    const double sa = g->sa, ca = g->ca, sb = g->sb, cb = g->cb, cab = g->cab;
    double t, p, q, tmp, psq;
    switch (w) {
        case 0: /* LSL */
            psq = 2.0 + dd - 2.0 * cab + 2.0 * d * (sa - sb);
            if (psq < 0.0) return 0;
            tmp = dm_atan2(cb - ca, d + sa - sb);
            t = dm_mod2pi(tmp - alpha); p = sqrt(psq); q = dm_mod2pi(beta - tmp);
            break;
        case 1: /* RSR */
            psq = 2.0 + dd - 2.0 * cab + 2.0 * d * (sb - sa);
            if (psq < 0.0) return 0;
            tmp = dm_atan2(ca - cb, d - sa + sb);
            t = dm_mod2pi(alpha - tmp); p = sqrt(psq); q = dm_mod2pi(tmp - beta);
            break;
        case 2: /* LSR */
            psq = dd - 2.0 + 2.0 * cab + 2.0 * d * (sa + sb);
            if (psq < 0.0) return 0;
            p = sqrt(psq);
            tmp = dm_atan2(-ca - cb, d + sa + sb) - dm_atan2(-2.0, p);
            t = dm_mod2pi(tmp - alpha); q = dm_mod2pi(tmp - dm_mod2pi(beta));
            break;
        case 3: /* RSL */
            psq = dd - 2.0 + 2.0 * cab - 2.0 * d * (sa + sb);
            if (psq < 0.0) return 0;
            p = sqrt(psq);
            tmp = dm_atan2(ca + cb, d - sa - sb) - dm_atan2(2.0, p);
            t = dm_mod2pi(alpha - tmp); q = dm_mod2pi(beta - tmp);
            break;
        case 4: /* RLR */
            tmp = (6.0 - dd + 2.0 * cab + 2.0 * d * (sa - sb)) * 0.125;
            if (fabs(tmp) > 1.0) return 0;
            p = dm_mod2pi(DM_TWO_PI - dm_acos(tmp));
            t = dm_mod2pi(alpha - dm_atan2(ca - cb, d - sa + sb) + p * 0.5);       /* the angle RSR uses */
            q = dm_mod2pi(alpha - beta - t + p);
            break;
        default: /* LRL */
            tmp = (6.0 - dd + 2.0 * cab + 2.0 * d * (sb - sa)) * 0.125;
            if (fabs(tmp) > 1.0) return 0;
            p = dm_mod2pi(DM_TWO_PI - dm_acos(tmp));
            t = dm_mod2pi(p * 0.5 - alpha + dm_atan2(cb - ca, d + sa - sb));       /* the angle LSL uses */
            q = dm_mod2pi(dm_mod2pi(beta) - alpha - t + p);
            break;
    }
    *ot = t; *op = p; *oq = q;
    return 1;
}

/* shortest word from (0, 0, heading h0) to (dx, dy, heading h1); headings are indices into NH */
void orc2_dubins(int dx, int dy, int h0, int h1, int NH, double rho, dubins_t *out)
{
    dubins_in_t g;
    dubins_setup(dx, dy, h0, h1, NH, rho, &g);
    out->word = -1; out->len = INFINITY; out->t = out->p = out->q = 0.0;
    for (int w = 0; w < 6; ++w) {
        double t, p, q;
        if (!dubins_word(&g, w, &t, &p, &q)) continue;
        const double len = ((t + p) + q) * rho;
        if (len < out->len) { out->len = len; out->word = w; out->t = t; out->p = p; out->q = q; }
    }
}

/* all six words of one query (tests): ok[w], tpq[3w..], for q = (x0, y0, h0, x1, y1, h1) */
void orc2_dubins_all(const int32_t *q, int NH, double rho, int32_t *ok, double *tpq)
{
    dubins_in_t g;
    dubins_setup(q[3] - q[0], q[4] - q[1], q[2], q[5], NH, rho, &g);
    for (int w = 0; w < 6; ++w) ok[w] = dubins_word(&g, w, tpq + 3 * w, tpq + 3 * w + 1, tpq + 3 * w + 2);
}

/* segment kinds of the six words: 0 = left arc, 1 = straight, 2 = right arc */
static const int kSeg[6][3] = {{0, 1, 0}, {2, 1, 2}, {0, 1, 2}, {2, 1, 0}, {2, 0, 2}, {0, 2, 0}};

static void advance(double *x, double *y, double *th, int kind, double len, double rho)
{
    double s0, c0, s1, c1;
    dm_sincos(*th, &s0, &c0);
    if (kind == 1) {
        *x = *x + rho * len * c0;
        *y = *y + rho * len * s0;
    } else if (kind == 0) {
        dm_sincos(*th + len, &s1, &c1);
        *x = *x + rho * (s1 - s0);
        *y = *y + rho * (c0 - c1);
        *th = *th + len;
    } else {
        dm_sincos(*th - len, &s1, &c1);
        *x = *x + rho * (s0 - s1);
        *y = *y + rho * (c1 - c0);
        *th = *th - len;
    }
}

/* point of the path at arc length s (cells) from its start (x0, y0, heading index h0) */
void orc2_dubins_point(int x0, int y0, int h0, int NH, double rho, const dubins_t *w, double s, double *ox, double *oy, double *oth)
{
    double x = (double)x0, y = (double)y0, th = (double)h0 * (DM_TWO_PI / (double)NH);
    const double u = s * (1.0 / rho);
    const int *k = kSeg[w->word];
    if (u < w->t) {
        advance(&x, &y, &th, k[0], u, rho);
    } else {
        advance(&x, &y, &th, k[0], w->t, rho);
        const double u2 = u - w->t;
        if (u2 < w->p) {
            advance(&x, &y, &th, k[1], u2, rho);
        } else {
            advance(&x, &y, &th, k[1], w->p, rho);
            advance(&x, &y, &th, k[2], u2 - w->p, rho);
        }
    }
    *ox = x; *oy = y;
    if (oth) *oth = th;
}

static int cell_blocked(const uint8_t *og, int W, int H, double x, double y)
{
    const double fx = floor(x + 0.5), fy = floor(y + 0.5);
    if (!(fx >= 0.0 && fx < (double)W && fy >= 0.0 && fy < (double)H)) return 1;
    return og[(size_t)(int)fx * H + (int)fy] != 0;
}

/* 1 = the sampled path is free */
int orc2_dubins_free(const uint8_t *og, int W, int H, int x0, int y0, int h0, int x1, int y1, int NH, double rho, double ds,
                     const dubins_t *w)
{
    if (w->word < 0) return 0;
    const long long ns = (long long)floor(w->len / ds);
    for (long long k = 0; k <= ns; ++k) {
        double x, y;
        orc2_dubins_point(x0, y0, h0, NH, rho, w, (double)k * ds, &x, &y, NULL);
        if (cell_blocked(og, W, H, x, y)) return 0;
    }
    return og[(size_t)x1 * H + y1] == 0;
}

/* batch forms for the primitive tests: q = (x0, y0, h0, x1, y1, h1) int32 */
void orc2_dubins_batch(const int32_t *q, long nq, int NH, double rho, int32_t *word, double *tpq, double *len)
{
    for (long i = 0; i < nq; ++i) {
        dubins_t w;
        orc2_dubins(q[6 * i + 3] - q[6 * i], q[6 * i + 4] - q[6 * i + 1], q[6 * i + 2], q[6 * i + 5], NH, rho, &w);
        word[i] = w.word; tpq[3 * i] = w.t; tpq[3 * i + 1] = w.p; tpq[3 * i + 2] = w.q; len[i] = w.len;
    }
}
void orc2_dubins_free_batch(const uint8_t *og, int W, int H, const int32_t *q, long nq, int NH, double rho, double ds, uint8_t *free_out)
{
    for (long i = 0; i < nq; ++i) {
        dubins_t w;
        orc2_dubins(q[6 * i + 3] - q[6 * i], q[6 * i + 4] - q[6 * i + 1], q[6 * i + 2], q[6 * i + 5], NH, rho, &w);
        free_out[i] = (uint8_t)orc2_dubins_free(og, W, H, q[6 * i], q[6 * i + 1], q[6 * i + 2], q[6 * i + 3], q[6 * i + 4], NH, rho, ds, &w);
    }
}
void orc2_dubins_points(const int32_t *q, int NH, double rho, const double *s, long ns, double *xy)
{
    dubins_t w;
    orc2_dubins(q[3] - q[0], q[4] - q[1], q[2], q[5], NH, rho, &w);
    for (long i = 0; i < ns; ++i) orc2_dubins_point(q[0], q[1], q[2], NH, rho, &w, s[i], xy + 3 * i, xy + 3 * i + 1, xy + 3 * i + 2);
}
void orc2_math(const double *a, const double *b, long n, double *at2, double *sn, double *cs)
{
    for (long i = 0; i < n; ++i) { at2[i] = dm_atan2(a[i], b[i]); dm_sincos(a[i], sn + i, cs + i); }
}

/* ---- rrt.py:183-229 again (kept local so this file stands alone) --------------------------------- */
static int line_free(const uint8_t *og, int H, int ax, int ay, int bx, int by)
{
    int adx = abs(bx - ax), ady = abs(by - ay);
    int stepx = ax < bx ? 1 : -1, stepy = ay < by ? 1 : -1;
    int acc = adx - ady, x = ax, y = ay;
    for (;;) {
        if (og[(size_t)x * H + y]) return 0;
        if (x == bx && y == by) return 1;
        int twice = 2 * acc;
        if (twice >= -ady) { acc -= ady; x += stepx; }
        if (twice <= adx)  { acc += adx; y += stepy; }
    }
}

typedef struct {
    int model, NH, W, H;
    double rho, ds;
    const uint8_t *og;
    int64_t *stats;
} world_t;

static double euclid(const int32_t *pts, int a, int x, int y)
{
    int64_t dx = (int64_t)pts[2 * a] - x, dy = (int64_t)pts[2 * a + 1] - y;
    return sqrt((double)(dx * dx + dy * dy));
}

/* length of the edge (ax, ay, ah) -> (bx, by, bh) and, on request, whether it is free */
static double edge(const world_t *w, int ax, int ay, int ah, int bx, int by, int bh, int *is_free)
{
    w->stats[S2_LEN_EVALS]++;
    if (w->model == MODEL_EUCLID) {
        int64_t dx = (int64_t)bx - ax, dy = (int64_t)by - ay;
        if (is_free) { *is_free = line_free(w->og, w->H, ax, ay, bx, by); w->stats[S2_CHECKS]++; }
        return sqrt((double)(dx * dx + dy * dy));
    }
    dubins_t d;
    orc2_dubins(bx - ax, by - ay, ah, bh, w->NH, w->rho, &d);
    if (is_free) { *is_free = orc2_dubins_free(w->og, w->W, w->H, ax, ay, ah, bx, by, w->NH, w->rho, w->ds, &d); w->stats[S2_CHECKS]++; }
    return d.len;
}

typedef struct { double c; int v; } cand2_t;
static int cand2_cmp(const void *a, const void *b)
{
    const cand2_t *p = a, *q = b;
    if (p->c < q->c) return -1;
    if (p->c > q->c) return 1;
    return (p->v > q->v) - (p->v < q->v);
}

/*
 * One plan.  star: 0 = parent is the nearest vertex (RRT), 1 = choose parent within r_rewire.
 * rewire: 0 = none (the reference's behaviour), 1 = as specified above.
 * samples: n x (x, y, h) int32; start / goal: (x, y, h).  Outputs have n + 1 rows: pts (x, y) int32
 * (INT32_MIN unfilled), head int32 (-1 unfilled), cost, elen (length of the edge from the parent;
 * 0 for the root, +inf unfilled), parent.
 */
/* rrt.py:589-599 + 615-625: the informed ellipse point for budget c and unit-disc draw ball (same operation order as
 * oracle/rrt_oracle.c:ellipse_point and the device's ellipse_sample) */
static void ellipse_xy(int W, int H, const double rot[4], const int32_t *s, const int32_t *g, double c, const double *ball, int *ox, int *oy)
{
    const double cx = (double)(s[0] + g[0]) / 2.0, cy = (double)(s[1] + g[1]) / 2.0;
    const double r1 = c / 2.0;
    const int64_t ddx = (int64_t)s[0] - g[0], ddy = (int64_t)s[1] - g[1];
    const double d2 = (double)(ddx * ddx + ddy * ddy);
    const double r2 = sqrt(fabs(c * c - d2)) / 2.0;
    const double m00 = rot[0] * r1, m01 = rot[1] * r2, m10 = rot[2] * r1, m11 = rot[3] * r2;
    const double x = (m00 * ball[0] + m01 * ball[1]) + cx;
    const double y = (m10 * ball[0] + m11 * ball[1]) + cy;
    double lx = (x < (double)(W - 1)) ? x : (double)(W - 1);
    double ly = (y < (double)(H - 1)) ? y : (double)(H - 1);
    lx = (lx > 0.0) ? lx : 0.0;
    ly = (ly > 0.0) ? ly : 0.0;
    *ox = (int)lx; *oy = (int)ly;
}

static int plan_core(int model, int star, int rewire, const uint8_t *og, int W, int H, int n, double r_rewire, int NH, double rho,
                     double ds, const int32_t *start, const int32_t *goal, const int32_t *samples, int32_t *pts, int32_t *head,
                     double *cost, double *elen, int32_t *parent, int64_t *stats, int informed, double r_goal, const double *rot,
                     const double *balls, double *ell_c)
{
    uint8_t *seen = calloc((size_t)W * H, 1);
    int32_t *ring = malloc(sizeof(int32_t) * (size_t)(n + 1));
    int32_t *first = malloc(sizeof(int32_t) * (size_t)(n + 1));   /* child lists: first child / next sibling */
    int32_t *next = malloc(sizeof(int32_t) * (size_t)(n + 1));
    int32_t *queue = malloc(sizeof(int32_t) * (size_t)(n + 1));
    cand2_t *cands = malloc(sizeof(cand2_t) * (size_t)(n + 1));
    int32_t *vsoln = malloc(sizeof(int32_t) * (size_t)(n + 1));
    int nsol = 0;
    if (!seen || !ring || !first || !next || !queue || !cands || !vsoln) return -1;
    if (informed && ell_c) for (int i = 0; i <= n; ++i) ell_c[i] = NAN;
    for (int i = 0; i <= n; ++i) {
        pts[2 * i] = pts[2 * i + 1] = INT32_MIN; head[i] = -1;
        cost[i] = INFINITY; elen[i] = INFINITY; parent[i] = -1; first[i] = -1; next[i] = -1;
    }
    memset(stats, 0, sizeof(int64_t) * S2_COUNT);
    stats[S2_FIRST_SOL] = -1;
    world_t w = {model, NH, W, H, rho, ds, og, stats};
    pts[0] = start[0]; pts[1] = start[1]; head[0] = start[2]; cost[0] = 0.0; elen[0] = 0.0;
    int j = 1;
    const double rr = r_rewire * r_rewire;

    for (int i = 0; i < n; ++i) {
        int x = samples[3 * i], y = samples[3 * i + 1];
        const int h = samples[3 * i + 2];
        if (informed && nsol > 0) {
            if (!balls) break;                                       /* probe */
            int vb = vsoln[0];
            for (int k = 1; k < nsol; ++k) if (cost[vsoln[k]] < cost[vb]) vb = vsoln[k];
            const double c = cost[vb] + euclid(pts, vb, goal[0], goal[1]);
            ellipse_xy(W, H, rot, start, goal, c, balls + 2 * i, &x, &y);
            if (ell_c) ell_c[j] = c;
            stats[S2_ELL_ITERS]++;
        }
        int vnear = 0;
        int64_t bd = INT64_MAX;
        for (int v = 0; v < j; ++v) {
            int64_t dx = (int64_t)pts[2 * v] - x, dy = (int64_t)pts[2 * v + 1] - y;
            int64_t d = dx * dx + dy * dy;
            if (d < bd) { bd = d; vnear = v; }
        }
        int ok;
        const double l0 = edge(&w, pts[2 * vnear], pts[2 * vnear + 1], head[vnear], x, y, h, &ok);
        if (!ok || seen[(size_t)x * H + y] || j == n) continue;
        seen[(size_t)x * H + y] = 1;
        const double c0 = cost[vnear] + l0;
        int vbest = vnear, m = 0;
        double cbest = c0, lbest = l0;
        if (star) {
            for (int v = 0; v < j; ++v) {
                int64_t dx = (int64_t)pts[2 * v] - x, dy = (int64_t)pts[2 * v + 1] - y;
                if ((double)(dx * dx + dy * dy) < rr) ring[m++] = v;
            }
            stats[S2_RING] += m;
            for (int k = 0; k < m; ++k) {
                const int vn = ring[k];
                if (vn == vnear) continue;
                if (!(cost[vn] + euclid(pts, vn, x, y) < c0)) continue;
                const double lc = edge(&w, pts[2 * vn], pts[2 * vn + 1], head[vn], x, y, h, NULL);
                const double cn = cost[vn] + lc;
                if (cn < cbest) {
                    int fr;
                    edge(&w, pts[2 * vn], pts[2 * vn + 1], head[vn], x, y, h, &fr);
                    if (fr) { vbest = vn; cbest = cn; lbest = lc; }
                }
            }
        }
        pts[2 * j] = x; pts[2 * j + 1] = y; head[j] = h; cost[j] = cbest; elen[j] = lbest; parent[j] = vbest;
        next[j] = first[vbest]; first[vbest] = j;
        if (star && rewire) {
            for (int k = 0; k < m; ++k) {
                const int vn = ring[k];
                if (vn == vbest) continue;
                if (!(cbest + euclid(pts, vn, x, y) < cost[vn])) continue;
                int fr;
                const double lr = edge(&w, x, y, h, pts[2 * vn], pts[2 * vn + 1], head[vn], NULL);
                const double cm = cbest + lr;
                if (!(cm < cost[vn])) continue;
                edge(&w, x, y, h, pts[2 * vn], pts[2 * vn + 1], head[vn], &fr);
                if (!fr) continue;
                /* unlink vn from its old parent's child list, hang it under the new vertex */
                const int op = parent[vn];
                if (first[op] == vn) first[op] = next[vn];
                else { int c = first[op]; while (next[c] != vn) c = next[c]; next[c] = next[vn]; }
                next[vn] = first[j]; first[j] = vn;
                parent[vn] = j; elen[vn] = lr; cost[vn] = cm;
                stats[S2_REWIRES]++;
                int qh = 0, qt = 0;
                queue[qt++] = vn;
                while (qh < qt) {
                    const int u = queue[qh++];
                    for (int c = first[u]; c >= 0; c = next[c]) { cost[c] = cost[u] + elen[c]; queue[qt++] = c; stats[S2_PROPAGATED]++; }
                }
            }
        }
        stats[S2_ACCEPTED]++;
        if (informed && euclid(pts, j, goal[0], goal[1]) < r_goal) {
            vsoln[nsol++] = j;
            if (nsol == 1) stats[S2_FIRST_SOL] = i;
        }
        ++j;
    }

    for (int v = 0; v < j; ++v) {
        cands[v].c = cost[v] + edge(&w, pts[2 * v], pts[2 * v + 1], head[v], goal[0], goal[1], goal[2], NULL);
        cands[v].v = v;
    }
    qsort(cands, (size_t)j, sizeof(cand2_t), cand2_cmp);
    int vgoal = 0, found = 0;
    for (int k = 0; k < j; ++k) {
        const int v = cands[k].v;
        int fr;
        const double lg = edge(&w, pts[2 * v], pts[2 * v + 1], head[v], goal[0], goal[1], goal[2], &fr);
        if (fr) {
            vgoal = j; found = 1;
            pts[2 * j] = goal[0]; pts[2 * j + 1] = goal[1]; head[j] = goal[2];
            cost[j] = cands[k].c; elen[j] = lg; parent[j] = v;
            break;
        }
    }
    stats[S2_J] = j; stats[S2_VGOAL] = vgoal; stats[S2_FOUND] = found;
    free(seen); free(ring); free(first); free(next); free(queue); free(cands); free(vsoln);
    return 0;
}

int orc2_plan(int model, int star, int rewire, const uint8_t *og, int W, int H, int n, double r_rewire, int NH, double rho,
              double ds, const int32_t *start, const int32_t *goal, const int32_t *samples, int32_t *pts, int32_t *head,
              double *cost, double *elen, int32_t *parent, int64_t *stats)
{
    return plan_core(model, star, rewire, og, W, H, n, r_rewire, NH, rho, ds, start, goal, samples, pts, head, cost, elen, parent, stats,
                     0, 0.0, NULL, NULL, NULL);
}

/* the same loop with the informed sampling rule; rot: 2 x 2 row-major rotation (rrt.py:601-613, computed by the caller),
 * balls: n x 2 unit-disc draws or NULL (probe), ell_c: n + 1 doubles (NaN where no ellipse sample was drawn) or NULL */
int orc2_plan_informed(int model, int star, int rewire, const uint8_t *og, int W, int H, int n, double r_rewire, int NH, double rho,
                       double ds, const int32_t *start, const int32_t *goal, const int32_t *samples, int32_t *pts, int32_t *head,
                       double *cost, double *elen, int32_t *parent, int64_t *stats, double r_goal, const double *rot,
                       const double *balls, double *ell_c)
{
    return plan_core(model, star, rewire, og, W, H, n, r_rewire, NH, rho, ds, start, goal, samples, pts, head, cost, elen, parent, stats,
                     1, r_goal, rot, balls, ell_c);
}
