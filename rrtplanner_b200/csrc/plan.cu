// K7: one thread block per plan.  The tree's vertex array (packed int16 pairs), the radius-set
// scratch list and -- when it fits -- the plan's bit-packed occupancy grid live in shared memory
// for the whole plan; cost-to-come and parents live in the output arrays (global, L1/L2 resident).
//
// Replaces the three plan() loops of the reference (rrt.py:418-437, 498-548, 690-748), go2goal
// (rrt.py:284-332) and the primitives they call (near :131-155, within :157-181, collisionfree
// :183-229, default costfn :70-78, informed sampler :579-633).  Results are identical to the
// reference run on the same sample stream with its two unstable argsorts pinned to "lowest index
// first" (SURVEY.md section 8(c)); tests/test_plan_gpu.py checks that bit for bit.
//
// Per iteration (block-wide, 1 barrier if the sample is a duplicate / the tree is full, else 2):
//   scan      every thread visits 4 vertices per 128-bit shared load: exact integer d^2, running
//             (min d^2, lowest index), and -- RRT*/informed -- one radius-set membership bit per
//             visited vertex in a word private to the thread (no atomics, no shared list).
//             Unfilled slots hold a far-away sentinel, so the loop has no tail predicates.
//             Vertex 0 (the start) is kept in registers and merged after the scan so that
//             "d^2 == 0 among vertices >= 1" is exactly the reference's `sampled` set test.
//   barrier 1 per-warp minima are combined by every warp redundantly (REDUX, no extra barrier).
//   gate      warp 0 walks nearest -> sample (rrt.py:424/506/706) and publishes the verdict ...
//   choose    ... while every warp already evaluates its own radius-set members: cost in FP64
//             (exact d^2, __dsqrt_rn, __dadd_rn); members that beat the nearest vertex's cost and
//             the warp's best so far are walked warp-cooperatively.
//   barrier 2 verdict read; per-warp (cost, index) minima combined with three REDUX; thread 0
//             stores cost/parent of the new vertex.
//
// The reference's "rewire" block (rrt.py:532-546, 732-742) tests vcosts[vn] + d < vcosts[vn] and
// can never fire with the default cost function (oracle/rrt_oracle.py counts it: always 0), so
// it has no device counterpart.
#include <math_constants.h>

#include <cstdlib>

#include "common.cuh"

namespace rrtk {

constexpr int kMaxWarps = 16;

struct PlanParams {
    const uint32_t *bits;
    size_t words_per_grid;
    int W, H, TY;
    const rrtk_plan_desc *plans;
    int n;
    uint32_t r2_excl;     // ring test  d2 < r2_excl   (= ceil(r_rewire^2), capped at 2^30)
    double r_goal;
    const short2 *samples;
    const double2 *balls;
    short2 *pts;
    double *cost;
    int *parent;
    long long *stats;
    double *ell_c;
};

__device__ __forceinline__ unsigned warp_min_u32(unsigned v) { return __reduce_min_sync(RRTK_FULL, v); }

// informed ellipse sample, rrt.py:589-599 + 615-625 (rotation computed on the host, rrt.py:601-613)
__device__ __forceinline__ void ellipse_sample(int W, int H, const double rot[4], int sx, int sy, int gx, int gy,
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

template <int KIND, bool GRID_SMEM>
__global__ void plan_kernel(PlanParams P)
{
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint2 s_near[2][kMaxWarps];            // per-warp (min d2, index), double-buffered
    __shared__ double s_bestc[kMaxWarps];
    __shared__ int s_bestv[kMaxWarps];
    __shared__ int s_gate;
    __shared__ unsigned long long s_goalc;
    __shared__ int s_goalv;
    __shared__ unsigned long long s_checks, s_cells, s_ring_total;

    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nw = T >> 5;
    const int plan = blockIdx.x;
    const int n = P.n;
    const int npad = (n + 1 + 3) & ~3;

    const int log2T = 31 - __clz(T);                                  // T is a power of two (checked by the launcher)
    const int hit_words_max = ((((npad >> 2) + T - 1) >> log2T) + 7) >> 3;
    uint32_t *s_pts = smem;                                           // npad words
    uint32_t *s_hit = smem + npad;                                    // hit_words_max * T words, thread-private
    uint32_t *s_grid = s_hit + ((hit_words_max * T + 3) & ~3);        // grid words (GRID_SMEM)

    const rrtk_plan_desc d = P.plans[plan];
    const uint32_t *gbits = P.bits + (size_t)d.world * P.words_per_grid;
    const int sx = d.start_x, sy = d.start_y, gx = d.goal_x, gy = d.goal_y;
    const uint32_t startp = pack_xy(sx, sy);

    for (int i = tid; i < npad; i += T) s_pts[i] = RRTK_FAR_VERTEX;
    if (GRID_SMEM) {
        const uint4 *src = reinterpret_cast<const uint4 *>(gbits);
        uint4 *dst = reinterpret_cast<uint4 *>(s_grid);
        for (size_t i = tid; i < P.words_per_grid / 4; i += T) dst[i] = __ldg(src + i);
    }
    if (tid == 0) {
        s_gate = 0;
        s_checks = s_cells = s_ring_total = 0ull;
    }
    __syncthreads();

    SharedGrid sg{s_grid};
    GlobalGrid gg{gbits};
    const int TY = P.TY;
#define WALK(ax_, ay_, bx_, by_) \
    (GRID_SMEM ? warp_first_hit(sg, TY, ax_, ay_, bx_, by_, lane) : warp_first_hit(gg, TY, ax_, ay_, bx_, by_, lane))

    double *cost = P.cost + (size_t)plan * (n + 1);
    int *parent = P.parent + (size_t)plan * (n + 1);
    const short2 *samples = P.samples + (size_t)plan * n;
    const double2 *balls = (KIND == RRTK_INFORMED) ? P.balls + (size_t)plan * n : nullptr;
    double *ell_c = (KIND == RRTK_INFORMED) ? P.ell_c + (size_t)plan * (n + 1) : nullptr;

    if (KIND == RRTK_INFORMED)
        for (int i = tid; i <= n; i += T) ell_c[i] = CUDART_NAN;

    // block-uniform state, replicated in every thread
    int j = 1;
    uint32_t lastp = RRTK_FAR_VERTEX;   // STANDARD: newest vertex, kept in registers (see below)
    int lastv = -1;
    double lastc = 0.0;
    bool have_sol = false;              // INFORMED: running least_cost over vsoln (rrt.py:627-633)
    int vsol = 0;
    double csol = 0.0;
    long long first_sol = -1, ell_iters = 0, nn_pairs = 0, accepted = 0;
    unsigned ring_count = 0;                // radius-set members seen by this thread
    unsigned my_checks = 0, my_cells = 0;   // per-warp counters (lane 0 meaningful)

    short2 snext = samples[0];
    for (int it = 0; it < n; ++it) {
        if (KIND != RRTK_INFORMED && j == n) break;    // tree full: every later sample is rejected
        const short2 s = snext;
        if (it + 1 < n) snext = samples[it + 1];
        int x = s.x, y = s.y;
        if (KIND == RRTK_INFORMED && have_sol && P.balls == nullptr) break;   // probe run: stop at first solution
        if (KIND == RRTK_INFORMED && have_sol) {
            const uint32_t pv = s_pts[vsol];
            const double c = reach_cost(csol, dist2(pv, gx, gy));                 // rrt.py:698-699
            ellipse_sample(P.W, P.H, d.rot, sx, sy, gx, gy, c, balls[it], x, y);
            if (tid == 0) ell_c[j] = c;                                           // rrt.py:701
            ++ell_iters;
        }
        const int par = it & 1;

        // ---- scan: nearest + radius set over vertices 1 .. j-1 ------------------------------
        // Radius-set membership is recorded as one bit per visited vertex in a word private to
        // the visiting thread (4 bits per step, 8 steps per word): no atomics, no shared list.
        uint32_t bd = 0xffffffffu, bi = 0;
        const int nquads = (j + 3) >> 2;
        const int hit_words = (((nquads + T - 1) >> log2T) + 7) >> 3;     // block-uniform
        {
            const uint4 *q4 = reinterpret_cast<const uint4 *>(s_pts);
            const uint32_t r2x = P.r2_excl;
            uint32_t hits = 0;
            int step = 0;
#define VISIT(word_, v_, bit_)                                               \
    {                                                                        \
        const uint32_t dd = dist2(word_, x, y);                              \
        if (dd < bd) { bd = dd; bi = (v_); }                                 \
        if (KIND != RRTK_STANDARD && dd < r2x) nib |= (bit_);                \
    }
#pragma unroll 2
            for (int q = tid; q < nquads; q += T, ++step) {
                const uint4 w = q4[q];
                const int v = q << 2;
                uint32_t nib = 0;
                VISIT(w.x, v, 1u)
                VISIT(w.y, v + 1, 2u)
                VISIT(w.z, v + 2, 4u)
                VISIT(w.w, v + 3, 8u)
                if (KIND != RRTK_STANDARD) {
                    hits |= nib << ((step & 7) << 2);
                    if ((step & 7) == 7) { s_hit[(step >> 3) * T + tid] = hits; hits = 0; }
                }
            }
#undef VISIT
            if (KIND != RRTK_STANDARD)
                for (int wd = step >> 3; wd < hit_words; ++wd) { s_hit[wd * T + tid] = hits; hits = 0; }
        }
        const uint32_t d2s = dist2(startp, x, y);
        {   // warp minimum, lowest index among equals
            const uint32_t wd = warp_min_u32(bd);
            const uint32_t wi = warp_min_u32(bd == wd ? bi : 0xffffffffu);
            if (lane == 0) s_near[par][warp] = make_uint2(wd, wi);
        }
        __syncthreads();                                                   // ---- barrier 1
        {
            const uint2 e = lane < nw ? s_near[par][lane] : make_uint2(0xffffffffu, 0xffffffffu);
            bd = warp_min_u32(e.x);
            bi = warp_min_u32(e.x == bd ? e.y : 0xffffffffu);
        }
        nn_pairs += j;
        bool dup = (bd == 0);                       // an accepted sample (vertex >= 1) at this cell
        if (KIND == RRTK_STANDARD && lastv > 0) {   // newest vertex may not be visible in smem yet
            const uint32_t dl = dist2(lastp, x, y);
            dup |= (dl == 0);
            if (dl < bd) { bd = dl; bi = lastv; }
        }
        if (d2s <= bd) { bd = d2s; bi = 0; }        // vertex 0 wins ties (lowest index)
        const int vnear = (int)bi;
        const uint32_t pnear = vnear == 0 ? startp : (KIND == RRTK_STANDARD && vnear == lastv ? lastp : s_pts[vnear]);
        // the reference walks nearest -> sample before looking at the other two gate terms
        // (rrt.py:424-425 / 506-507 / 706-707); they do not depend on the walk, so test them first
        if (dup || j == n) continue;

        const double cnear = vnear == 0 ? 0.0 : (KIND == RRTK_STANDARD && vnear == lastv ? lastc : cost[vnear]);
        const double c0 = reach_cost(cnear, bd);
        int vbest = vnear;
        double cbest = c0;
        const uint32_t pnew = pack_xy(x, y);

        if (KIND == RRTK_STANDARD) {
            const int hit = WALK(px(pnear), py(pnear), x, y);              // every warp, redundantly: no 2nd barrier
            if (warp == 0) { my_checks += 1; my_cells += cells_tested(hit); }
            if (hit >= 0) continue;
            if (tid == 0) { s_pts[j] = pnew; cost[j] = c0; parent[j] = vnear; }
            lastp = pnew; lastv = j; lastc = c0;
        } else {
            // ---- gate (warp 0 only) overlapped with choose-parent (all warps): rrt.py:506-521 ----
            double wc = CUDART_INF;                 // this warp's best (cost, vertex)
            int wv = 0x7fffffff;
            bool warp_active = true;
            if (warp == 0) {
                const int hit = WALK(px(pnear), py(pnear), x, y);
                my_checks += 1; my_cells += cells_tested(hit);
                if (lane == 0) {
                    s_gate = hit < 0;
                    if (hit < 0) s_pts[j] = pnew;   // nobody reads slot j before barrier 2
                }
                warp_active = hit < 0;
            }
            unsigned ring_it = 0;
            if (warp_active && warp == nw - 1 && d2s < P.r2_excl) {            // vertex 0 lives in registers
                const double cn = reach_cost(0.0, d2s);
                if (cn < c0) {
                    const int h = WALK(sx, sy, x, y);
                    my_checks += 1; my_cells += cells_tested(h);
                    if (h < 0) { wc = cn; wv = 0; }
                }
                if (lane == 0) ring_it = 1;
            }
            for (int wd = 0; warp_active && wd < hit_words; ++wd) {
                uint32_t bits = s_hit[wd * T + tid];             // written by this very thread
                ring_it += __popc(bits);
                while (__any_sync(RRTK_FULL, bits != 0)) {
                    const bool has = bits != 0;
                    const int b = has ? __ffs(bits) - 1 : 0;
                    bits &= bits - 1;
                    const int v = ((tid + (wd * 8 + (b >> 2)) * T) << 2) + (b & 3);
                    const uint32_t p = has ? s_pts[v] : 0u;
                    double cn = CUDART_INF;
                    if (has) cn = reach_cost(cost[v], dist2(p, x, y));
                    unsigned m = __ballot_sync(RRTK_FULL, cn < c0);
                    while (m) {
                        const int l = __ffs(m) - 1;
                        m &= m - 1;
                        const double cv = __shfl_sync(RRTK_FULL, cn, l);
                        const int vv = __shfl_sync(RRTK_FULL, v, l);
                        if (cv < wc || (cv == wc && vv < wv)) {
                            const uint32_t pp = __shfl_sync(RRTK_FULL, p, l);
                            const int h = WALK(px(pp), py(pp), x, y);
                            my_checks += 1; my_cells += cells_tested(h);
                            if (h < 0) { wc = cv; wv = vv; }
                        }
                    }
                }
            }
            if (lane == 0) { s_bestc[warp] = wc; s_bestv[warp] = wv; }
            __syncthreads();                                               // ---- barrier 2
            if (!s_gate) continue;
            ring_count += ring_it;
            {   // minimum (cost, vertex) over the warps: costs are positive doubles, so their bit
                // patterns order like unsigned integers
                const double c = lane < nw ? s_bestc[lane] : CUDART_INF;
                const uint32_t v = lane < nw ? (uint32_t)s_bestv[lane] : 0x7fffffffu;
                const uint32_t hi = (uint32_t)__double2hiint(c), lo = (uint32_t)__double2loint(c);
                const uint32_t mhi = warp_min_u32(hi);
                const uint32_t mlo = warp_min_u32(hi == mhi ? lo : 0xffffffffu);
                const uint32_t mv = warp_min_u32(hi == mhi && lo == mlo ? v : 0xffffffffu);
                if (mv != 0x7fffffffu) { vbest = (int)mv; cbest = __hiloint2double((int)mhi, (int)mlo); }
            }
            if (tid == 0) { cost[j] = cbest; parent[j] = vbest; }          // rrt.py:524-529
        }
        if (KIND == RRTK_INFORMED) {
            const uint32_t dg = dist2(pnew, gx, gy);
            if (__dsqrt_rn((double)dg) < P.r_goal) {                       // rrt.py:744-745
                if (!have_sol) first_sol = it;
                if (!have_sol || cbest < csol) { csol = cbest; vsol = j; }
                have_sol = true;
            }
        }
        ++accepted;
        ++j;
    }

    // ---- goal connection: rrt.py:284-332, ascending (cost, index), filled vertices only ------
    if (tid == 0) { s_goalc = 0x7ff0000000000000ull; s_goalv = 0x7fffffff; }
    __syncthreads();
    for (int base = warp * 32; base < j; base += nw * 32) {
        const int v = base + lane;
        const bool valid = v < j;
        const uint32_t p = (!valid || v == 0) ? startp : s_pts[v];
        double cg = CUDART_INF;
        if (valid) cg = reach_cost(v == 0 ? 0.0 : cost[v], dist2(p, gx, gy));
        unsigned m = __ballot_sync(RRTK_FULL, valid && cg < __longlong_as_double(*(volatile unsigned long long *)&s_goalc));
        while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            const double cv = __shfl_sync(RRTK_FULL, cg, l);
            const uint32_t pp = __shfl_sync(RRTK_FULL, p, l);
            if (cv < __longlong_as_double(*(volatile unsigned long long *)&s_goalc)) {
                const int h = WALK(px(pp), py(pp), gx, gy);
                my_checks += 1; my_cells += cells_tested(h);
                if (h < 0 && lane == 0) atomicMin(&s_goalc, (unsigned long long)__double_as_longlong(cv));
            }
        }
    }
    __syncthreads();
    const unsigned long long cstar_bits = s_goalc;
    const bool reachable = cstar_bits != 0x7ff0000000000000ull;
    if (reachable) {   // lowest index among vertices with exactly the minimum cost and a free walk
        for (int base = warp * 32; base < j; base += nw * 32) {
            const int v = base + lane;
            const bool valid = v < j;
            const uint32_t p = (!valid || v == 0) ? startp : s_pts[v];
            double cg = CUDART_INF;
            if (valid) cg = reach_cost(v == 0 ? 0.0 : cost[v], dist2(p, gx, gy));
            unsigned m = __ballot_sync(RRTK_FULL, valid && (unsigned long long)__double_as_longlong(cg) == cstar_bits);
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t pp = __shfl_sync(RRTK_FULL, p, l);
                const int h = WALK(px(pp), py(pp), gx, gy);
                my_checks += 1; my_cells += cells_tested(h);
                if (h < 0 && lane == 0) atomicMin(&s_goalv, base + l);
            }
        }
    }
    ring_count = __reduce_add_sync(RRTK_FULL, ring_count);
    if (lane == 0) {
        atomicAdd(&s_checks, (unsigned long long)my_checks);
        atomicAdd(&s_cells, (unsigned long long)my_cells);
        atomicAdd(&s_ring_total, (unsigned long long)ring_count);
    }
    __syncthreads();

    // ---- outputs -------------------------------------------------------------------------------
    const int vparent = s_goalv;
    const bool found = reachable && vparent != 0x7fffffff;
    const int top = found ? j + 1 : j;     // rows holding real vertices
    short2 *opts = P.pts + (size_t)plan * (n + 1);
    for (int v = tid; v <= n; v += T) {
        short2 o = make_short2(-32768, -32768);
        if (v < j) {
            const uint32_t p = v == 0 ? startp : s_pts[v];
            o = make_short2((short)px(p), (short)py(p));
        } else if (v == j && found) {
            o = make_short2((short)gx, (short)gy);
        }
        opts[v] = o;
        if (v >= top) { cost[v] = CUDART_INF; parent[v] = -1; }
    }
    if (tid == 0) {
        cost[0] = 0.0;
        parent[0] = -1;
        if (found) { cost[j] = __longlong_as_double((long long)cstar_bits); parent[j] = vparent; }
        long long *st = P.stats + (size_t)plan * RRTK_STAT_COUNT;
        st[RRTK_STAT_J] = j;
        st[RRTK_STAT_VGOAL] = found ? j : 0;
        st[RRTK_STAT_FOUND] = found ? 1 : 0;
        st[RRTK_STAT_CHECKS] = (long long)s_checks;
        st[RRTK_STAT_CELLS] = (long long)s_cells;
        st[RRTK_STAT_FIRST_SOL_ITER] = first_sol;
        st[RRTK_STAT_ELL_ITERS] = ell_iters;
        st[RRTK_STAT_NN_PAIRS] = nn_pairs;
        st[RRTK_STAT_RING_MEMBERS] = (long long)s_ring_total;
        st[RRTK_STAT_ACCEPTED] = accepted;
        st[RRTK_STAT_RESERVED0] = 0;
        st[RRTK_STAT_RESERVED1] = 0;
    }
#undef WALK
}

static size_t plan_smem_bytes(int W, int H, int n, int threads, bool grid_smem)
{
    const size_t npad = (size_t)((n + 1 + 3) & ~3);
    const size_t hit_words = ((((npad >> 2) + threads - 1) / threads) + 7) >> 3;
    size_t words = npad + ((hit_words * threads + 3) & ~(size_t)3);
    if (grid_smem) words += grid_words(W, H);
    return words * 4;
}

// Where does the bit grid live?  Shared memory gives the shortest walk latency, but the tree
// (4 B / vertex) is what must stay on chip; when staging the grid as well would lower the number
// of resident plan blocks per SM, the grid is left in global memory (read-only path, L1/L2
// resident: one tile = one 128-byte line).  RRTK_GRID_SMEM=0/1 overrides for experiments.
static int blocks_by_smem(size_t bytes, int threads, int sm_smem)
{
    int by_smem = (int)((size_t)sm_smem / (bytes + 1024 + 1024));   // + static + per-block reservation
    int by_threads = 2048 / threads;
    int r = by_smem < by_threads ? by_smem : by_threads;
    return r < 1 ? 1 : (r > 32 ? 32 : r);
}
static bool choose_grid_smem(int W, int H, int n, int threads, int optin, int sm_smem, size_t *bytes)
{
    const size_t budget = (size_t)optin - 2048;
    const size_t with_grid = plan_smem_bytes(W, H, n, threads, true);
    const size_t without = plan_smem_bytes(W, H, n, threads, false);
    bool in_smem = with_grid <= budget && blocks_by_smem(with_grid, threads, sm_smem) >= blocks_by_smem(without, threads, sm_smem);
    const char *force = getenv("RRTK_GRID_SMEM");
    if (force && force[0] == '0') in_smem = false;
    if (force && force[0] == '1' && with_grid <= budget) in_smem = true;
    *bytes = in_smem ? with_grid : without;
    return in_smem;
}

template <int KIND>
static int launch_kind(const PlanParams &P, int nplans, int W, int H, int n, int threads, int optin, int sm_smem, cudaStream_t st)
{
    size_t bytes = 0;
    const bool in_smem = choose_grid_smem(W, H, n, threads, optin, sm_smem, &bytes);
    if (bytes > (size_t)optin - 2048) {   // static shared (slots, counters) is < 1 KB; keep 2 KB headroom
        set_error("plan does not fit shared memory: n=%d needs %zu bytes, device allows %zu", n, bytes, (size_t)optin - 2048);
        return RRTK_ERR_CAPACITY;
    }
    if (in_smem) {
        RRTK_CUDA(cudaFuncSetAttribute(plan_kernel<KIND, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        plan_kernel<KIND, true><<<nplans, threads, bytes, st>>>(P);
    } else {
        RRTK_CUDA(cudaFuncSetAttribute(plan_kernel<KIND, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        plan_kernel<KIND, false><<<nplans, threads, bytes, st>>>(P);
    }
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

int plan_default_threads(int n) { return n >= 1024 ? 128 : 64; }

int plan_footprint(int kind, int W, int H, int n, int threads, int optin, int sm_smem, int *smem_bytes, int *blocks_per_sm)
{
    (void)kind;
    if (threads <= 0) threads = plan_default_threads(n);
    size_t b = 0;
    choose_grid_smem(W, H, n, threads, optin, sm_smem, &b);
    if (b > (size_t)optin - 2048) return RRTK_ERR_CAPACITY;
    if (smem_bytes) *smem_bytes = (int)b;
    if (blocks_per_sm) *blocks_per_sm = blocks_by_smem(b, threads, sm_smem);
    return RRTK_OK;
}

int plan_launch(int kind, const uint32_t *d_bits, int W, int H, const rrtk_plan_desc *d_plans, int nplans, int n,
                double r_rewire, double r_goal, const int16_t *d_samples, const double *d_balls, int16_t *d_pts,
                double *d_cost, int32_t *d_parent, int64_t *d_stats, double *d_ell_c, int threads, int optin,
                int sm_smem, cudaStream_t st)
{
    if (threads <= 0) threads = plan_default_threads(n);
    if (threads < 32 || threads > 32 * kMaxWarps || (threads & (threads - 1))) {
        set_error("threads must be a power of two in [32, %d]", 32 * kMaxWarps);
        return RRTK_ERR_INVALID;
    }
    PlanParams P;
    P.bits = d_bits;
    P.words_per_grid = grid_words(W, H);
    P.W = W; P.H = H; P.TY = tiles_y(H);
    P.plans = d_plans;
    P.n = n;
    double rr = r_rewire * r_rewire;
    double lim = ceil(rr);
    P.r2_excl = (kind == RRTK_STANDARD) ? 0u : (lim >= 1073741824.0 ? 1073741824u : (lim <= 0.0 ? 0u : (uint32_t)lim));
    P.r_goal = r_goal;
    P.samples = reinterpret_cast<const short2 *>(d_samples);
    P.balls = reinterpret_cast<const double2 *>(d_balls);
    P.pts = reinterpret_cast<short2 *>(d_pts);
    P.cost = d_cost;
    P.parent = d_parent;
    P.stats = reinterpret_cast<long long *>(d_stats);
    P.ell_c = d_ell_c;
    switch (kind) {
        case RRTK_STANDARD: return launch_kind<RRTK_STANDARD>(P, nplans, W, H, n, threads, optin, sm_smem, st);
        case RRTK_STAR: return launch_kind<RRTK_STAR>(P, nplans, W, H, n, threads, optin, sm_smem, st);
        case RRTK_INFORMED: return launch_kind<RRTK_INFORMED>(P, nplans, W, H, n, threads, optin, sm_smem, st);
    }
    set_error("unknown planner kind %d", kind);
    return RRTK_ERR_INVALID;
}

// ---- root -> goal paths (RRT.route2gv, rrt.py:87-107, on a tree) ------------------------------
__global__ void paths_kernel(const int *parent, const long long *stats, int nplans, int n, int cap, int *path, int *len)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nplans) return;
    const int *par = parent + (size_t)p * (n + 1);
    int *out = path + (size_t)p * cap;
    int v = (int)stats[(size_t)p * RRTK_STAT_COUNT + RRTK_STAT_VGOAL];
    int depth = 0;
    for (int u = v; u > 0 && depth <= n; u = par[u]) ++depth;    // edges from v up to the root
    const int L = depth + 1;
    len[p] = L;
    if (L > cap) return;
    int u = v;
    for (int k = L - 1; k >= 0; --k) { out[k] = u; u = u > 0 ? par[u] : 0; }
}

int paths_launch(const int32_t *d_parent, const int64_t *d_stats, int nplans, int n, int cap, int32_t *d_path,
                 int32_t *d_len, cudaStream_t st)
{
    paths_kernel<<<(nplans + 127) / 128, 128, 0, st>>>(d_parent, reinterpret_cast<const long long *>(d_stats), nplans, n,
                                                      cap, d_path, d_len);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

}  // namespace rrtk
