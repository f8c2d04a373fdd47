// K7: one thread block per plan, K samples per round (K = warps per block).
//
// Replaces the three plan() loops of the reference (rrt.py:418-437, 498-548, 690-748), go2goal
// (rrt.py:284-332) and the primitives they call (near :131-155, within :157-181, collisionfree
// :183-229, default costfn :70-78, informed sampler :579-633).  Results are identical to the
// reference run on the same sample stream with its two unstable argsorts pinned to "lowest index
// first" (SURVEY.md section 8(c)); tests/test_gpu_parity.py checks that bit for bit.
//
// On chip for the whole plan: the tree's vertices packed x | y << 16 (4 B, one LDS.128 = 4
// vertices; unfilled slots hold a far-away sentinel so scans need no tail predicates), per-thread
// radius-set membership words, one compacted candidate list per warp and -- when it does not cost
// resident blocks -- the bit-packed grid.  Cost-to-come and parents live in the output arrays.
//
// The reference's loop is sequential (iteration i sees the tree iteration i-1 left).  A round
// takes the next K samples and keeps that semantics exactly:
//   scan     all threads: each vertex quad is loaded once and compared with all K samples -- exact
//            integer d^2, running (min d^2, lowest index) per sample, one radius-set membership
//            bit per (vertex, sample) in thread-private words.
//   barrier
//   owner    warp k owns sample k and evaluates it against the tree as it stood at the start of
//            the round: combine the per-warp minima (REDUX), duplicate test, walk nearest ->
//            sample (rrt.py:424/506/706), FP64 cost via the nearest vertex, compaction of the
//            membership bits into a dense candidate list, choose-parent (rrt.py:510-521) best
//            first: cheapest candidate that beats the incumbent is walked, first free one wins.
//   barrier
//   commit   warp 0 replays the K results in sample order against the vertices accepted earlier
//            in the same round: equal cell -> duplicate; inside the radius -> extra candidate
//            (cost, walk); strictly nearer than the recorded nearest vertex, or a change of the
//            informed sampler's state -> the round is cut there and the remaining samples are
//            redone next round (rare).  Accepted vertices are appended in order.
//   barrier
// so every decision is the one the sequential loop would take.  The reference's "rewire" block
// (rrt.py:532-546, 732-742) tests vcosts[vn] + d < vcosts[vn] and can never fire with the default
// cost function (oracle/rrt_oracle.py counts it: always 0), so it has no device counterpart.
#include <cstdlib>

#include "plan_common.cuh"

namespace rrtk {

constexpr int kWideListCap = 256;     // dense candidate list per warp (uint16 entries); larger radius sets take the sparse path

// resident blocks per SM the register allocation aims for (shared memory allows 8 at n = 5000)
#ifndef RRTK_MINBLOCKS
#define RRTK_MINBLOCKS(K) ((K) == 2 ? 12 : (K) == 4 ? 8 : 2)
#endif

template <int KIND, bool GRID_SMEM, int K>
__global__ void __launch_bounds__(32 * K, RRTK_MINBLOCKS(K)) plan_wide_kernel(PlanParams P)
{
    constexpr int T = 32 * K;
    constexpr int LOG2T = (K == 2 ? 6 : K == 4 ? 7 : 8);
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint2 s_near[K][K];                    // [sample][warp] (min d2, index)
    __shared__ SampleRec s_rec[K];
    __shared__ RoundSummary s_sum;
    __shared__ short2 s_samp[K];                      // informed: ellipse samples of the round
    __shared__ double s_sampc[K];
    __shared__ unsigned long long s_goalc;
    __shared__ int s_goalv;
    __shared__ unsigned long long s_checks, s_cells;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int plan = blockIdx.x;
    const int n = P.n;
    const int npad = (n + 1 + 3) & ~3;

    const int hit_words_max = ((((npad >> 2) + T - 1) >> LOG2T) + 7) >> 3;
    uint32_t *s_pts = smem;                                               // npad words
    uint32_t *s_hit = smem + npad;                                        // [K][hit_words_max][T], thread-private words
    uint16_t *s_list = reinterpret_cast<uint16_t *>(s_hit + K * hit_words_max * T);   // [K][kWideListCap]
    uint32_t *s_grid = reinterpret_cast<uint32_t *>(s_list + K * kWideListCap);           // grid words (GRID_SMEM)
    s_grid = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(s_grid) + 15) & ~uintptr_t(15));

    const rrtk_plan_desc *dsc = P.plans + plan;
    const uint32_t *gbits = P.bits + (size_t)dsc->world * P.words_per_grid;
    const int sx = dsc->start_x, sy = dsc->start_y, gx = dsc->goal_x, gy = dsc->goal_y;
    const uint32_t startp = pack_xy(sx, sy);

    for (int i = tid; i < npad; i += T) s_pts[i] = RRTK_FAR_VERTEX;
    if (GRID_SMEM) {
        const uint4 *src = reinterpret_cast<const uint4 *>(gbits);
        uint4 *dst = reinterpret_cast<uint4 *>(s_grid);
        for (size_t i = tid; i < P.words_per_grid / 4; i += T) dst[i] = __ldg(src + i);
    }
    if (tid == 0) s_checks = s_cells = 0ull;
    __syncthreads();

    SharedGrid sg{s_grid};
    GlobalGrid gg{gbits};
    const int TY = P.TY;
#define WALK(ax_, ay_, bx_, by_) \
    (GRID_SMEM ? warp_first_hit(sg, TY, ax_, ay_, bx_, by_, lane) : warp_first_hit(gg, TY, ax_, ay_, bx_, by_, lane))

    double *cost = P.cost + (size_t)plan * (n + 1);
    int *parent = P.parent + (size_t)plan * (n + 1);
    const short2 *samples = P.samples + (size_t)plan * n;
    const double2 *balls = (KIND == RRTK_INFORMED && P.balls) ? P.balls + (size_t)plan * n : nullptr;
    double *ell_c = (KIND == RRTK_INFORMED) ? P.ell_c + (size_t)plan * (n + 1) : nullptr;
    const uint32_t r2x = P.r2_excl;

    if (KIND == RRTK_INFORMED)
        for (int i = tid; i <= n; i += T) ell_c[i] = CUDART_NAN;

    // block-uniform state, replicated in every thread (refreshed from s_sum after each round)
    int j = 1, it0 = 0;
    bool have_sol = false;              // INFORMED: running least_cost over vsoln (rrt.py:627-633)
    int vsol = 0;
    double csol = 0.0;
    long long first_sol = -1;
    // counters kept by warp 0 (commit phase) / per warp (walks)
    long long ell_iters = 0, nn_pairs = 0, ring_members = 0, accepted = 0;
    unsigned my_checks = 0, my_cells = 0;

    while (it0 < n) {
        if (KIND != RRTK_INFORMED && j == n) break;                       // tree full: every later sample is rejected
        if (KIND == RRTK_INFORMED && have_sol && balls == nullptr) break; // probe run: stop at first solution
        const bool ellipse_mode = (KIND == RRTK_INFORMED) && have_sol;

        // ---- the K samples of this round -----------------------------------------------------
        int qx[K], qy[K];
        if (ellipse_mode) {
            if (lane == 0 && it0 + warp < n) {
                const double c = reach_cost(csol, dist2(s_pts[vsol], gx, gy));            // rrt.py:698-699
                int ex, ey;
                ellipse_sample(P.W, P.H, dsc->rot, sx, sy, gx, gy, c, balls[it0 + warp], ex, ey);
                s_samp[warp] = make_short2((short)ex, (short)ey);
                s_sampc[warp] = c;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const short2 s = s_samp[it0 + k < n ? k : 0];
                qx[k] = s.x; qy[k] = s.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const short2 s = samples[min(it0 + k, n - 1)];
                qx[k] = s.x; qy[k] = s.y;
            }
        }

        // ---- scan: nearest + radius-set bits for K samples over vertices 1 .. j-1 ---------------
        const int nquads = (j + 3) >> 2;
        const int hit_words = (((nquads + T - 1) >> LOG2T) + 7) >> 3;     // block-uniform
        {
            uint32_t bd[K], bi[K], hits[K];
#pragma unroll
            for (int k = 0; k < K; ++k) { bd[k] = 0xffffffffu; bi[k] = 0; hits[k] = 0; }
            const uint4 *q4 = reinterpret_cast<const uint4 *>(s_pts);
            int step = 0;
            for (int q = tid; q < nquads; q += T, ++step) {
                const uint4 w = q4[q];
                const uint32_t wv[4] = {w.x, w.y, w.z, w.w};
                const int v = q << 2;
                const int sh = (step & 7) << 2;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int vx = px(wv[e]), vy = py(wv[e]);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int dx = vx - qx[k], dy = vy - qy[k];
                        const uint32_t dd = (uint32_t)(dx * dx) + (uint32_t)(dy * dy);
                        if (dd < bd[k]) { bd[k] = dd; bi[k] = v + e; }
                        if (KIND != RRTK_STANDARD && dd < r2x) hits[k] |= 1u << (sh + e);
                    }
                }
                if (KIND != RRTK_STANDARD && (step & 7) == 7) {
#pragma unroll
                    for (int k = 0; k < K; ++k) { s_hit[(k * hit_words_max + (step >> 3)) * T + tid] = hits[k]; hits[k] = 0; }
                }
            }
            if (KIND != RRTK_STANDARD)
                for (int wd = step >> 3; wd < hit_words; ++wd) {
#pragma unroll
                    for (int k = 0; k < K; ++k) { s_hit[(k * hit_words_max + wd) * T + tid] = hits[k]; hits[k] = 0; }
                }
#pragma unroll
            for (int k = 0; k < K; ++k) {   // warp minimum, lowest index among equals
                const uint32_t wd = warp_min_u32(bd[k]);
                const uint32_t wi = warp_min_u32(bd[k] == wd ? bi[k] : 0xffffffffu);
                if (lane == 0) s_near[k][warp] = make_uint2(wd, wi);
            }
        }
        __syncthreads();                                                   // ---- barrier: scan results visible

        // ---- owner phase: warp k evaluates sample k against the round-start tree --------------
        if (it0 + warp < n) {
            const int it = it0 + warp;
            int x, y;
            if (ellipse_mode) { const short2 s = s_samp[warp]; x = s.x; y = s.y; }
            else { const short2 s = samples[it]; x = s.x; y = s.y; }
            uint32_t bd, bi;
            {
                const uint2 e = lane < K ? s_near[warp][lane] : make_uint2(0xffffffffu, 0xffffffffu);
                bd = warp_min_u32(e.x);
                bi = warp_min_u32(e.x == bd ? e.y : 0xffffffffu);
            }
            const bool dup = (bd == 0);             // an accepted sample (vertex >= 1) at this cell
            const uint32_t d2s = dist2(startp, x, y);
            if (d2s <= bd) { bd = d2s; bi = 0; }    // vertex 0 wins ties (lowest index)
            const int vnear = (int)bi;
            const uint32_t pnear = vnear == 0 ? startp : s_pts[vnear];
            int flags = 1 | (dup ? 2 : 0);
            double c0 = 0.0, wc = CUDART_INF;
            int wv = 0x7fffffff, ring = 0;
            // the reference walks nearest -> sample before looking at the duplicate test
            // (rrt.py:424-425 / 506-507 / 706-707); the verdicts are independent, so skip the walk
            if (!dup) {
                const int hit = WALK(px(pnear), py(pnear), x, y);
                my_checks += 1; my_cells += cells_tested(hit);
                if (hit < 0) {
                    flags |= 4;
                    c0 = reach_cost(vnear == 0 ? 0.0 : cost[vnear], bd);
                    if (KIND != RRTK_STANDARD) {
                        // candidates held one per lane: while one beats the incumbent, walk the cheapest
                        auto consider = [&](bool has, int v, uint32_t p, double cn) {
                            bool live = has && cn < c0;
                            for (;;) {
                                const bool cand = live && (cn < wc || (cn == wc && v < wv));
                                if (!__any_sync(RRTK_FULL, cand)) break;
                                // positive doubles order like their bit patterns
                                const uint32_t hi = cand ? (uint32_t)__double2hiint(cn) : 0xffffffffu;
                                const uint32_t mhi = warp_min_u32(hi);
                                const uint32_t lo = (cand && hi == mhi) ? (uint32_t)__double2loint(cn) : 0xffffffffu;
                                const uint32_t mlo = warp_min_u32(lo);
                                const uint32_t vv = (cand && hi == mhi && lo == mlo) ? (uint32_t)v : 0xffffffffu;
                                const uint32_t mv = warp_min_u32(vv);
                                const int src = __ffs(__ballot_sync(RRTK_FULL, vv == mv && mv != 0xffffffffu)) - 1;
                                const uint32_t pp = __shfl_sync(RRTK_FULL, p, src);
                                const int h = WALK(px(pp), py(pp), x, y);
                                my_checks += 1; my_cells += cells_tested(h);
                                if (h < 0) { wc = __hiloint2double((int)mhi, (int)mlo); wv = (int)mv; }
                                else if (lane == src) live = false;
                            }
                        };
                        if (d2s < r2x) {            // vertex 0 lives in registers
                            ring = 1;
                            consider(lane == 0, 0, startp, reach_cost(0.0, d2s));
                        }
                        // membership words of this sample: hit_words rows of T words, K per lane per row
                        const uint32_t *hw = s_hit + (size_t)warp * hit_words_max * T;
                        int mine = 0;
                        for (int wd = 0; wd < hit_words; ++wd)
#pragma unroll
                            for (int c = 0; c < K; ++c) mine += __popc(hw[wd * T + c * 32 + lane]);
                        const int total = __reduce_add_sync(RRTK_FULL, mine);
                        ring += total;
                        if (total <= kWideListCap) {
                            // compaction: exclusive prefix of the per-lane counts, then every lane lists its members
                            int incl = mine;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const int t = __shfl_up_sync(RRTK_FULL, incl, o);
                                if (lane >= o) incl += t;
                            }
                            uint16_t *list = s_list + warp * kWideListCap;
                            int pos = incl - mine;
                            for (int wd = 0; wd < hit_words; ++wd)
#pragma unroll
                                for (int c = 0; c < K; ++c) {
                                    uint32_t bits = hw[wd * T + c * 32 + lane];
                                    const int vbase = ((c * 32 + lane) + wd * 8 * T) << 2;
                                    while (bits) {
                                        const int b = __ffs(bits) - 1;
                                        bits &= bits - 1;
                                        list[pos++] = (uint16_t)(vbase + (((b >> 2) * T) << 2) + (b & 3));
                                    }
                                }
                            __syncwarp();
                            for (int base = 0; base < total; base += 32) {
                                const bool has = base + lane < total;
                                const int v = has ? (int)list[base + lane] : 0;
                                const uint32_t p = has ? s_pts[v] : 0u;
                                double cn = CUDART_INF;
                                if (has) cn = reach_cost(cost[v], dist2(p, x, y));
                                consider(has, v, p, cn);
                            }
                            __syncwarp();
                        } else {
                            // very large radius sets: walk the membership words directly, one bit per lane per step
                            for (int wd = 0; wd < hit_words; ++wd)
                                for (int c = 0; c < K; ++c) {
                                    uint32_t bits = hw[wd * T + c * 32 + lane];
                                    const int vbase = ((c * 32 + lane) + wd * 8 * T) << 2;
                                    while (__any_sync(RRTK_FULL, bits != 0)) {
                                        const bool has = bits != 0;
                                        const int b = has ? __ffs(bits) - 1 : 0;
                                        bits &= bits - 1;
                                        const int v = vbase + (((b >> 2) * T) << 2) + (b & 3);
                                        const uint32_t p = has ? s_pts[v] : 0u;
                                        double cn = CUDART_INF;
                                        if (has) cn = reach_cost(cost[v], dist2(p, x, y));
                                        consider(has, v, p, cn);
                                    }
                                }
                        }
                    }
                }
            }
            if (lane == 0) {
                SampleRec r;
                r.pnew = pack_xy(x, y); r.bd = bd; r.vnear = vnear; r.flags = flags; r.bv = wv; r.ring = ring;
                r.c0 = c0; r.bc = wc; r.ell = ellipse_mode ? s_sampc[warp] : 0.0;
                s_rec[warp] = r;
            }
        } else if (lane == 0) {
            s_rec[warp].flags = 0;
        }
        __syncthreads();                                                   // ---- barrier: K results visible

        // ---- commit phase: warp 0 replays the results in sample order -----------------------------
        if (warp == 0) {
            uint32_t newp[K];
            int newv[K];
            double newc[K];
            int nnew = 0, consumed = 0, jc = j;
            bool hs = have_sol, finished = false, cut = false;
            int vs = vsol;
            double cs = csol;
            long long fs = first_sol;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (cut || finished) continue;
                const SampleRec r = s_rec[k];
                if (!(r.flags & 1)) { finished = true; continue; }          // ran past iteration n
                if (KIND != RRTK_INFORMED && jc == n) { finished = true; continue; }
                const int x = px(r.pnew), y = py(r.pnew);
                bool reject = (r.flags & 2) || jc == n || !(r.flags & 4);
                double bc = r.bc;
                int bv = r.bv;
                // vertices accepted earlier in this round are not in the scan this sample was compared with
#pragma unroll
                for (int m = 0; m < K; ++m) {
                    if (m >= nnew || cut) continue;
                    const uint32_t du = dist2(newp[m], x, y);
                    if (du == 0) { reject = true; continue; }                // now in `sampled` (rrt.py:426/508/708)
                    if (du < r.bd) { cut = true; continue; }                  // it would be the nearest vertex: redo
                    if (KIND != RRTK_STANDARD && !reject && du < r2x) {
                        const double cn = reach_cost(newc[m], du);
                        if (cn < r.c0 && cn < bc) {                           // higher index: loses cost ties
                            const int h = WALK(px(newp[m]), py(newp[m]), x, y);
                            my_checks += 1; my_cells += cells_tested(h);
                            if (h < 0) { bc = cn; bv = newv[m]; }
                        }
                    }
                }
                if (cut) continue;                                            // this sample starts the next round
                ++consumed;
                nn_pairs += jc;
                if (KIND == RRTK_INFORMED && ellipse_mode) {
                    if (lane == 0) ell_c[jc] = r.ell;                          // rrt.py:701
                    ++ell_iters;
                }
                if (reject) continue;
                int ringm = r.ring;
                if (KIND != RRTK_STANDARD)
#pragma unroll
                    for (int m = 0; m < K; ++m)
                        if (m < nnew && dist2(newp[m], x, y) < r2x) ++ringm;
                ring_members += ringm;
                const int vbest = (bv != 0x7fffffff) ? bv : r.vnear;
                const double cbest = (bv != 0x7fffffff) ? bc : r.c0;
                if (lane == 0) { s_pts[jc] = r.pnew; cost[jc] = cbest; parent[jc] = vbest; }   // rrt.py:524-529
#pragma unroll
                for (int m = 0; m < K; ++m)
                    if (m == nnew) { newp[m] = r.pnew; newv[m] = jc; newc[m] = cbest; }
                ++nnew;
                ++accepted;
                if (KIND == RRTK_INFORMED) {
                    const uint32_t dg = dist2(r.pnew, gx, gy);
                    if (__dsqrt_rn((double)dg) < P.r_goal) {                   // rrt.py:744-745
                        const bool changed = !hs || cbest < cs;
                        if (!hs) fs = it0 + k;
                        if (changed) { cs = cbest; vs = jc; }
                        hs = true;
                        if (changed) cut = true;                               // later samples of the round used the old sampler state
                    }
                }
                ++jc;
            }
            if (lane == 0) {
                RoundSummary s;
                s.j = jc; s.consumed = consumed; s.flags = (hs ? 1 : 0) | (finished ? 2 : 0);
                s.vsol = vs; s.csol = cs; s.first_sol = fs;
                s_sum = s;
            }
        }
        __syncthreads();                                                   // ---- barrier: tree updated
        {
            const RoundSummary s = s_sum;
            j = s.j;
            it0 += s.consumed;
            have_sol = s.flags & 1;
            vsol = s.vsol; csol = s.csol; first_sol = s.first_sol;
            if (s.flags & 2) break;
        }
    }

    // ---- goal connection: rrt.py:284-332, ascending (cost, index), filled vertices only ------
    if (tid == 0) { s_goalc = 0x7ff0000000000000ull; s_goalv = 0x7fffffff; }
    __syncthreads();
    for (int base = warp * 32; base < j; base += K * 32) {
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
        for (int base = warp * 32; base < j; base += K * 32) {
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
    if (lane == 0) {
        atomicAdd(&s_checks, (unsigned long long)my_checks);
        atomicAdd(&s_cells, (unsigned long long)my_cells);
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
    if (tid == 0) {      // warp 0 carries the commit-phase counters
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
        st[RRTK_STAT_RING_MEMBERS] = ring_members;
        st[RRTK_STAT_ACCEPTED] = accepted;
        st[RRTK_STAT_RESERVED0] = 0;
        st[RRTK_STAT_RESERVED1] = 0;
    }
#undef WALK
}

static size_t wide_smem_bytes(int W, int H, int n, int threads, bool grid_smem)
{
    const size_t K = (size_t)threads / 32;
    const size_t npad = (size_t)((n + 1 + 3) & ~3);
    const size_t hit_words = ((((npad >> 2) + threads - 1) / threads) + 7) >> 3;
    size_t bytes = npad * 4 + K * hit_words * threads * 4 + K * kWideListCap * 2 + 16;
    if (grid_smem) bytes += grid_words(W, H) * 4;
    return (bytes + 15) & ~(size_t)15;
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
    const size_t with_grid = wide_smem_bytes(W, H, n, threads, true);
    const size_t without = wide_smem_bytes(W, H, n, threads, false);
    bool in_smem = with_grid <= budget && blocks_by_smem(with_grid, threads, sm_smem) >= blocks_by_smem(without, threads, sm_smem);
    const char *force = getenv("RRTK_GRID_SMEM");
    if (force && force[0] == '0') in_smem = false;
    if (force && force[0] == '1' && with_grid <= budget) in_smem = true;
    *bytes = in_smem ? with_grid : without;
    return in_smem;
}

template <int KIND, int K>
static int launch_kind(const PlanParams &P, int nplans, int W, int H, int n, int optin, int sm_smem, cudaStream_t st)
{
    const int threads = 32 * K;
    size_t bytes = 0;
    const bool in_smem = choose_grid_smem(W, H, n, threads, optin, sm_smem, &bytes);
    if (bytes > (size_t)optin - 2048) {   // static shared (records, slots) is < 1 KB; keep 2 KB headroom
        set_error("plan does not fit shared memory: n=%d needs %zu bytes, device allows %zu", n, bytes, (size_t)optin - 2048);
        return RRTK_ERR_CAPACITY;
    }
    if (in_smem) {
        RRTK_CUDA(cudaFuncSetAttribute(plan_wide_kernel<KIND, true, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        plan_wide_kernel<KIND, true, K><<<nplans, threads, bytes, st>>>(P);
    } else {
        RRTK_CUDA(cudaFuncSetAttribute(plan_wide_kernel<KIND, false, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        plan_wide_kernel<KIND, false, K><<<nplans, threads, bytes, st>>>(P);
    }
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

template <int KIND>
static int launch_threads(const PlanParams &P, int nplans, int W, int H, int n, int threads, int optin, int sm_smem, cudaStream_t st)
{
    switch (threads) {
        case 64: return launch_kind<KIND, 2>(P, nplans, W, H, n, optin, sm_smem, st);
        case 128: return launch_kind<KIND, 4>(P, nplans, W, H, n, optin, sm_smem, st);
        case 256: return launch_kind<KIND, 8>(P, nplans, W, H, n, optin, sm_smem, st);
    }
    set_error("threads must be 64, 128 or 256 (samples per round = threads / 32)");
    return RRTK_ERR_INVALID;
}

int wide_default_threads(int n) { return n >= 512 ? 128 : 64; }

int wide_plan_footprint(int kind, int W, int H, int n, int threads, int optin, int sm_smem, int *smem_bytes, int *blocks_per_sm)
{
    (void)kind;
    if (threads <= 0) threads = wide_default_threads(n);
    size_t b = 0;
    choose_grid_smem(W, H, n, threads, optin, sm_smem, &b);
    if (b > (size_t)optin - 2048) return RRTK_ERR_CAPACITY;
    if (smem_bytes) *smem_bytes = (int)b;
    if (blocks_per_sm) *blocks_per_sm = blocks_by_smem(b, threads, sm_smem);
    return RRTK_OK;
}

int wide_plan_launch(int kind, const uint32_t *d_bits, int W, int H, const rrtk_plan_desc *d_plans, int nplans, int n,
                double r_rewire, double r_goal, const int16_t *d_samples, const double *d_balls, int16_t *d_pts,
                double *d_cost, int32_t *d_parent, int64_t *d_stats, double *d_ell_c, int threads, int optin,
                int sm_smem, cudaStream_t st)
{
    if (threads <= 0) threads = wide_default_threads(n);
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
        case RRTK_STANDARD: return launch_threads<RRTK_STANDARD>(P, nplans, W, H, n, threads, optin, sm_smem, st);
        case RRTK_STAR: return launch_threads<RRTK_STAR>(P, nplans, W, H, n, threads, optin, sm_smem, st);
        case RRTK_INFORMED: return launch_threads<RRTK_INFORMED>(P, nplans, W, H, n, threads, optin, sm_smem, st);
    }
    set_error("unknown planner kind %d", kind);
    return RRTK_ERR_INVALID;
}

}  // namespace rrtk
