// K7, bucket form (RRTStandard / RRTStar): one thread block per plan, K samples per round, NO brute-force scan.
//
// Same loop, same results as plan_scan.cuh (rrt.py:418-437, 498-548, go2goal :284-332; near :131-155, within :157-181,
// collisionfree :183-229, default costfn :70-78) -- tests/test_gpu_parity.py runs every golden tree through both.  What
// changes is how `near` and `within` are answered.  Every vertex the tree will ever hold is one of the n samples, and those
// are known when the plan starts.  So the block first sorts the SAMPLES (and the start) by spatial bucket (2^bshift cells in
// x, 2^bshy in y, x-major: 32 x 4 cells for cfg3) and gives each a fixed slot of the on-chip tree in that order; a slot is
// EMPTY until its sample is accepted, then holds  id << (xb + yb) | y << xb | x.  `near` / `within` for a sample then read
// only the slots of the buckets that hold every cell within rr >= r_rewire of it on both axes -- whole bucket columns in x (one
// run of consecutive slots each, ~4 for cfg3), the exact range of the fine buckets in y: ~260 slots in 11 groups of 32 instead
// of the whole tree -- straight from shared memory, by the warp that owns the sample:
//   * every vertex closer than rr on both axes lies in those buckets, so the radius set is complete, and the nearest vertex
//     found there is THE nearest vertex whenever it is closer than rr (lowest index among equals: ids are in the entries).
//     If it is not (a sparse tree; 13 samples per plan on cfg3), the warp scans all slots for the nearest vertex; while the
//     tree has at most kFirst vertices it scans a by-index copy of them instead of the buckets.
//   * members go to the warp's list as they are found (one ballot per 32 slots); no membership words, no compaction, no
//     block-wide scan phase and no barrier for it: a round is owner phase, barrier, commit phase, barrier.
//   * KEY32: when squared distances and ids fit one word together (cfg3: 19 + 13 bits) a lane's nearest vertex is one
//     integer minimum per slot, and the nearest vertex's point is read back from the `pts` output (written at insertion).
// The owner phase from the costing on and the whole commit phase are those of plan_scan.cuh (see its header for why a round
// of K samples replays to exactly the sequential result).  Against that kernel on cfg3 (profiles/r2_v5_plan_ncu.txt): 6.4 M
// instead of 8.3 M warp instructions per plan (no 12.5 M vertex-sample pairs), two barriers per round instead of three, K = 16
// samples per round pay (with the scan they did not), 113 k instead of 89.6 k plans/s.  The kernel is bound by issue slots
// (28 warps per SM, one instruction per warp every ~10 cycles whatever the dependences) and register-limited: what helped
// was removing instructions and live values, not overlapping chains.
//
// Slot of sample i: found with shared-memory atomics in the prologue (the order inside a bucket is arbitrary and immaterial:
// every tie is broken on the vertex id), kept until the sample's turn in row i + 1 of the plan's `parent` output, which no
// vertex can occupy before iteration i has been committed (the tree has at most i + 1 vertices then).
//
// Limits (plan.cu falls back to plan_scan.cuh): bits(W - 1) + bits(H - 1) + bits(n + 1) <= 32 so that an entry fits a word
// (cfg3: 9 + 9 + 13), at most 2048 buckets and fewer than n (coarser ones otherwise), n >= 256 by default, RRTStandard / RRTStar only
// (an informed plan draws its samples from the tree's own state, so they are not known in advance).
#pragma once
#include "plan_common.cuh"

namespace rrtk {

constexpr uint32_t kSlotEmpty = 0xffffffffu;
constexpr int kFirst = 256;              // vertices also kept by index: the scan of a small tree

#ifndef RRTK_GRID_MINB
#define RRTK_GRID_MINB 7
#endif

// KEY32: squared distances and vertex ids of the plan fit one 32-bit word together (cfg3: 19 + 13 bits), so a lane's nearest
// vertex is one integer minimum per slot; otherwise distance and id are compared separately.
template <int KIND, int K, int T, bool KEY32>
__global__ void __launch_bounds__(T, RRTK_GRID_MINB) plan_grid_kernel(PlanParams P)
{
    constexpr int NW = T / 32;
    static_assert(T % 32 == 0 && K <= 16 && K >= 1, "block shape");
    static_assert(KIND == RRTK_STANDARD || KIND == RRTK_STAR, "the samples of an informed plan are not known in advance");
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ SampleRec s_rec[K];
    __shared__ RoundSummary s_sum;
    __shared__ short2 s_q[K];                         // samples of the round
    __shared__ int s_sig[K];                          // their slots
    __shared__ int s_next;                            // owner phase: next sample nobody has taken yet
    __shared__ int s_rootslot;
    __shared__ unsigned long long s_goalc;
    __shared__ int s_goalv;
    __shared__ unsigned long long s_checks, s_cells;
    __shared__ unsigned long long s_cnt[4];           // commit-phase counters of all warps, summed at the end
    __shared__ uint32_t s_first[kFirst];              // entries of vertices 0 .. kFirst-1 by index

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int plan = blockIdx.x;
    const int n = P.n;
    const int cap = P.list_cap;
    const int nslots = n + 1;
    const int xb = P.g_xb, xyb = P.g_xb + P.g_yb;
    const uint32_t xmask = (1u << P.g_xb) - 1u, ymask = (1u << P.g_yb) - 1u;
    const int bsx = P.g_bshift, bsy = P.g_bshy, NBY = P.g_nby, NB = P.g_nbx * P.g_nby, rr = P.g_rad;
    const uint32_t near_ok2 = P.g_near_ok2;

    uint32_t *s_ent = smem;                                               // nslots entries in bucket order (padded to 32)
    const int ent_words = P.g_ent_words;
    // offsets from the kernel parameters: constant-bank operands instead of address arithmetic the compiler redoes at every use
    uint32_t *s_list = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(smem) + P.g_off_list);      // [NW][cap] entries: one radius-set list per owner warp
    uint16_t *s_bstart = reinterpret_cast<uint16_t *>(reinterpret_cast<unsigned char *>(smem) + P.g_off_bstart);  // NB + 1 first slots
    const int kb = P.g_kb;                                                // KEY32: id bits of a key
    short2 *opts = P.pts + (size_t)plan * (n + 1);

    const rrtk_plan_desc *dsc = P.plans + plan;
    const uint32_t *gbits = P.bits + (size_t)dsc->world * P.words_per_grid;
    const int sx = dsc->start_x, sy = dsc->start_y, gx = dsc->goal_x, gy = dsc->goal_y;

    double *cost = P.cost + (size_t)plan * (n + 1);
    int *parent = P.parent + (size_t)plan * (n + 1);
    const short2 *samples = P.samples + (size_t)plan * n;
    const uint32_t r2x = P.r2_excl;

    auto bucket_of = [&](int x, int y) { return (x >> bsx) * NBY + (y >> bsy); };
    auto make_entry = [&](int x, int y, int id) { return ((uint32_t)id << xyb) | ((uint32_t)y << xb) | (uint32_t)x; };
    auto ent_xy = [&](uint32_t e) { return pack_xy((int)(e & xmask), (int)((e >> xb) & ymask)); };
    auto ent_id = [&](uint32_t e) { return (int)(e >> xyb); };

    // ---- prologue: slots of the samples and of the start, in bucket order -----------------------------------
    {
        uint32_t *cnt = s_ent;                                            // NB + 1 <= nslots counters for now
        for (int b = tid; b <= NB; b += T) cnt[b] = 0u;
        __syncthreads();
        for (int i = tid; i <= n; i += T) {
            const short2 q = i < n ? samples[i] : make_short2((short)sx, (short)sy);
            atomicAdd(&cnt[bucket_of(q.x, q.y)], 1u);
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t run = 0;
            for (int base = 0; base <= NB; base += 32) {
                const int b = base + lane;
                const uint32_t c = b < NB ? cnt[b] : 0u;
                uint32_t incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(RRTK_FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                if (b <= NB) { s_bstart[b] = (uint16_t)(run + incl - c); cnt[b] = run + incl - c; }
                run += __shfl_sync(RRTK_FULL, incl, 31);
            }
        }
        __syncthreads();
        for (int i = tid; i <= n; i += T) {
            const short2 q = i < n ? samples[i] : make_short2((short)sx, (short)sy);
            const int slot = (int)atomicAdd(&cnt[bucket_of(q.x, q.y)], 1u);
            if (i < n) parent[i + 1] = slot;                               // parked until iteration i (see the header)
            else s_rootslot = slot;
        }
        __syncthreads();
        for (int i = tid; i < ent_words; i += T) s_ent[i] = kSlotEmpty;
        for (int i = tid; i < kFirst; i += T) s_first[i] = kSlotEmpty;
        __syncthreads();
    }
    if (tid < K) {
        const int i0 = min(tid, n - 1);
        s_q[tid] = samples[i0];
        s_sig[tid] = parent[i0 + 1];
    }
    if (tid == 0) {
        const uint32_t e0 = make_entry(sx, sy, 0);
        s_ent[s_rootslot] = e0; s_first[0] = e0;
        opts[0] = make_short2((short)sx, (short)sy);
        s_next = NW;
        s_checks = s_cells = 0ull; cost[0] = 0.0; parent[0] = -1;
    }
    if (tid < 4) s_cnt[tid] = 0ull;
    __syncthreads();

    GlobalGrid gg{gbits};
    const int TY = P.TY;
#define WALK(ax_, ay_, bx_, by_) warp_first_hit(gg, TY, ax_, ay_, bx_, by_, lane)

    // block-uniform state, replicated in every thread (refreshed from s_sum after each round)
    int j = 1, it0 = 0, kwant = 1;
    int cw = 0;                          // the commit duty rotates over the warps (plan_scan.cuh)
    // the statistics the commit phase adds up live in shared memory (s_cnt, one writer per round): the kernel is at its register
    // limit, and three 64-bit counters carried through the whole loop cost more than three shared-memory updates per round
    unsigned my_checks = 0, my_cells = 0;
    const unsigned ltmask = (1u << lane) - 1u;
    // -DRRTK_PHASE_CLOCKS (experiment builds, scripts/phase_clocks.py): cycles per phase in spare stats slots
#ifdef RRTK_PHASE_CLOCKS
    long long clk_scan = 0, clk_owner = 0, clk_commit = 0, clk_ownwork = 0, rounds = 0;
#define GPHASE_T(var) const long long var = clock64()
#define GPHASE_ADD(acc, a, b) acc += (b) - (a)
#else
#define GPHASE_T(var)
#define GPHASE_ADD(acc, a, b)
#endif

    while (it0 < n) {
        GPHASE_T(t_0);
        if (j == n) break;                                                    // tree full: every later sample is rejected
        const int kact = min(min(K, n - it0), kwant);                         // adaptive samples per round (plan_scan.cuh)
        short2 ahead = make_short2(0, 0);
        int ahead_sig = 0;
        if (warp == cw && lane < 2 * K) {
            const int ia = min(it0 + lane, n - 1);
            ahead = samples[ia];
            ahead_sig = parent[ia + 1];                                       // still the parked slot: row ia + 1 >= j + ... is not a vertex yet
        }

        // ---- owner phase: the warps take the samples one at a time and evaluate each against the round-start tree ---------
        for (int k = warp; k < K;) {
            if (k >= kact) {
                if (lane == 0) s_rec[k].flags = 0;
                int nk = 0;
                if (lane == 0) nk = atomicAdd(&s_next, 1);
                k = __shfl_sync(RRTK_FULL, nk, 0);
                continue;
            }
            const short2 smp = s_q[k];
            const int x = smp.x, y = smp.y;
            uint32_t *list = s_list + warp * cap;
            GPHASE_T(t_s0);
            // ---- near + within on the buckets around the sample ----
            uint32_t bd = 0xffffffffu, bv = 0xffffffffu, be = 0u;            // this lane's nearest: distance (KEY32: key), id, entry
            int total = 0;
            // 32 slots per step.  (Four steps unrolled into independent chains ran no faster: with 28 warps per SM the kernel is
            // bound by issue slots, not by the chains, so what counts is not to touch a group of slots that the run does not have.)
            constexpr int kVis = 1;
            auto visit = [&](const uint32_t *src, int base, int end, bool members) {
                const int idx = base + lane;
                const uint32_t e = idx < end ? src[idx] : kSlotEmpty;
                const int dx = (int)(e & xmask) - x, dy = (int)((e >> xb) & ymask) - y;
                const bool valid = e != kSlotEmpty;
                const uint32_t d2r = (uint32_t)(dx * dx + dy * dy);
                const uint32_t id = e >> xyb;
                bool mem;
                if (KEY32) {
                    const uint32_t key = valid ? (d2r << kb) + id : 0xffffffffu;
                    bd = min(bd, key);
                    mem = valid && d2r < r2x;
                } else {
                    const uint32_t d2 = valid ? d2r : 0xffffffffu;
                    if (d2 < bd || (d2 == bd && id < bv)) { bd = d2; bv = id; be = e; }
                    mem = d2 < r2x;
                }
                if (KIND != RRTK_STANDARD && members) {
                    const unsigned mm = __ballot_sync(RRTK_FULL, mem);
                    const int at = total + __popc(mm & ltmask);
                    if (mem && at < cap) list[at] = e;
                    total += __popc(mm);
                }
            };
            auto nearest_d2 = [&]() { const uint32_t m = __reduce_min_sync(RRTK_FULL, bd); return KEY32 ? (m >> kb) : m; };
            __syncwarp();                                                  // the lanes are done reading the previous sample's list
            if (j <= kFirst) {
                for (int base = 0; base < j; base += 32 * kVis) visit(s_first, base, j, true);
            } else {
                // every vertex within rr cells of the sample on both axes: whole buckets in x (one run of slots each), the exact
                // range of the (much finer) buckets in y
                const int bx0 = max(x - rr, 0) >> bsx, bx1 = min(x + rr, P.W - 1) >> bsx;
                const int by0 = max(y - rr, 0) >> bsy, by1 = min(y + rr, P.H - 1) >> bsy;
                for (int bxx = bx0; bxx <= bx1; ++bxx) {
                    const int s0 = s_bstart[bxx * NBY + by0], s1 = s_bstart[bxx * NBY + by1 + 1];
                    for (int base = s0; base < s1; base += 32 * kVis) visit(s_ent, base, s1, true);
                }
                if (nearest_d2() > near_ok2) {
                    // nothing that close in these buckets: the nearest vertex may lie anywhere (the radius set is complete as it is)
                    for (int base = 0; base < nslots; base += 32 * kVis) visit(s_ent, base, nslots, false);
                }
            }
            __syncwarp();                                                  // list entries written by other lanes are read below
            GPHASE_T(t_s1);
            GPHASE_ADD(clk_scan, t_s0, t_s1);
            int vnear;
            uint32_t pnear;
            if (KEY32) {
                const uint32_t mk = __reduce_min_sync(RRTK_FULL, bd);
                bd = mk >> kb;
                vnear = (int)(mk & ((1u << kb) - 1u));
                const short2 pn = opts[vnear];                                // written when the vertex was inserted
                pnear = pack_xy(pn.x, pn.y);
            } else {
                const uint32_t md = __reduce_min_sync(RRTK_FULL, bd);
                const uint32_t mv = __reduce_min_sync(RRTK_FULL, bd == md ? bv : 0xffffffffu);
                const uint32_t near_e = __shfl_sync(RRTK_FULL, be, __ffs(__ballot_sync(RRTK_FULL, bd == md && bv == mv)) - 1);
                bd = md;
                vnear = (int)mv;
                pnear = ent_xy(near_e);
            }
            // `sampled` holds accepted samples only, not xstart (rrt.py:410,426): a sample on the root's cell is a
            // duplicate only if some vertex >= 1 sits there too
            bool dup = bd == 0 && vnear >= 1;
            if (bd == 0 && vnear == 0) {
                bool f = false;
                if (j <= kFirst) {
                    for (int v = 1 + lane; v < j; v += 32) f |= ent_xy(s_first[v]) == pack_xy(x, y);
                } else {
                    const int b = bucket_of(x, y);
                    for (int s = s_bstart[b] + lane; s < s_bstart[b + 1]; s += 32) {
                        const uint32_t e = s_ent[s];
                        f |= e != kSlotEmpty && ent_id(e) >= 1 && ent_xy(e) == pack_xy(x, y);
                    }
                }
                dup = __any_sync(RRTK_FULL, f);
            }
            int flags = 1 | (dup ? 2 : 0);
            double c0 = 0.0, wc = CUDART_INF;
            int wv = 0x7fffffff, ring = 0;
            // the reference walks nearest -> sample before looking at the duplicate test (rrt.py:424-425 / 506-507); the
            // verdicts are independent, so skip the walk
            if (!dup) {
                const double cnear = cost[vnear];
                const int hit = WALK(px(pnear), py(pnear), x, y);
                my_checks += 1; my_cells += cells_tested(hit);
                if (hit < 0) {
                    flags |= 4;
                    c0 = reach_cost(cnear, bd);
                    if (KIND != RRTK_STANDARD) {
                        ring = total;
                        if (total <= cap) {
                            // choose-parent (rrt.py:510-521), cheapest first: up to three candidates per lane have their costs
                            // loaded and evaluated together; the cheapest live one of the warp is walked, and the first
                            // free one is the minimum over (cost, index) of everything that beats the nearest vertex.
                            // (the list is in slot order, not index order: every tie is resolved on the vertex index)
                            for (int base = 0; base < total; base += 96) {
                                int cv[3];
                                uint32_t cp[3];
                                unsigned long long ck[3];                     // cost bits; ~0 = not a candidate
                                double ccv[3];
                                const int left = total - base;                // slots u with 32 u >= left are empty for every lane
#pragma unroll
                                for (int u = 0; u < 3; ++u) {
                                    ck[u] = ~0ull; cv[u] = 0; cp[u] = 0u; ccv[u] = CUDART_INF;
                                    if (32 * u < left) {
                                        const int idx = base + 32 * u + lane;
                                        const bool has = idx < total;
                                        const uint32_t e = has ? list[idx] : 0u;
                                        cv[u] = ent_id(e);
                                        cp[u] = ent_xy(e);
                                        if (has) ccv[u] = cost[cv[u]];
                                    }
                                }
#pragma unroll
                                for (int u = 0; u < 3; ++u) {
                                    if (32 * u < left) {
                                        const double cn = reach_cost(ccv[u], dist2(cp[u], x, y));
                                        const bool live = cn < c0 && (cn < wc || (cn == wc && cv[u] < wv));
                                        if (live) ck[u] = (unsigned long long)__double_as_longlong(cn);   // positive doubles order like their bits
                                    }
                                }
                                for (;;) {
                                    // this lane's cheapest live candidate, lowest vertex index among equal costs
                                    unsigned long long bk = ck[0];
                                    int bu = 0, bvx = cv[0];
                                    if (ck[1] < bk || (ck[1] == bk && cv[1] < bvx)) { bk = ck[1]; bu = 1; bvx = cv[1]; }
                                    if (ck[2] < bk || (ck[2] == bk && cv[2] < bvx)) { bk = ck[2]; bu = 2; bvx = cv[2]; }
                                    const uint32_t hi = (uint32_t)(bk >> 32);
                                    const uint32_t mhi = warp_min_u32(hi);
                                    if (mhi == 0xffffffffu) break;                    // nothing left that beats the incumbent
                                    const uint32_t lo = (hi == mhi) ? (uint32_t)bk : 0xffffffffu;
                                    const uint32_t mlo = warp_min_u32(lo);
                                    const uint32_t vv = (hi == mhi && lo == mlo) ? (uint32_t)bvx : 0xffffffffu;
                                    const uint32_t mvv = warp_min_u32(vv);
                                    const int src = __ffs(__ballot_sync(RRTK_FULL, vv == mvv)) - 1;
                                    const uint32_t bpx = bu == 0 ? cp[0] : bu == 1 ? cp[1] : cp[2];
                                    const uint32_t pp = __shfl_sync(RRTK_FULL, bpx, src);
                                    const int h = WALK(px(pp), py(pp), x, y);
                                    my_checks += 1; my_cells += cells_tested(h);
                                    if (h < 0) { wc = __hiloint2double((int)mhi, (int)mlo); wv = (int)mvv; break; }
                                    if (lane == src) {
                                        if (bu == 0) ck[0] = ~0ull; else if (bu == 1) ck[1] = ~0ull; else ck[2] = ~0ull;
                                    }
                                }
                            }
                        } else {
                            // very large radius sets: test every slot directly, 32 per step; while one beats the incumbent, walk the cheapest
                            auto consider = [&](bool has, int v, uint32_t p, double cn) {
                                bool live = has && cn < c0;
                                for (;;) {
                                    const bool cand = live && (cn < wc || (cn == wc && v < wv));
                                    if (!__any_sync(RRTK_FULL, cand)) break;
                                    const uint32_t hi = cand ? (uint32_t)__double2hiint(cn) : 0xffffffffu;
                                    const uint32_t mhi = warp_min_u32(hi);
                                    const uint32_t lo = (cand && hi == mhi) ? (uint32_t)__double2loint(cn) : 0xffffffffu;
                                    const uint32_t mlo = warp_min_u32(lo);
                                    const uint32_t vv = (cand && hi == mhi && lo == mlo) ? (uint32_t)v : 0xffffffffu;
                                    const uint32_t mvv = warp_min_u32(vv);
                                    const int src = __ffs(__ballot_sync(RRTK_FULL, vv == mvv && mvv != 0xffffffffu)) - 1;
                                    const uint32_t pp = __shfl_sync(RRTK_FULL, p, src);
                                    const int h = WALK(px(pp), py(pp), x, y);
                                    my_checks += 1; my_cells += cells_tested(h);
                                    if (h < 0) { wc = __hiloint2double((int)mhi, (int)mlo); wv = (int)mvv; }
                                    else if (lane == src) live = false;
                                }
                            };
                            for (int base = 0; base < nslots; base += 32) {
                                const uint32_t e = base + lane < nslots ? s_ent[base + lane] : kSlotEmpty;
                                const bool valid = e != kSlotEmpty;
                                const uint32_t p = ent_xy(e);
                                const uint32_t dd = dist2(p, x, y);
                                const int v = ent_id(e);
                                const bool has = valid && dd < r2x;
                                double cn = CUDART_INF;
                                if (has) cn = reach_cost(cost[v], dd);
                                consider(has, v, p, cn);
                            }
                        }
                    }
                }
            }
            if (lane == 0) {
                SampleRec r;
                r.pnew = pack_xy(x, y); r.bd = bd; r.vnear = vnear; r.flags = flags; r.bv = wv; r.ring = ring;
                r.c0 = c0; r.bc = wc; r.ell = 0.0;
                s_rec[k] = r;
            }
            {
                int nk = 0;
                if (lane == 0) nk = atomicAdd(&s_next, 1);
                k = __shfl_sync(RRTK_FULL, nk, 0);
            }
        }
        GPHASE_T(t_1b);
        __syncthreads();                                                   // ---- barrier: K results visible
        GPHASE_T(t_2);

        // ---- commit phase (plan_scan.cuh): one warp replays the results in sample order; lane m holds the m-th vertex
        //      accepted in this round ----------------------------------------------------------------
        if (warp == cw) {
            uint32_t newp = 0;
            double newc = 0.0;
            int nnew = 0, consumed = 0, jc = j;
            bool finished = false;
            bool fast = false;
            auto insert = [&](int v, uint32_t pnew, int slot) {             // the new vertex becomes visible to the next round's scans
                const uint32_t e = make_entry(px(pnew), py(pnew), v);
                s_ent[slot] = e;
                if (v < kFirst) s_first[v] = e;
                if (KEY32) opts[v] = make_short2((short)px(pnew), (short)py(pnew));
            };
            if (j + kact <= n) {
                const SampleRec r = s_rec[min(lane, K - 1)];
                const int myslot = s_sig[min(lane, K - 1)];
                const bool in = lane < kact;
                bool a = in && !(r.flags & 2) && (r.flags & 4);              // accepted unless the round interferes
                double bc = r.bc;
                int bv = r.bv;
                const int x = px(r.pnew), y = py(r.pnew);
                int extra = 0, stop = kact, nacc = 0;                        // samples [stop, kact) are left for the next round
                unsigned addm = __ballot_sync(RRTK_FULL, a);                 // who adds a vertex, as far as settled (bit kk is final at step kk)
                for (int kk = 0; kk < stop; ++kk) {
                    if (!((addm >> kk) & 1u)) continue;                      // sample kk adds no vertex
                    const uint32_t pk = __shfl_sync(RRTK_FULL, r.pnew, kk);
                    const uint32_t du = dist2(pk, x, y);
                    const bool later = in && lane > kk && lane < stop;
                    // the common step: nobody later in the round is touched by this vertex (one ballot, nothing else to exchange)
                    if (!__any_sync(RRTK_FULL, later && (du < r.bd || (a && (du == 0 || du < r2x))))) { ++nacc; continue; }
                    const double myc = (bv != 0x7fffffff) ? bc : r.c0;       // final for lane kk
                    const double ck = __shfl_sync(RRTK_FULL, myc, kk);
                    const unsigned cutm = __ballot_sync(RRTK_FULL, later && du != 0 && du < r.bd);
                    if (cutm) stop = min(stop, __ffs(cutm) - 1);             // it would be their nearest vertex: redo from there
                    const bool still = later && lane < stop;
                    if (still && du == 0 && a) a = false;                     // now in `sampled` (rrt.py:426/508)
                    addm = __ballot_sync(RRTK_FULL, a);
                    const bool inr = KIND != RRTK_STANDARD && still && a && du < r2x;
                    if (__any_sync(RRTK_FULL, inr)) {
                        double cn = CUDART_INF;
                        if (inr) { ++extra; cn = reach_cost(ck, du); }
                        unsigned wm = __ballot_sync(RRTK_FULL, inr && cn < r.c0 && cn < bc);    // higher index: loses cost ties
                        while (wm) {
                            const int dst = __ffs(wm) - 1;
                            wm &= wm - 1;
                            const uint32_t pq = __shfl_sync(RRTK_FULL, r.pnew, dst);
                            const int h = WALK(px(pk), py(pk), px(pq), py(pq));
                            my_checks += 1; my_cells += cells_tested(h);
                            if (h < 0 && lane == dst) { bc = cn; bv = j + nacc; }
                        }
                    }
                    ++nacc;
                }
                const bool cons = lane < stop;
                const bool acc = a && cons;
                const unsigned accm = __ballot_sync(RRTK_FULL, acc);
                const int myj = j + __popc(accm & ltmask);                    // the tree size when this sample is reached
                if (acc) {                                                    // rrt.py:524-529
                    const bool hasbv = bv != 0x7fffffff;
                    insert(myj, r.pnew, myslot);
                    cost[myj] = hasbv ? bc : r.c0; parent[myj] = hasbv ? bv : r.vnear;
                }
                consumed = stop;
                jc = j + __popc(accm);
                const int np_ = __reduce_add_sync(RRTK_FULL, cons ? myj : 0), rm_ = __reduce_add_sync(RRTK_FULL, acc ? r.ring + extra : 0);
                if (lane == 0) { s_cnt[1] += (unsigned long long)np_; s_cnt[2] += (unsigned long long)rm_; s_cnt[3] += (unsigned long long)__popc(accm); }
                fast = true;
            }
            for (int k = 0; k < kact && !fast; ++k) {                         // rounds in which the tree may fill up: one sample at a time
                const SampleRec r = s_rec[k];
                if (jc == n) { finished = true; break; }
                const int x = px(r.pnew), y = py(r.pnew);
                bool reject = (r.flags & 2) || jc == n || !(r.flags & 4);
                double bc = r.bc;
                int bv = r.bv;
                const bool mine = lane < nnew;
                const uint32_t du = mine ? dist2(newp, x, y) : 0xffffffffu;
                if (__any_sync(RRTK_FULL, mine && du != 0 && du < r.bd)) break;           // it would be the nearest vertex: redo
                if (__any_sync(RRTK_FULL, mine && du == 0)) reject = true;             // now in `sampled` (rrt.py:426/508)
                int ringm = r.ring;
                if (KIND != RRTK_STANDARD && !reject) {
                    const bool inr = mine && du < r2x;
                    const unsigned inm = __ballot_sync(RRTK_FULL, inr);
                    ringm += __popc(inm);
                    if (inm) {
                        double cn = CUDART_INF;
                        if (inr) cn = reach_cost(newc, du);
                        bool live = inr && cn < r.c0 && cn < bc;              // higher index: loses cost ties
                        for (;;) {                                            // cheapest first; equal cost -> lower lane
                            if (!__any_sync(RRTK_FULL, live)) break;
                            const uint32_t hi = live ? (uint32_t)__double2hiint(cn) : 0xffffffffu;
                            const uint32_t mhi = warp_min_u32(hi);
                            const uint32_t lo = (live && hi == mhi) ? (uint32_t)__double2loint(cn) : 0xffffffffu;
                            const uint32_t mlo = warp_min_u32(lo);
                            const int src = __ffs(__ballot_sync(RRTK_FULL, live && hi == mhi && lo == mlo)) - 1;
                            const uint32_t pp = __shfl_sync(RRTK_FULL, newp, src);
                            const int h = WALK(px(pp), py(pp), x, y);
                            my_checks += 1; my_cells += cells_tested(h);
                            if (h < 0) { bc = __hiloint2double((int)mhi, (int)mlo); bv = j + src; break; }
                            if (lane == src) live = false;
                        }
                    }
                }
                ++consumed;
                if (lane == 0) s_cnt[1] += (unsigned long long)jc;
                if (reject) continue;
                if (lane == 0) { s_cnt[2] += (unsigned long long)ringm; s_cnt[3] += 1ull; }
                const int vbest = (bv != 0x7fffffff) ? bv : r.vnear;
                const double cbest = (bv != 0x7fffffff) ? bc : r.c0;
                if (lane == 0) { insert(jc, r.pnew, s_sig[k]); cost[jc] = cbest; parent[jc] = vbest; }   // rrt.py:524-529
                if (lane == nnew) { newp = r.pnew; newc = cbest; }
                ++nnew;
                ++jc;
            }
            if (it0 + consumed >= n) finished = true;
            // stage the next round's samples and their slots
            __syncwarp();                                                      // lane 0's tree writes -> all lanes
            const int kwant_next = consumed < kact ? max(min(K, 1 + (jc >> 3)), max(1, kact >> 1)) : min(K, 2 * kact);
            if (lane == 0) s_next = NW;
            {
                const int src = min(consumed + lane, 31);
                const int ax_ = __shfl_sync(RRTK_FULL, (int)ahead.x, src), ay_ = __shfl_sync(RRTK_FULL, (int)ahead.y, src);
                const int sg_ = __shfl_sync(RRTK_FULL, ahead_sig, src);
                if (lane < K) { s_q[lane] = make_short2((short)ax_, (short)ay_); s_sig[lane] = sg_; }
            }
            if (lane == 0) {
                RoundSummary s;
                s.j = jc; s.consumed = consumed; s.flags = finished ? 2 : 0;
                s.vsol = 0; s.csol = 0.0; s.first_sol = -1;
                s.kwant = kwant_next; s.pad = 0;
                s_sum = s;
            }
        }
        __syncthreads();                                                   // ---- barrier: tree updated
        GPHASE_T(t_3);
        GPHASE_ADD(clk_owner, t_0, t_2); GPHASE_ADD(clk_commit, t_2, t_3); GPHASE_ADD(clk_ownwork, t_0, t_1b);
#ifdef RRTK_PHASE_CLOCKS
        ++rounds;
#endif
        {
            const RoundSummary s = s_sum;
            cw = (cw + 1 == NW) ? 0 : cw + 1;
            j = s.j;
            it0 += s.consumed;
            kwant = s.kwant;
            if (s.flags & 2) break;
        }
    }

    // ---- goal connection: rrt.py:284-332, ascending (cost, index), filled vertices only (here: in slot order; the result is
    //      the minimum over (cost, index), whatever the order) ------
    if (tid == 0) { s_goalc = 0x7ff0000000000000ull; s_goalv = 0x7fffffff; }
    __syncthreads();
    for (int base = warp * 32; base < nslots; base += NW * 32) {
        const uint32_t e = base + lane < nslots ? s_ent[base + lane] : kSlotEmpty;
        const bool valid = e != kSlotEmpty;
        const uint32_t p = ent_xy(e);
        double cg = CUDART_INF;
        if (valid) cg = reach_cost(cost[ent_id(e)], dist2(p, gx, gy));
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
        for (int base = warp * 32; base < nslots; base += NW * 32) {
            const uint32_t e = base + lane < nslots ? s_ent[base + lane] : kSlotEmpty;
            const bool valid = e != kSlotEmpty;
            const uint32_t p = ent_xy(e);
            const int v = ent_id(e);
            double cg = CUDART_INF;
            if (valid) cg = reach_cost(cost[v], dist2(p, gx, gy));
            unsigned m = __ballot_sync(RRTK_FULL, valid && (unsigned long long)__double_as_longlong(cg) == cstar_bits);
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t pp = __shfl_sync(RRTK_FULL, p, l);
                const int vv = __shfl_sync(RRTK_FULL, v, l);
                const int h = WALK(px(pp), py(pp), gx, gy);
                my_checks += 1; my_cells += cells_tested(h);
                if (h < 0 && lane == 0) atomicMin(&s_goalv, vv);
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
    for (int v = j + tid; v <= n; v += T) {                                    // rows without a vertex (the parked slots among them)
        opts[v] = (v == j && found) ? make_short2((short)gx, (short)gy) : make_short2(-32768, -32768);
        if (v >= top) { cost[v] = CUDART_INF; parent[v] = -1; }
    }
    if (!KEY32)
        for (int s = tid; s < nslots; s += T) {
            const uint32_t e = s_ent[s];
            if (e != kSlotEmpty) opts[ent_id(e)] = make_short2((short)(e & xmask), (short)((e >> xb) & ymask));
        }
    __syncthreads();
    if (tid == 0) {
        if (found) { cost[j] = __longlong_as_double((long long)cstar_bits); parent[j] = vparent; }
        long long *st = P.stats + (size_t)plan * RRTK_STAT_COUNT;
        st[RRTK_STAT_J] = j;
        st[RRTK_STAT_VGOAL] = found ? j : 0;
        st[RRTK_STAT_FOUND] = found ? 1 : 0;
        st[RRTK_STAT_CHECKS] = (long long)s_checks;
        st[RRTK_STAT_CELLS] = (long long)s_cells;
        st[RRTK_STAT_FIRST_SOL_ITER] = -1;
        st[RRTK_STAT_ELL_ITERS] = 0;
        st[RRTK_STAT_NN_PAIRS] = (long long)s_cnt[1];
        st[RRTK_STAT_RING_MEMBERS] = (long long)s_cnt[2];
        st[RRTK_STAT_ACCEPTED] = (long long)s_cnt[3];
        st[RRTK_STAT_RESERVED0] = 0;
        st[RRTK_STAT_RESERVED1] = 0;
#ifdef RRTK_PHASE_CLOCKS                      // thread 0's view: scan = its own neighbourhood scans, owner = round start .. first barrier
        st[RRTK_STAT_RESERVED0] = clk_scan;
        st[RRTK_STAT_RESERVED1] = clk_owner;
        st[RRTK_STAT_ELL_ITERS] = clk_commit;
        st[RRTK_STAT_FIRST_SOL_ITER] = clk_ownwork;
        st[RRTK_STAT_RING_MEMBERS] = rounds;
#endif
    }
#undef WALK
}

}  // namespace rrtk
