// K7 dispatch: which plan kernel runs a batch, with what block shape and shared-memory budget.
//
// 1. RRTStandard / RRTStar with n >= 256 whose tree entries fit one word (bits(W-1) + bits(H-1) + bits(n+1) <= 32, cfg3:
//    9 + 9 + 13) and whose tree fits shared memory: the bucket kernel (plan_grid.cuh) -- the tree in shared memory in
//    bucket order of the samples, near / within from the buckets around a sample, 16 samples per round, 128 threads.
// 2. Otherwise the packed-key scan kernel (plan_scan.cuh): brute-force scan over the whole tree.  Its keys hold
//    (d2 - |q|^2) * 2^sbits + row in 32 bits, so it takes grids up to 2896 / 2048 / 1448 cells a side (1 / 2 / 3-4
//    membership words per thread, i.e. n up to 32 / 64 / 128 vertices per thread).
// 3. Anything larger -- up to the 16384-cell limit of the packed vertex format -- runs on the 32-bit-distance kernel of
//    plan_wide.cu.
// RRTK_PLAN_IMPL=grid|scan|wide, RRTK_PLAN_K / RRTK_GRID_K=<samples per round> override the choice for experiments and tests
// (grid: forced for any size that fits).
#include <cstdlib>

#include "plan_common.cuh"

namespace rrtk {

int scan_launch_standard(const PlanParams &, int, int, int, int, size_t, cudaStream_t);
int scan_launch_star(const PlanParams &, int, int, int, int, size_t, cudaStream_t);
int scan_launch_informed(const PlanParams &, int, int, int, int, size_t, cudaStream_t);
int scan_occupancy_standard(int, int, int, size_t);
int scan_occupancy_star(int, int, int, size_t);
int scan_occupancy_informed(int, int, int, size_t);
int grid_launch(int, int, const PlanParams &, int, size_t, cudaStream_t);
int grid_occupancy(int, int, bool, size_t);
int wide_plan_launch(int, const uint32_t *, int, int, const rrtk_plan_desc *, int, int, double, double, const int16_t *,
                     const double *, int16_t *, double *, int32_t *, int64_t *, double *, int, int, int, cudaStream_t);
int wide_plan_footprint(int, int, int, int, int, int, int, int *, int *);

struct ScanShape {
    int T, K, hit_words, tail_bytes, sbits, steps_max, list_cap, blocks_per_sm, two_per_sm;
    size_t smem;
};

static int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// registers: the launch bounds of plan_scan_kernel (ScanCfg::kMinBlocks)
#ifndef RRTK_MINB128
#define RRTK_MINB128 7
#endif
static int scan_min_blocks(int T) { return T <= 64 ? 12 : T <= 128 ? RRTK_MINB128 : T <= 160 ? 5 : T <= 256 ? 3 : 1; }

// ---- the bucket form (plan_grid.cuh): RRTStandard / RRTStar whose tree entries fit one word ---------------------------------
struct GridShape {
    int xb, yb, bshift, bshy, nbx, nby, rad, list_cap, blocks_per_sm, K, kb, ent_words, off_list, off_bstart;
    uint32_t near_ok2;
    size_t smem;
};

static int bits_for(int v) { int b = 0; while ((1 << b) <= v) ++b; return b; }   // bits that hold 0 .. v

static bool grid_shape(int kind, int W, int H, int n, double r_rewire, int threads, int optin, int sm_smem, GridShape *out)
{
    const char *impl = getenv("RRTK_PLAN_IMPL");
    if (impl && (impl[0] == 'w' || impl[0] == 's')) return false;
#ifdef RRTK_GRID_OPT_IN
    if (!(impl && impl[0] == 'g')) return false;                               // experiment builds: only with RRTK_PLAN_IMPL=grid
#endif
    if (kind != RRTK_STANDARD && kind != RRTK_STAR) return false;
    if (threads != 0 && threads != 128) return false;
    if (env_int("RRTK_PLAN_T", 0) != 0) return false;
    // measured against the scan form on 512 x 512 worlds, r = 50 (scripts/size_sweep.py): n = 400 / 1000 / 2048 / 5000 / 5500 / 8000 ->
    // 1.09 / 1.19 / 1.22 / 1.27 / 1.82 / 1.71 x; trees of fewer than 256 vertices never leave the by-index path, so they stay
    // with the scan form unless the bucket form is forced (tests)
    if (n < 256 && !(impl && impl[0] == 'g')) return false;
    GridShape g;
    g.K = env_int("RRTK_GRID_K", 16);                                          // measured on cfg3: 100.4 k plans/s against 99.2 k with 8
    if (g.K != 8 && g.K != 16) return false;
    g.xb = bits_for(W - 1); g.yb = bits_for(H - 1);
    const int ib = 32 - g.xb - g.yb;
    if (ib < 1 || ib > 31 || (long long)n >= (1ll << ib) - 1) return false;    // ids 0 .. n, all ones = an empty slot
    // buckets: 32 cells in x (a sample reads ~4 runs of slots), 4 cells in y (a run covers just the y range it needs);
    // coarser while there would be more than 2048 of them (cfg3: 16 x 128)
    g.bshift = env_int("RRTK_GRID_BSX", 5); g.bshy = env_int("RRTK_GRID_BSY", 2);
    for (;;) {
        g.nbx = (W + (1 << g.bshift) - 1) >> g.bshift; g.nby = (H + (1 << g.bshy) - 1) >> g.bshy;
        if (g.nbx * g.nby <= 2048 && g.nbx * g.nby + 1 <= n) break;           // (the prologue's counters live in the entry array)
        if (g.bshift >= 15) return false;
        if (g.bshy < g.bshift) ++g.bshy; else ++g.bshift;
    }
    const double r = kind == RRTK_STAR ? (r_rewire > 0.0 ? r_rewire : 0.0) : 0.0;
    double rad = ceil(r);
    if (rad < (double)(1 << g.bshift)) rad = (double)(1 << g.bshift);          // near: look at least one x bucket around
    if (rad > 32768.0) rad = 32768.0;
    g.rad = (int)rad;
    g.near_ok2 = (uint32_t)(rad * rad);
    int cap = env_int("RRTK_PLAN_CAP", 256);
    const int need = (n + 1 + 31) & ~31;
    g.list_cap = kind == RRTK_STANDARD ? 0 : (cap < need ? cap : need);
    g.ent_words = (n + 1 + 31) & ~31;
    g.off_list = 4 * g.ent_words;
    g.off_bstart = g.off_list + 4 * 4 * g.list_cap;
    g.smem = (size_t)g.off_bstart + (((size_t)2 * (g.nbx * g.nby + 1) + 15) & ~(size_t)15);
    // one-word (distance, id) keys: ids 0 .. n in kb bits, squared distances up to (W-1)^2 + (H-1)^2 in the rest
    g.kb = bits_for(n);
    {
        const long long d2max = (long long)(W - 1) * (W - 1) + (long long)(H - 1) * (H - 1);
        if (g.kb >= 32 || d2max >= (1ll << (32 - g.kb)) - 1 || env_int("RRTK_GRID_KEY32", 1) == 0) g.kb = 0;
    }
    if (g.smem > (size_t)optin - 4096) return false;
    int b = grid_occupancy(kind, g.K, g.kb > 0, g.smem);
    if (b <= 0) {
        b = (int)((size_t)sm_smem / (g.smem + 3072));
        if (b > 7) b = 7;
    }
    g.blocks_per_sm = b < 1 ? 1 : b;
    *out = g;
    return true;
}

static bool scan_shape(int kind, int W, int H, int n, int threads, int optin, int sm_smem, ScanShape *out)
{
    const char *impl = getenv("RRTK_PLAN_IMPL");
    if (impl && impl[0] == 'w') return false;
    if (n + 1 > 65535) return false;                       // uint16 list entries
    ScanShape s;
    s.K = env_int("RRTK_PLAN_K", 8);
    if (s.K != 4 && s.K != 8 && s.K != 16) return false;
    int T = threads > 0 ? threads : env_int("RRTK_PLAN_T", 0);
    if (T <= 0) T = n < 2048 ? 64 : n < 5120 ? 128 : n < 32768 ? 256 : 512;   // measured: scripts/sweep_tk.sh
    if (T != 64 && T != 128 && T != 160 && T != 256 && T != 512) return false;
    const int rows = (n + 1 + T - 1) / T;
    s.T = T;
    s.hit_words = (rows + 31) / 32;
    if (s.hit_words > 4) return false;
    s.sbits = s.hit_words == 1 ? 5 : s.hit_words == 2 ? 6 : 7;
    const long long side = W > H ? W : H;
    if (((2 * side * side) << s.sbits) > (1ll << 29)) return false;     // |key| and |key - threshold| stay below 2^31
    s.steps_max = (rows + 3) / 4;
    // radius-set list per owner warp: larger sets take the exact brute-force path.  448 rather than 512 entries for informed
    // plans lets two blocks of cfg4 (n = 20000: an 80 KB tree) share an SM: 2.24 k instead of 1.68 k plans/s
    int cap = env_int("RRTK_PLAN_CAP", kind == RRTK_INFORMED ? 448 : 256);
    const int need = (n + 1 + 31) & ~31;
    s.list_cap = kind == RRTK_STANDARD ? 0 : (cap < need ? cap : need);
    s.smem = (size_t)4 * 4 * T * s.steps_max;
    const int tail_rows = rows - 32 * (s.hit_words - 1);                 // rows of the last membership word a full tree uses
    s.tail_bytes = tail_rows <= 8 ? 1 : tail_rows <= 16 ? 2 : 4;
    if (kind != RRTK_STANDARD)
        s.smem += (size_t)4 * (s.hit_words - 1) * s.K * T + (((size_t)s.tail_bytes * s.K * T + 3) & ~(size_t)3) + (size_t)2 * (T / 32) * s.list_cap;   // membership words, one list per warp
    s.smem = (s.smem + 15) & ~(size_t)15;
    if (s.smem > (size_t)optin - 2048) return false;
    // 256-thread blocks whose shared memory admits at most two per SM run the build that is bounded for two (128 registers)
    s.two_per_sm = (T == 256 && (size_t)sm_smem / (s.smem + 2048) <= 2) ? 1 : 0;
    // resident blocks per SM: asked of the runtime for the kernel that will run (no device -> the static estimate)
    int b = kind == RRTK_STANDARD ? scan_occupancy_standard(T, s.K, s.two_per_sm, s.smem)
            : kind == RRTK_STAR   ? scan_occupancy_star(T, s.K, s.two_per_sm, s.smem)
                                  : scan_occupancy_informed(T, s.K, s.two_per_sm, s.smem);
    if (b <= 0) {
        const int by_smem = (int)((size_t)sm_smem / (s.smem + 1024 + 1024));   // + static + per-block reservation
        const int by_threads = 2048 / T;
        b = by_smem < by_threads ? by_smem : by_threads;
        const int mb = s.two_per_sm ? 2 : scan_min_blocks(T);
        const int by_regs = 65536 / (T * (65536 / (T * mb) / 8 * 8));
        if (b > by_regs) b = by_regs;
    }
    s.blocks_per_sm = b < 1 ? 1 : (b > 32 ? 32 : b);
    *out = s;
    return true;
}

// which of the three kernels plan_launch would run: "grid", "scan" or "wide"
const char *plan_kernel_name(int kind, int W, int H, int n, int threads, int optin, int sm_smem)
{
    GridShape g;
    if (grid_shape(kind, W, H, n, 0.0, threads, optin, sm_smem, &g)) return "grid";
    ScanShape s;
    return scan_shape(kind, W, H, n, threads, optin, sm_smem, &s) ? "scan" : "wide";
}

int plan_footprint(int kind, int W, int H, int n, int threads, int optin, int sm_smem, int *smem_bytes, int *blocks_per_sm)
{
    GridShape g;
    if (grid_shape(kind, W, H, n, 0.0, threads, optin, sm_smem, &g)) {
        if (smem_bytes) *smem_bytes = (int)g.smem;
        if (blocks_per_sm) *blocks_per_sm = g.blocks_per_sm;
        return RRTK_OK;
    }
    ScanShape s;
    if (!scan_shape(kind, W, H, n, threads, optin, sm_smem, &s))
        return wide_plan_footprint(kind, W, H, n, threads == 160 || threads == 512 ? 0 : threads, optin, sm_smem, smem_bytes, blocks_per_sm);
    if (smem_bytes) *smem_bytes = (int)s.smem;
    if (blocks_per_sm) *blocks_per_sm = s.blocks_per_sm;
    return RRTK_OK;
}

int plan_launch(int kind, const uint32_t *d_bits, int W, int H, const rrtk_plan_desc *d_plans, int nplans, int n,
                double r_rewire, double r_goal, const int16_t *d_samples, const double *d_balls, int16_t *d_pts,
                double *d_cost, int32_t *d_parent, int64_t *d_stats, double *d_ell_c, int threads, int optin,
                int sm_smem, cudaStream_t st)
{
    if (kind != RRTK_STANDARD && kind != RRTK_STAR && kind != RRTK_INFORMED) {
        set_error("unknown planner kind %d", kind);
        return RRTK_ERR_INVALID;
    }
    GridShape g;
    const bool use_grid = grid_shape(kind, W, H, n, r_rewire, threads, optin, sm_smem, &g);
    ScanShape s;
    if (!use_grid && !scan_shape(kind, W, H, n, threads, optin, sm_smem, &s))
        return wide_plan_launch(kind, d_bits, W, H, d_plans, nplans, n, r_rewire, r_goal, d_samples, d_balls, d_pts, d_cost,
                                d_parent, d_stats, d_ell_c, threads == 160 || threads == 512 ? 0 : threads, optin, sm_smem, st);
    PlanParams P;
    P.bits = d_bits;
    P.words_per_grid = grid_words(W, H);
    P.W = W; P.H = H; P.TY = tiles_y(H);
    P.plans = d_plans;
    P.n = n;
    const double lim = ceil(r_rewire * r_rewire);
    P.r2_excl = (kind == RRTK_STANDARD) ? 0u : (lim >= 1073741824.0 ? 1073741824u : (lim <= 0.0 ? 0u : (uint32_t)lim));
    P.r_goal = r_goal;
    P.samples = reinterpret_cast<const short2 *>(d_samples);
    P.balls = reinterpret_cast<const double2 *>(d_balls);
    P.pts = reinterpret_cast<short2 *>(d_pts);
    P.cost = d_cost;
    P.parent = d_parent;
    P.stats = reinterpret_cast<long long *>(d_stats);
    P.ell_c = d_ell_c;
    if (use_grid) {
        P.sbits = 0; P.hit_words = 0; P.tail_bytes = 0; P.steps_max = 0;
        P.list_cap = g.list_cap;
        P.g_xb = g.xb; P.g_yb = g.yb; P.g_bshift = g.bshift; P.g_bshy = g.bshy; P.g_nbx = g.nbx; P.g_nby = g.nby; P.g_rad = g.rad; P.g_near_ok2 = g.near_ok2;
        P.g_ent_words = g.ent_words; P.g_off_list = g.off_list; P.g_off_bstart = g.off_bstart; P.g_kb = g.kb;
        return grid_launch(kind, g.K, P, nplans, g.smem, st);
    }
    P.g_xb = P.g_yb = P.g_bshift = P.g_bshy = P.g_nbx = P.g_nby = P.g_rad = 0; P.g_near_ok2 = 0;
    P.g_ent_words = P.g_off_list = P.g_off_bstart = P.g_kb = 0;
    P.sbits = s.sbits;
    P.list_cap = s.list_cap;
    P.hit_words = s.hit_words;
    P.tail_bytes = s.tail_bytes;
    P.steps_max = s.steps_max;
    switch (kind) {
        case RRTK_STANDARD: return scan_launch_standard(P, nplans, s.T, s.K, s.two_per_sm, s.smem, st);
        case RRTK_STAR: return scan_launch_star(P, nplans, s.T, s.K, s.two_per_sm, s.smem, st);
        default: return scan_launch_informed(P, nplans, s.T, s.K, s.two_per_sm, s.smem, st);
    }
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

// the same walk, also emitting what a caller of the reference ends up holding for a plan: the path's points
// (vertices_as_ndarray, rrt.py:109-129, as the point sequence) and its cost (vcosts[vgoal])
__global__ void paths_xy_kernel(const int *parent, const short2 *pts, const double *cost, const long long *stats, int nplans, int n,
                                int cap, int *path, short2 *xy, int *len, double *pcost)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nplans) return;
    const size_t row = (size_t)p * (n + 1);
    const int *par = parent + row;
    int v = (int)stats[(size_t)p * RRTK_STAT_COUNT + RRTK_STAT_VGOAL];
    int depth = 0;
    for (int u = v; u > 0 && depth <= n; u = par[u]) ++depth;
    const int L = depth + 1;
    len[p] = L;
    pcost[p] = cost[row + v];
    if (L > cap) return;
    int u = v;
    for (int k = L - 1; k >= 0; --k) {
        path[(size_t)p * cap + k] = u;
        xy[(size_t)p * cap + k] = pts[row + u];
        u = u > 0 ? par[u] : 0;
    }
    for (int k = L; k < cap; ++k) { path[(size_t)p * cap + k] = -1; xy[(size_t)p * cap + k] = make_short2(-32768, -32768); }
}

int paths_xy_launch(const int32_t *d_parent, const int16_t *d_pts, const double *d_cost, const int64_t *d_stats, int nplans, int n, int cap,
                    int32_t *d_path, int16_t *d_xy, int32_t *d_len, double *d_pcost, cudaStream_t st)
{
    paths_xy_kernel<<<(nplans + 127) / 128, 128, 0, st>>>(d_parent, reinterpret_cast<const short2 *>(d_pts), d_cost,
                                                         reinterpret_cast<const long long *>(d_stats), nplans, n, cap, d_path,
                                                         reinterpret_cast<short2 *>(d_xy), d_len, d_pcost);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

// headings of the path vertices (K8 trees): what a caller needs beside the points to re-draw a Dubins path
__global__ void path_heads_kernel(const int *path, const uint8_t *head, int nplans, int n, int cap, uint8_t *out)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (size_t)nplans * cap) return;
    const int v = path[i];
    out[i] = v < 0 ? (uint8_t)255 : head[(i / cap) * (size_t)(n + 1) + v];
}

int path_heads_launch(const int32_t *d_path, const uint8_t *d_head, int nplans, int n, int cap, uint8_t *d_out, cudaStream_t st)
{
    const size_t total = (size_t)nplans * cap;
    if (total == 0) return RRTK_OK;
    path_heads_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_path, d_head, nplans, n, cap, d_out);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
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
