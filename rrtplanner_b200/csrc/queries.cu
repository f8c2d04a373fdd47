// K2 / K3: stand-alone batched forms of RRT.near(...)[0] (rrt.py:131-155) and RRT.within
// (rrt.py:157-181) over caller-supplied vertex arrays, plus the distance keys / stable ordering
// needed to return the full permutation RRT.near yields.  Inside plan() these queries are fused
// into the plan kernel (plan.cu); the kernels here serve the static-method API and the
// micro-benchmarks.  int32 coordinates, exact int64 squared distances.
#include "common.cuh"

namespace rrtk {

// Coordinate traits.  Integer vertices (what plan() holds, rrt.py:408) use exact int64 squared
// distances; ordering by d^2 equals ordering by np.linalg.norm for |d^2| < 2^52.  Floating-point
// vertices (the reference's own within() test passes x = [0.5, 0.5], tests/test_rrt.py:116-119)
// follow numpy literally: d2 = dx*dx + dy*dy without contraction, near() orders by sqrt(d2).
struct IntCoord {
    typedef int2 P;
    typedef long long K;
    static __device__ __forceinline__ K d2(P p, P q)
    {
        const long long dx = (long long)p.x - q.x, dy = (long long)p.y - q.y;
        return dx * dx + dy * dy;
    }
    static __device__ __forceinline__ K near_key(P p, P q) { return d2(p, q); }
    static __device__ __forceinline__ K worst() { return 0x7fffffffffffffffll; }
};
struct F64Coord {
    typedef double2 P;
    typedef double K;
    static __device__ __forceinline__ K d2(P p, P q)
    {
        const double dx = __dsub_rn(p.x, q.x), dy = __dsub_rn(p.y, q.y);
        return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    }
    static __device__ __forceinline__ K near_key(P p, P q) { return __dsqrt_rn(d2(p, q)); }
    static __device__ __forceinline__ K worst() { return __longlong_as_double(0x7ff0000000000000ll); }
};

// one block per query; coalesced loads; warp-shuffle argmin with lowest-index tie rule
template <class C>
__global__ void nearest_kernel(const typename C::P *__restrict__ pts, int npts, const typename C::P *__restrict__ queries,
                               const int *__restrict__ count, int *__restrict__ idx_out, typename C::K *__restrict__ key_out)
{
    typedef typename C::K K;
    __shared__ K s_d[32];
    __shared__ int s_i[32];
    const int q = blockIdx.x;
    const typename C::P qp = queries[q];
    const int m = count ? min(__ldg(count + q), npts) : npts;
    K bd = C::worst();
    int bi = 0x7fffffff;
    for (int v = threadIdx.x; v < m; v += blockDim.x) {
        const K d = C::near_key(pts[v], qp);
        if (d < bd || bi == 0x7fffffff) { bd = d; bi = v; }      // ascending v per thread: first minimum kept
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const K od = __shfl_xor_sync(RRTK_FULL, bd, o);
        const int oi = __shfl_xor_sync(RRTK_FULL, bi, o);
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) { s_d[warp] = bd; s_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        bd = lane < nw ? s_d[lane] : C::worst();
        bi = lane < nw ? s_i[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const K od = __shfl_xor_sync(RRTK_FULL, bd, o);
            const int oi = __shfl_xor_sync(RRTK_FULL, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if (lane == 0) {
            idx_out[q] = m > 0 ? bi : -1;
            if (key_out) key_out[q] = bd;
        }
    }
}

int nearest_launch(const int32_t *d_pts, int npts, const int32_t *d_queries, const int32_t *d_count, int nq, int32_t *d_idx,
                   int64_t *d_d2, cudaStream_t st)
{
    if (nq == 0) return RRTK_OK;
    const int threads = npts >= 4096 ? 256 : 128;
    nearest_kernel<IntCoord><<<nq, threads, 0, st>>>(reinterpret_cast<const int2 *>(d_pts), npts,
                                                    reinterpret_cast<const int2 *>(d_queries), d_count, d_idx,
                                                    reinterpret_cast<long long *>(d_d2));
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}
int nearest_launch_f64(const double *d_pts, int npts, const double *d_queries, const int32_t *d_count, int nq, int32_t *d_idx,
                       double *d_dist, cudaStream_t st)
{
    if (nq == 0) return RRTK_OK;
    const int threads = npts >= 4096 ? 256 : 128;
    nearest_kernel<F64Coord><<<nq, threads, 0, st>>>(reinterpret_cast<const double2 *>(d_pts), npts,
                                                    reinterpret_cast<const double2 *>(d_queries), d_count, d_idx, d_dist);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

// one block per query; ballot + prefix compaction keeps np.argwhere's ascending order (rrt.py:180)
template <class C>
__global__ void within_kernel(const typename C::P *__restrict__ pts, int npts, const typename C::P *__restrict__ queries,
                              const int *__restrict__ count, typename C::K r2_excl, int cap, int *__restrict__ out,
                              int *__restrict__ len_out)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int q = blockIdx.x;
    const typename C::P qp = queries[q];
    const int m = count ? min(__ldg(count + q), npts) : npts;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int *o = out + (size_t)q * cap;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < m; base += blockDim.x) {
        const int v = base + threadIdx.x;
        const bool in = v < m && C::d2(pts[v], qp) < r2_excl;
        const unsigned b = __ballot_sync(RRTK_FULL, in);
        if (lane == 0) s_warp[warp] = __popc(b);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        const int pos = before + __popc(b & ((1u << lane) - 1));
        if (in && pos < cap) o[pos] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = s_base;
            for (int w = 0; w < nw; ++w) t += s_warp[w];
            s_base = t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) len_out[q] = s_base;
}

int within_launch(const int32_t *d_pts, int npts, const int32_t *d_queries, const int32_t *d_count, int nq, double r, int cap,
                  int32_t *d_out, int32_t *d_len, cudaStream_t st)
{
    if (nq == 0) return RRTK_OK;
    // (double)d2 < r*r  <=>  d2 < ceil(r*r) for integer d2 >= 0
    const double rr = r * r;
    long long excl = rr >= 9.0e18 ? 0x7fffffffffffffffll : (long long)ceil(rr);
    if (!(rr > 0.0)) excl = 0;
    within_kernel<IntCoord><<<nq, 256, 0, st>>>(reinterpret_cast<const int2 *>(d_pts), npts,
                                               reinterpret_cast<const int2 *>(d_queries), d_count, excl, cap, d_out, d_len);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}
int within_launch_f64(const double *d_pts, int npts, const double *d_queries, const int32_t *d_count, int nq, double r, int cap,
                      int32_t *d_out, int32_t *d_len, cudaStream_t st)
{
    if (nq == 0) return RRTK_OK;
    within_kernel<F64Coord><<<nq, 256, 0, st>>>(reinterpret_cast<const double2 *>(d_pts), npts,
                                               reinterpret_cast<const double2 *>(d_queries), d_count, r * r, cap, d_out, d_len);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

template <class C>
__global__ void dist_key_kernel(const typename C::P *__restrict__ pts, int npts, typename C::P q, typename C::K *__restrict__ out)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < npts) out[v] = C::near_key(pts[v], q);
}

int dist2_launch(const int32_t *d_pts, int npts, int qx, int qy, int64_t *d_d2, cudaStream_t st)
{
    if (npts == 0) return RRTK_OK;
    dist_key_kernel<IntCoord><<<(npts + 255) / 256, 256, 0, st>>>(reinterpret_cast<const int2 *>(d_pts), npts, make_int2(qx, qy),
                                                                 reinterpret_cast<long long *>(d_d2));
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}
int dist_launch_f64(const double *d_pts, int npts, double qx, double qy, double *d_dist, cudaStream_t st)
{
    if (npts == 0) return RRTK_OK;
    dist_key_kernel<F64Coord><<<(npts + 255) / 256, 256, 0, st>>>(reinterpret_cast<const double2 *>(d_pts), npts,
                                                                 make_double2(qx, qy), d_dist);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

// ---- stable argsort of int64 keys: bitonic network over (key, index) pairs -----------------------
struct KeyIdx { long long k; long long i; };

__global__ void argsort_fill_kernel(const long long *__restrict__ keys, int n, int npow2, KeyIdx *__restrict__ a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < npow2) a[t] = t < n ? KeyIdx{keys[t], t} : KeyIdx{0x7fffffffffffffffll, 0x7fffffffll + t};
}
__global__ void argsort_step_kernel(KeyIdx *__restrict__ a, int npow2, int k, int jj)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = t ^ jj;
    if (t >= npow2 || p <= t) return;
    const KeyIdx x = a[t], y = a[p];
    const bool up = (t & k) == 0;
    const bool gt = x.k > y.k || (x.k == y.k && x.i > y.i);
    if (gt == up) { a[t] = y; a[p] = x; }
}
__global__ void argsort_out_kernel(const KeyIdx *__restrict__ a, int n, int *__restrict__ perm)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) perm[t] = (int)a[t].i;
}

static int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }
size_t argsort_scratch(int n) { return (size_t)next_pow2(n < 1 ? 1 : n) * sizeof(KeyIdx); }

int argsort_launch(const int64_t *d_keys, int n, int32_t *d_perm, void *d_scratch, size_t scratch_bytes, cudaStream_t st)
{
    if (n == 0) return RRTK_OK;
    const int np2 = next_pow2(n);
    if (scratch_bytes < (size_t)np2 * sizeof(KeyIdx)) {
        set_error("argsort scratch too small: need %zu bytes", (size_t)np2 * sizeof(KeyIdx));
        return RRTK_ERR_INVALID;
    }
    KeyIdx *a = reinterpret_cast<KeyIdx *>(d_scratch);
    const int blocks = (np2 + 255) / 256;
    argsort_fill_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const long long *>(d_keys), n, np2, a);
    for (int k = 2; k <= np2; k <<= 1)
        for (int jj = k >> 1; jj > 0; jj >>= 1) argsort_step_kernel<<<blocks, 256, 0, st>>>(a, np2, k, jj);
    argsort_out_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, n, d_perm);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

}  // namespace rrtk
