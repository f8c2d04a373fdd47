// Sample streams: RRT.sample_all_free (rrt.py:231-240) with the reference's random generator.
//
// The reference draws free[rand_gen.choice(nfree)] with rand_gen = np.random.default_rng(seed)
// (rrt.py:85), i.e. numpy's PCG64 (128-bit LCG, XSL-RR 64-bit output, stepped before output)
// feeding Lemire's bounded 32-bit rejection method from buffered 32-bit halves (low half first).
// numpy is not part of /root/reference; the algorithm is restated from numpy's published source
// (numpy/random/src/pcg64/pcg64.h, numpy/random/src/distributions/distributions.c, numpy 1.21 ..
// 2.3 identical here) and pinned by tests against default_rng(seed).integers(0, nfree, n) and
// against tests/golden/sampler_seed12345.npz (drawn by the reference's own sampler).
//
// rank -> cell: free = argwhere(og == 0) (rrt.py:64) lists free cells row-major, so rank k lies in
// the row x with rowcum[x] <= k < rowcum[x+1] and is the (k - rowcum[x])-th clear bit of that row.
#include "common.cuh"

namespace rrtk {

struct Pcg64 {
    unsigned long long hi, lo, inc_hi, inc_lo;
    bool has32;
    uint32_t buf32;

    __device__ __forceinline__ void step()
    {
        // state = state * 0x2360ED051FC65DA44385DF649FCCF645 + inc  (mod 2^128)
        const unsigned long long mh = 0x2360ED051FC65DA4ull, ml = 0x4385DF649FCCF645ull;
        const unsigned long long l = lo * ml;
        const unsigned long long h = __umul64hi(lo, ml) + hi * ml + lo * mh;
        const unsigned long long nl = l + inc_lo;
        hi = h + inc_hi + (nl < l ? 1ull : 0ull);
        lo = nl;
    }
    __device__ __forceinline__ unsigned long long next64()
    {
        step();
        const unsigned long long x = hi ^ lo;
        const unsigned r = (unsigned)(hi >> 58);
        return (x >> r) | (x << ((64 - r) & 63));
    }
    __device__ __forceinline__ uint32_t next32()
    {
        if (has32) { has32 = false; return buf32; }
        const unsigned long long v = next64();
        has32 = true;
        buf32 = (uint32_t)(v >> 32);
        return (uint32_t)v;
    }
    // Generator.integers(0, bound) for 0 < bound <= 2^32 - 1  (Lemire, 32-bit)
    __device__ __forceinline__ uint32_t bounded(uint32_t bound)
    {
        const uint32_t rng = bound - 1;
        if (rng == 0) return 0;
        unsigned long long m = (unsigned long long)next32() * bound;
        uint32_t left = (uint32_t)m;
        if (left < bound) {
            const uint32_t thr = (0xffffffffu - rng) % bound;
            while (left < thr) {
                m = (unsigned long long)next32() * bound;
                left = (uint32_t)m;
            }
        }
        return (uint32_t)(m >> 32);
    }
};

__device__ __forceinline__ short2 rank_to_cell(const uint32_t *__restrict__ g, const int *__restrict__ rowcum, int W, int TY, int k)
{
    int lo = 0, hi = W;                    // largest x with rowcum[x] <= k
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(rowcum + mid) <= k) lo = mid; else hi = mid;
    }
    int rem = k - __ldg(rowcum + lo);
    int y = 0;
    for (int ty = 0; ty < TY; ++ty) {
        const uint32_t fr = ~__ldg(g + word_index(lo, ty << 5, TY));
        const int c = __popc(fr);
        if (rem < c) { y = (ty << 5) + __fns(fr, 0, rem + 1); break; }
        rem -= c;
    }
    return make_short2((short)lo, (short)y);
}

// one block per plan: thread 0 runs the (inherently sequential) generator into shared memory,
// then the block maps ranks to cells in parallel
__global__ void sample_stream_kernel(const uint32_t *__restrict__ bits, const int *__restrict__ rowcum, int W, int H,
                                     const rrtk_plan_desc *__restrict__ plans, const unsigned long long *__restrict__ state, int n,
                                     short2 *__restrict__ samples, unsigned long long *__restrict__ state_out, uint32_t *__restrict__ carry)
{
    extern __shared__ uint32_t s_rank[];
    const int plan = blockIdx.x;
    const int world = plans[plan].world;
    const int *rc = rowcum + (size_t)world * (W + 1);
    const uint32_t nfree = (uint32_t)__ldg(rc + W);
    if (threadIdx.x == 0) {
        Pcg64 g;
        g.hi = state[4 * plan]; g.lo = state[4 * plan + 1];
        g.inc_hi = state[4 * plan + 2]; g.inc_lo = state[4 * plan + 3];
        g.has32 = false; g.buf32 = 0;
        if (carry) { g.has32 = carry[2 * plan] != 0; g.buf32 = carry[2 * plan + 1]; }   // numpy's has_uint32 / uinteger
        for (int i = 0; i < n; ++i) s_rank[i] = nfree ? g.bounded(nfree) : 0;
        if (state_out) {       // the generator as the next plan() of the same planner object finds it (rrt.py:85: one rand_gen per object)
            state_out[4 * plan] = g.hi; state_out[4 * plan + 1] = g.lo;
            state_out[4 * plan + 2] = g.inc_hi; state_out[4 * plan + 3] = g.inc_lo;
        }
        if (carry) { carry[2 * plan] = g.has32 ? 1u : 0u; carry[2 * plan + 1] = g.buf32; }
    }
    __syncthreads();
    const uint32_t *g = bits + (size_t)world * grid_words(W, H);
    const int TY = tiles_y(H);
    short2 *out = samples + (size_t)plan * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = rank_to_cell(g, rc, W, TY, (int)s_rank[i]);
}

int sample_streams_launch(const uint32_t *d_bits, const int32_t *d_rowcum, int W, int H, const rrtk_plan_desc *d_plans,
                          int nplans, const uint64_t *d_state, int n, int16_t *d_samples, int optin, cudaStream_t st,
                          uint64_t *d_state_out, uint32_t *d_carry)
{
    if (nplans == 0 || n == 0) return RRTK_OK;
    const size_t smem = (size_t)n * 4;
    if (smem > (size_t)optin) {
        set_error("sample stream of n=%d does not fit shared memory", n);
        return RRTK_ERR_CAPACITY;
    }
    RRTK_CUDA(cudaFuncSetAttribute(sample_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sample_stream_kernel<<<nplans, 128, smem, st>>>(d_bits, d_rowcum, W, H, d_plans,
                                                    reinterpret_cast<const unsigned long long *>(d_state), n,
                                                    reinterpret_cast<short2 *>(d_samples),
                                                    reinterpret_cast<unsigned long long *>(d_state_out), d_carry);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

}  // namespace rrtk
