// Measured roofline denominators for the two on-chip levels the plan and collision kernels live in.
//
// MEASURED_PEAKS.json (driver-written) carries HBM and bf16 only; SURVEY.md section 8(d) bounds the
// nearest / radius scan by shared-memory bandwidth and the cfg2 collision walk by L2 bandwidth, so
// bench.py measures those two on the box it runs on, with the simplest kernels that can saturate them:
//
//   L2     every thread streams a buffer that fits L2 but not L1 (default 48 MB) with 16-byte loads,
//          eight independent loads in flight per thread, `passes` times; the first pass (HBM -> L2) is
//          issued untimed by the caller as warm-up.  Bytes / time = L2 -> SM read bandwidth.
//   smem   every block fills its shared memory once, then every thread reads it with conflict-free
//          LDS.128 (consecutive threads, consecutive 16-byte words), eight loads in flight per thread.
//
// Both fold what they read into one word per thread that is stored behind a condition the compiler
// cannot resolve, so no load is dropped.
#include "common.cuh"

namespace rrtk {

__global__ void __launch_bounds__(512) l2_read_kernel(const uint4 *__restrict__ buf, size_t n16, int passes, uint32_t *sink)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (int p = 0; p < passes; ++p) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 7 * stride < n16; i += 8 * stride) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(buf + i + u * stride));
#pragma unroll
            for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
        }
        for (; i < n16; i += stride) {
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + i));
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x9e3779b9u && passes < 0) sink[0] = acc;
}

__global__ void __launch_bounds__(1024) smem_read_kernel(int words16, int iters, uint32_t *sink)
{
    extern __shared__ __align__(16) uint4 s_buf[];
    for (int i = threadIdx.x; i < words16; i += blockDim.x) s_buf[i] = make_uint4(i, i * 3, i * 5, i * 7);
    __syncthreads();
    uint32_t acc = 0;
    const int per = words16 / (int)blockDim.x;       // 16-byte words per thread and sweep (host makes it a multiple of 8)
    for (int it = 0; it < iters; ++it) {
        for (int k = 0; k < per; k += 8) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = s_buf[(k + u) * blockDim.x + threadIdx.x];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
        }
        asm volatile("" ::: "memory");
    }
    if (acc == 0x9e3779b9u && iters < 0) sink[0] = acc;
}

int l2_read_launch(const void *d_buf, size_t bytes, int passes, int sm_count, uint32_t *d_sink, cudaStream_t st)
{
    const size_t n16 = bytes / 16;
    l2_read_kernel<<<sm_count * 4, 512, 0, st>>>(reinterpret_cast<const uint4 *>(d_buf), n16, passes, d_sink);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

int smem_read_launch(int smem_bytes, int iters, int sm_count, uint32_t *d_sink, cudaStream_t st)
{
    const int threads = 1024;
    int words16 = smem_bytes / 16;
    words16 -= words16 % (threads * 8);
    if (words16 <= 0) return RRTK_ERR_INVALID;
    RRTK_CUDA(cudaFuncSetAttribute(smem_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, words16 * 16));
    smem_read_kernel<<<sm_count, threads, (size_t)words16 * 16, st>>>(words16, iters, d_sink);
    RRTK_CUDA(cudaGetLastError());
    return RRTK_OK;
}

}  // namespace rrtk
