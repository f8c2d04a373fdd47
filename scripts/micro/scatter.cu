// Micro-benchmark: cost of scattered 1-byte reads of a 4 MB field (the access pattern of K1b, csrc/clearance.cu)
// through (a) LDG, (b) a pitch-linear 2-D texture, (c) a cudaArray texture.    nvcc -arch=sm_100a -O3 -o scatter scatter.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint32_t mix(uint32_t h) { h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16; return h; }
template <int MODE>
__global__ void k(const uint8_t *buf, cudaTextureObject_t tex, int n, int iters, uint32_t *out, int local)
{
    uint32_t h = mix(blockIdx.x * blockDim.x + threadIdx.x + 12345u);
    uint32_t acc = 0;
    int x = h & 2047, y = (h >> 11) & 2047;
    for (int i = 0; i < iters; ++i) {
        uint32_t v;
        if (MODE == 0) v = __ldg(buf + x * 2048 + y);
        else v = tex2D<unsigned char>(tex, (float)y, (float)x);
        acc += v;
        h = mix(h + v);                        // next address depends on the value, as in the clearance walk
        if (local) { x = (x + (h & 15) + 1) & 2047; y = (y + ((h >> 4) & 15)) & 2047; }   // a walk: next cell 1..16 away
        else { x = h & 2047; y = (h >> 11) & 2047; }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
int main()
{
    const int n = 2048 * 2048;
    uint8_t *h = (uint8_t *)malloc(n);
    for (int i = 0; i < n; ++i) h[i] = (uint8_t)(i * 2654435761u >> 24);
    uint8_t *d; uint32_t *out;
    CK(cudaMalloc(&d, n)); CK(cudaMalloc(&out, 4)); CK(cudaMemcpy(d, h, n, cudaMemcpyHostToDevice));
    cudaResourceDesc rd = {}; cudaTextureDesc td = {};
    rd.resType = cudaResourceTypePitch2D; rd.res.pitch2D.devPtr = d; rd.res.pitch2D.desc = cudaCreateChannelDesc<unsigned char>();
    rd.res.pitch2D.width = 2048; rd.res.pitch2D.height = 2048; rd.res.pitch2D.pitchInBytes = 2048;
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tp, ta;
    CK(cudaCreateTextureObject(&tp, &rd, &td, nullptr));
    cudaArray_t arr; cudaChannelFormatDesc cd = cudaCreateChannelDesc<unsigned char>();
    CK(cudaMallocArray(&arr, &cd, 2048, 2048));
    CK(cudaMemcpy2DToArray(arr, 0, 0, h, 2048, 2048, 2048, cudaMemcpyHostToDevice));
    cudaResourceDesc ra = {}; ra.resType = cudaResourceTypeArray; ra.res.array.array = arr;
    CK(cudaCreateTextureObject(&ta, &ra, &td, nullptr));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256, iters = 64;
    for (int local = 0; local < 2; ++local)
        for (int mode = 0; mode < 3; ++mode) {
            float best = 1e9;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<blocks, threads>>>(d, tp, n, iters, out, local);
                else if (mode == 1) k<1><<<blocks, threads>>>(d, tp, n, iters, out, local);
                else k<1><<<blocks, threads>>>(d, ta, n, iters, out, local);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            const double loads = (double)blocks * threads * iters;
            printf("%s %s: %.3f ms, %.1f G loads/s, %.2f cycles/lane-load/SM at 1965 MHz\n", local ? "walk  " : "random", mode == 0 ? "ldg      " : mode == 1 ? "tex pitch" : "tex array",
                   best, loads / best / 1e6, best * 1e-3 * 1.965e9 * 148 / loads);
        }
    return 0;
}
