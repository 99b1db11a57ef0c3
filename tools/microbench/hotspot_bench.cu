// hotspot_bench.cu - does it cost anything that all 148 CTAs read the SAME 32 KB vector at the same instant?
// Every CTA's 256 threads each issue 8 x 16-byte ld.relaxed.gpu loads (the step kernel's input-word fetch: 4 KB per
// warp) after a grid-wide rendezvous; `copies` = how many replicas of the vector the CTAs are spread over.
// Prints the mean / max cycles from issue to the last load's arrival over CTAs.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(256, 1) fetch(const uint4* buf, int copies, int iters, long long* out, unsigned* sink) {
    cg::grid_group grid = cg::this_grid();
    const uint4* base = buf + (size_t)(blockIdx.x % copies) * 2048;      // 32 KB = 2048 x 16 B per copy
    long long tot = 0;
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        grid.sync();
        const long long t0 = clock64();
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4* p = base + (threadIdx.x >> 5) * 256 + i * 32 + (threadIdx.x & 31);
            asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w) : "l"(p) : "memory");
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc ^= v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
        __syncthreads();
        const long long t1 = clock64();
        tot += t1 - t0 + (acc == 0x1234567u);
    }
    if (threadIdx.x == 0) out[blockIdx.x] = tot / iters;
    if (acc == 0x7654321u) *sink = acc;
}

int main() {
    const int G = 148, iters = 200;
    uint4* buf; long long* out; unsigned* sink;
    cudaMalloc(&buf, 32768 * 16); cudaMalloc(&out, G * 8); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, 32768 * 16);
    for (int copies : {1, 2, 4, 8, 16}) {
        int it = iters;
        void* args[] = {&buf, &copies, &it, &out, &sink};
        cudaLaunchCooperativeKernel((void*)fetch, dim3(G), dim3(256), args, 0, 0);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[G];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0, sum = 0;
        for (int i = 0; i < G; ++i) { sum += h[i]; if (h[i] > mx) mx = h[i]; }
        printf("copies %2d: issue -> all 8 loads of the CTA back: mean %lld cycles, slowest CTA %lld cycles (%s)\n", copies, sum / G, mx, cudaGetErrorString(e));
    }
    return 0;
}
