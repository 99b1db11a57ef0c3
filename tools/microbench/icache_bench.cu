// icache_bench.cu - what does instruction footprint cost on B200?
// N distinct straight-line blocks of ~2 KB SASS each are executed round-robin by 8 warps of one CTA per SM;
// prints cycles per block as the total footprint grows past the instruction caches.
#include <cstdio>
#include <cuda_runtime.h>

template <int I>
__device__ __noinline__ float blk(float x, float y) {
    float a = x, b = y, c = x + 1.f, d = y + 2.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        a = fmaf(a, 1.0001f + I * 0.001f + j * 0.01f, 0.5f);
        b = fmaf(b, 0.9999f + I * 0.002f + j * 0.02f, 0.25f);
        c = fmaf(c, 1.0002f + I * 0.003f + j * 0.03f, 0.125f);
        d = fmaf(d, 0.9998f + I * 0.004f + j * 0.04f, 0.0625f);
    }
    return a + b + c + d;
}

template <int N>
struct Run {
    static __device__ __forceinline__ float go(float x, int n_active) {
        x = Run<N - 1>::go(x, n_active);
        if (N - 1 < n_active) x = blk<N - 1>(x, x * 0.5f);
        return x;
    }
};
template <>
struct Run<0> {
    static __device__ __forceinline__ float go(float x, int) { return x; }
};

constexpr int kMaxBlocks = 128;

__global__ void __launch_bounds__(256, 1) bench(float* out, long long* cyc, int n_active, int iters) {
    float x = threadIdx.x * 0.001f;
    x = Run<kMaxBlocks>::go(x, n_active);          // warm
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) x = Run<kMaxBlocks>::go(x, n_active);
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 256 * 4);
    cudaMalloc(&cyc, 8);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, bench);
    printf("kernel binary %zu bytes, %d regs\n", (size_t)fa.binaryVersion, fa.numRegs);
    for (int n : {1, 2, 4, 8, 12, 16, 20, 24, 32, 48, 64, 96, 128}) {
        const int iters = 2048 / n + 1;
        bench<<<148, 256>>>(out, cyc, n, iters);
        cudaDeviceSynchronize();
        bench<<<148, 256>>>(out, cyc, n, iters);
        cudaDeviceSynchronize();
        long long h;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("blocks %3d (~%4d KB of code): %7.1f cycles per block per warp (8 warps/SM)\n", n, n * 2,
               (double)h / ((double)iters * n));
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
