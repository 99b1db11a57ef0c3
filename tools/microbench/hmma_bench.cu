// Microbenchmark: legacy mma.sync (HMMA.16816 bf16) and ldmatrix latency / throughput on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int CHAINS>
__global__ void k_mma(long long* out, float* sink, int iters) {
    float acc[CHAINS][4];
    for (int c = 0; c < CHAINS; ++c) for (int j = 0; j < 4; ++j) acc[c][j] = threadIdx.x * 0.001f + c;
    uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
    uint32_t b0 = 0x3f803f80u, b1 = 0x3f803f80u;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) mma(acc[c], a, b0, b1);
    }
    long long t1 = clock64();
    float s = 0; for (int c = 0; c < CHAINS; ++c) for (int j = 0; j < 4; ++j) s += acc[c][j];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
__global__ void k_ffma2(long long* out, float* sink, int iters) {
    unsigned long long acc[8]; for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x + i;
    unsigned long long x = 0x3f8000003f800000ull, w = 0x3f0000003f000000ull;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 8; ++c) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[c]) : "l"(x), "l"(w));
    }
    long long t1 = clock64();
    unsigned long long s = 0; for (int c = 0; c < 8; ++c) s += acc[c];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
int main() {
    long long* d; float* sink; cudaMalloc(&d, 64); cudaMalloc(&sink, 1 << 20);
    long long h;
    const int iters = 2000;
    auto run = [&](const char* name, auto kern, int warps, int chains) {
        kern<<<148, warps * 32>>>(d, sink, iters); cudaDeviceSynchronize();
        kern<<<148, warps * 32>>>(d, sink, iters); cudaDeviceSynchronize();
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("%-28s warps/SM=%2d chains=%d : %.1f cycles per op per warp, %.2f cycles/op per SM\n", name, warps, chains,
               (double)h / iters / chains, (double)h / iters / chains / warps);
    };
    run("HMMA dependent (latency)", k_mma<1>, 1, 1);
    run("HMMA 2 chains", k_mma<2>, 1, 2);
    run("HMMA 4 chains", k_mma<4>, 1, 4);
    run("HMMA 8 chains", k_mma<8>, 1, 8);
    run("HMMA 8 chains", k_mma<8>, 4, 8);
    run("HMMA 8 chains", k_mma<8>, 8, 8);
    run("HMMA 4 chains", k_mma<4>, 16, 4);
    run("FFMA2 8 chains", k_ffma2, 1, 8);
    run("FFMA2 8 chains", k_ffma2, 8, 8);
    run("FFMA2 8 chains", k_ffma2, 16, 8);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
