// What does one pass of the batched kernel's MMA loop cost its issuing warp?  One warp, M = 128 x N = 32 x K = 16 MMAs on
// four accumulators (operands: whatever is in shared memory), with / without tcgen05.commit, mbarrier try_wait, fences.
// nvcc -gencode arch=compute_100a,code=sm_100a -o umma_issue_bench.bin umma_issue_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a) {
    const uint32_t lo = ((a & 0x3FFFFu) >> 4) | (1u << 16), hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
#ifndef WAITKIND
#define WAITKIND 0
#endif
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
#if WAITKIND == 0
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#elif WAITKIND == 1
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.relaxed.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#else
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
    return ok != 0;
}
template <int NMMA, int NCOMMIT, int NWAIT, int FENCE, int MDIM, int NDIM = 32, int NW = 1>
__global__ void __launch_bounds__(160, 1) bench(long long* out, int passes) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars[8];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(1 << 20));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NDIM >> 3) << 17) | ((uint32_t)(MDIM >> 4) << 24);
    if (warp < NW) {
        const long long t0 = clock64();
        for (int it = 0; it < passes; ++it) {
#pragma unroll
            for (int w = 0; w < NWAIT; ++w)
                while (!try_wait(&bars[4 + w], 1u)) {}
            if (FENCE) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a = smem_u32(smem) + (uint32_t)(it & 7) * 16384, b = smem_u32(smem) + 131072 + (uint32_t)(it & 3) * 8192;
            const uint64_t ad = desc(a), bd = desc(b);
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < NMMA; ++j)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem + warp * 128 + (j & 3) * 32), "l"(ad + 2 * (j & 3)), "l"(bd + 2 * (j & 3)), "r"(idesc), "r"(it ? 1u : 0u) : "memory");
#pragma unroll
                for (int cc = 0; cc < NCOMMIT; ++cc)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[cc])) : "memory");
            }
            __syncwarp();
        }
        // drain
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[3])) : "memory");
        __syncwarp();
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) out[2 * warp] = t1 - t0;
        // wait until everything retired: poll pending count is awkward; sleep instead
        for (int i = 0; i < 2000; ++i) __nanosleep(1000);
        if ((threadIdx.x & 31) == 0) out[2 * warp + 1] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int NMMA, int NCOMMIT, int NWAIT, int FENCE, int MDIM, int NDIM = 32, int NW = 1>
void run(const char* name, long long* d) {
    const int smem = 200 * 1024, passes = 4000;
    cudaFuncSetAttribute(bench<NMMA, NCOMMIT, NWAIT, FENCE, MDIM, NDIM, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long h[2];
    for (int rep = 0; rep < 2; ++rep) {
        bench<NMMA, NCOMMIT, NWAIT, FENCE, MDIM, NDIM, NW><<<1, 160, smem>>>(d, passes);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-44s %7.1f cycles / pass (issue side)\n", name, (double)h[0] / passes);
}
int main() {
    long long* d;
    cudaMalloc(&d, 256);
    printf("wait kind %d (0 try_wait, 1 try_wait.relaxed.cta, 2 test_wait)\n", WAITKIND);
    run<0, 0, 1, 0, 128>("1 wait only", d);
    run<0, 0, 3, 0, 128>("3 waits only", d);
    run<4, 0, 1, 0, 128>("4 MMA + 1 wait", d);
    run<4, 0, 3, 0, 128>("4 MMA + 3 waits", d);
    run<0, 0, 0, 0, 128>("empty pass (elect only)", d);
    run<4, 0, 0, 0, 128>("4 MMA (M128 N32)", d);
    run<8, 0, 0, 0, 128>("8 MMA (M128 N32)", d);
    run<4, 0, 0, 0, 64>("4 MMA (M64 N32)", d);
    run<4, 1, 0, 0, 128>("4 MMA + 1 commit", d);
    run<4, 2, 0, 0, 128>("4 MMA + 2 commits", d);
    run<4, 3, 0, 0, 128>("4 MMA + 3 commits", d);
    run<0, 1, 0, 0, 128>("1 commit only", d);
    run<0, 2, 0, 0, 128>("2 commits only", d);
    run<4, 2, 1, 0, 128>("4 MMA + 2 commits + 1 try_wait", d);
    run<4, 2, 3, 0, 128>("4 MMA + 2 commits + 3 try_wait", d);
    run<4, 2, 3, 1, 128>("4 MMA + 2 commits + 3 try_wait + fence", d);
    run<16, 2, 1, 1, 128>("16 MMA + 2 commits + 1 try_wait + fence", d);
    run<8, 0, 0, 0, 128, 16>("8 MMA M128 N16", d);
    run<8, 0, 0, 0, 128, 64>("8 MMA M128 N64", d);
    run<8, 0, 0, 0, 128, 128>("8 MMA M128 N128", d);
    run<8, 0, 0, 0, 128, 256>("8 MMA M128 N256 (TMEM cols wrap, timing only)", d);
    run<8, 0, 0, 0, 64, 8>("8 MMA M64 N8", d);
    run<8, 0, 0, 0, 64, 32>("8 MMA M64 N32", d);
    run<8, 0, 0, 0, 64, 112>("8 MMA M64 N112", d);
    run<8, 0, 0, 0, 128, 32, 2>("8 MMA M128 N32, 2 warps issuing (per warp)", d);
    run<8, 0, 0, 0, 128, 32, 4>("8 MMA M128 N32, 4 warps issuing (per warp)", d);
    run<8, 1, 0, 0, 128, 32, 4>("8 MMA + 1 commit, 4 warps issuing (per warp)", d);
    return 0;
}
