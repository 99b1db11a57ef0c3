// stream_bench.cu - what read bandwidth does the step kernel's weight-streaming scheme reach on its own?
// 148 CTAs (one per SM), each walks ITS OWN contiguous byte range with 1-D bulk async copies (TMA engine) into a
// shared-memory ring; 8 consumer warps only wait for a slot and release it.  Sweeps slot size and ring depth, and a
// plain ld.global.v4 grid-stride read for comparison.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

__global__ void __launch_bounds__(320, 1) ring_kernel(const unsigned char* base, size_t per_cta, int slot_bytes, int n_slots, int hint) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)slot_bytes * n_slots);
    uint64_t* empty = full + n_slots;
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_slots; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned char* src = base + (size_t)blockIdx.x * per_cta;
    const unsigned n = (unsigned)(per_cta / slot_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 9) {
        if (lane == 0) {
            uint64_t pol;
            if (hint) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
            for (unsigned i = 0; i < n; ++i) {
                const unsigned s = i % n_slots, ph = (i / n_slots) & 1u;
                mbar_wait(empty + s, ph ^ 1u);
                mbar_expect(full + s, slot_bytes);
                bulk(smem + (size_t)s * slot_bytes, src + (size_t)i * slot_bytes, slot_bytes, full + s, pol);
            }
        }
    } else if (warp < 8) {
        for (unsigned i = warp; i < n; i += 8) {
            const unsigned s = i % n_slots, ph = (i / n_slots) & 1u;
            mbar_wait(empty + s, ph ^ 1u);        // generation g-1 released: pins the full barrier to generation g
            mbar_wait(full + s, ph);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
    }
}

__global__ void __launch_bounds__(1024, 1) ldg_kernel(const uint4* base, size_t n16, unsigned* sink) {
    unsigned acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(base + i));
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main() {
    const int G = 148;
    const size_t per_cta = 24ull << 20;              // 24 MB per CTA, 3.5 GB in total (>> L2)
    unsigned char* buf; unsigned* sink;
    cudaMalloc(&buf, per_cta * G); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, per_cta * G);
    printf("alloc: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    const int cfgs[][2] = {{8192, 20}, {8192, 24}, {7168, 20}, {16384, 10}, {16384, 12}, {4096, 40}, {32768, 6}, {8192, 8}, {8192, 12}};
    for (auto& c : cfgs) for (int hint = 0; hint < 2; ++hint) {
        const size_t pc = per_cta / c[0] * c[0];
        const size_t smem = (size_t)c[0] * c[1] + 16 * c[1];
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            ring_kernel<<<G, 320, smem>>>(buf, pc, c[0], c[1], hint);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("ring  slot %5d B x %2d slots (%3zu KB/SM) %s: %7.1f GB/s  (%s)\n", c[0], c[1], smem / 1024,
               hint ? "evict_first " : "evict_normal", pc * G / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    for (int blocks : {148, 296, 592}) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            ldg_kernel<<<blocks, 1024>>>(reinterpret_cast<const uint4*>(buf), per_cta * G / 16, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("ld.global.v4 grid-stride, %3d x 1024 threads: %7.1f GB/s  (%s)\n", blocks, per_cta * G / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
