// mma_sp_probe.cu - which metadata bits of mma.sp m16n8k32 (bf16) govern which (row, 4-column group) of A?
// A (compressed 16 x 16) = 2^(group) for every kept element, B[k][n] = k, baseline metadata = indices (0, 1) in every
// group.  Moving ONE nibble of ONE thread's metadata word to (2, 3) adds 4 * 2^group to exactly one row of D: the
// table of (thread, nibble) -> (row, group) is the metadata layout.  Run for both sparsity selectors.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

template <int SEL>
__device__ __forceinline__ void mma_sp(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[4], uint32_t e) {
    asm volatile(
        "mma.sp::ordered_metadata.sync.aligned.m16n8k32.row.col.f32.bf16.bf16.f32 "
        "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9,%10,%11}, {%0,%1,%2,%3}, %12, %13;"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(e), "n"(SEL));
}
__device__ uint32_t pack_bf16(float lo, float hi) {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
}

template <int SEL>
__global__ void probe(float* out /*[1 + 32*8][16]*/) {
    const int lane = threadIdx.x, g = lane >> 2, q = lane & 3;
    // A compressed [16 rows][16 ccols], dense m16n8k16 fragment layout: a0 (row g, cc 2q,2q+1), a1 (row g+8, same),
    // a2 (row g, cc 2q+8, 2q+9), a3 (row g+8, ...).  value = 2^(cc / 2)
    uint32_t a[4];
    a[0] = pack_bf16(exp2f((float)q), exp2f((float)q));
    a[1] = a[0];
    a[2] = pack_bf16(exp2f((float)(q + 4)), exp2f((float)(q + 4)));
    a[3] = a[2];
    // B [32 k][8 n] = k: b_i holds k = 2q + 8 i, 2q + 8 i + 1 of column n = g
    uint32_t b[4];
    for (int i = 0; i < 4; ++i) b[i] = pack_bf16((float)(2 * q + 8 * i), (float)(2 * q + 8 * i + 1));
    for (int t = -1; t < 32 * 8; ++t) {
        uint32_t e = 0x44444444u;
        if (t >= 0 && (t >> 3) == lane) { const int j = t & 7; e = (e & ~(0xFu << (4 * j))) | (0xEu << (4 * j)); }
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        mma_sp<SEL>(d, a, b, e);
        // column 0 of D: held by lanes with q == 0 (c0 = row g, c2 = row g + 8)
        if (q == 0) { out[(t + 1) * 16 + g] = d[0]; out[(t + 1) * 16 + g + 8] = d[2]; }
    }
}

int main() {
    float* d; cudaMalloc(&d, sizeof(float) * (1 + 256) * 16);
    static float h[(1 + 256) * 16];
    for (int sel = 0; sel < 2; ++sel) {
        cudaMemset(d, 0, sizeof(h));
        if (sel == 0) probe<0><<<1, 32>>>(d); else probe<1><<<1, 32>>>(d);
        cudaError_t err = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("selector %d: %s; baseline D[:,0] =", sel, cudaGetErrorString(err));
        for (int r = 0; r < 16; ++r) printf(" %.0f", h[r]);
        printf("\n");
        for (int t = 0; t < 256; ++t) {
            for (int r = 0; r < 16; ++r) {
                const float delta = h[(t + 1) * 16 + r] - h[r];
                if (delta != 0.f) {
                    int grp = -1;
                    for (int gg = 0; gg < 8; ++gg) if (delta == 4.f * (float)(1 << gg)) grp = gg;
                    printf("  sel %d lane %2d (g %d q %d) nibble %d -> row %2d group %d (delta %.0f)\n", sel, t >> 3, (t >> 3) >> 2,
                           (t >> 3) & 3, t & 7, r, grp, delta);
                }
            }
        }
    }
    return 0;
}
