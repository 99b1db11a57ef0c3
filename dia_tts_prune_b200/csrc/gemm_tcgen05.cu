// gemm_tcgen05.cu - the dense contractions OFF the decode step (encoder, cross-attention K/V precompute, prompt
// prefill: SURVEY.md 8(f) rank 1) on the 5th-generation tensor cores.
//
//   Y[M][N] (fp32) = X[M][K] (fp32) . W[K][N] (bf16)           DenseGeneral.forward, dia/layers.py:55-66, M = B*T rows
//
// The reference runs these in fp32 and the K/V a prefill writes are re-read by every later greedy step, so the
// activations may not be rounded to bf16 (hard part 2).  X is therefore split into three bf16 terms
// (x = hi + lo + lo2, exact to 24 bits) and each k-step issues three tcgen05.mma into the SAME fp32 accumulator in
// tensor memory: the product has fp32-operand accuracy at bf16 tensor-core speed.
//
// Structure (persistent: one CTA per SM walks the 128 x 256 output tiles, warp-specialised):
//   warp 0    one lane: TMA producer - cp.async.bulk.tensor loads of the weight tile [256 x 64] of a k-block into a
//             3-slot ring and of the three A-term tiles [128 x 64] into a 7-slot ring (both K-major, 128-byte swizzle);
//             the rings run on across tile boundaries, so the next tile's operands are in flight during a tile's tail
//   warp 1    allocates all 512 TMEM columns = two 128 x 256 fp32 accumulators; the warp walks the k-blocks and one
//             elected lane issues tcgen05.mma (M = 128, N = 256, K = 16, kind::f16, bf16 inputs, fp32 accumulate),
//             releases ring slots with tcgen05.commit and hands a finished accumulator to the epilogue
//   warps 2-5 epilogue of tile i (tcgen05.ld, 32 lanes x 32 columns per instruction -> registers -> residual add ->
//             fp32 stores) while the MMAs of tile i + 1 fill the other accumulator
// W is kept as a K-major bf16 copy [N][K] made once per weight (transpose_to_bf16_kernel), so that A and B use the
// same canonical UMMA layout.  Every wait is bounded (a stuck pipeline traps instead of hanging the GPU).
#include <algorithm>

#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "engine_internal.h"

namespace dia {

namespace {

// BN = 256: the three A-term tiles (48 KB per k-block) are shared by twice as many output columns - the kernel is bound
// by the operand traffic from L2 (BN = 128 moved 4.3 GB for M = 2048, N = 16384, K = 2048: 13 TB/s at 326 us), and a
// 256-wide MMA (128-cycle floor) also hides the ~100 cycles one tcgen05.mma takes to issue.
constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kTerms = 3;
constexpr int kASlots = 7, kBSlots = 3;                     // A-term ring / weight ring: > 2 k-blocks of operands in flight
constexpr int kTileBytes = BM * BK * 2;                     // 16 KB: one [128 x 64] bf16 tile
constexpr int kBTileBytes = BN * BK * 2;                    // 32 KB: the weight tile
constexpr int kGemmThreads = 192;                           // 6 warps
constexpr int kGemmBars = 2 * kASlots + 2 * kBSlots + 4;
constexpr int kGemmSmem = kASlots * kTileBytes + kBSlots * kBTileBytes + 1024 /* alignment slack */ + 256 /* barriers */;
static_assert(kGemmBars * 8 + 8 <= 256, "barrier region");

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    unsigned polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++polls > 20000000u) __trap();                  // seconds: the pipeline is stuck
    }
}
// K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (stride byte offset), version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    const uint32_t lo = ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace

// x = hi + lo + lo2 per element, three stacked bf16 matrices [3][M][K]
__global__ void split3_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xs, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const float r1 = v - __bfloat162float(h);
        const __nv_bfloat16 l = __float2bfloat16_rn(r1);
        const float r2 = r1 - __bfloat162float(l);
        xs[i] = h;
        xs[n + i] = l;
        xs[2 * n + i] = __float2bfloat16_rn(r2);
    }
}

// the same with torch.nn.RMSNorm of the row fused in front (dia/layers.py:541,560,579,714: the norm that precedes
// every projection): xs = split3((x * rsqrt(mean(x^2) + eps)) * w).  One CTA per row.
__global__ void __launch_bounds__(256) split3_norm_rows_kernel(const float* __restrict__ x, const float* __restrict__ w, float eps,
                                                               __nv_bfloat16* __restrict__ xs, int M, int K) {
    __shared__ float part[8];
    const int row = blockIdx.x;
    const float* xr = x + (size_t)row * K;
    float ss = 0.f;
    for (int i = threadIdx.x; i < K; i += 256) { const float v = xr[i]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += part[i];
    const float inv = 1.0f / sqrtf(tot / (float)K + eps);
    const size_t n = (size_t)M * K;
    for (int i = threadIdx.x; i < K; i += 256) {
        const float v = (xr[i] * inv) * w[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const float r1 = v - __bfloat162float(h);
        const __nv_bfloat16 l = __float2bfloat16_rn(r1);
        const float r2 = r1 - __bfloat162float(l);
        const size_t o = (size_t)row * K + i;
        xs[o] = h;
        xs[n + o] = l;
        xs[2 * n + o] = __float2bfloat16_rn(r2);
    }
}

// W [K][N] (fp32 or bf16, N contiguous) -> Wt [N][K] bf16 (K contiguous)
__global__ void transpose_to_bf16_kernel(const void* __restrict__ w, int src_bf16, __nv_bfloat16* __restrict__ wt, int K, int N) {
    __shared__ float tile[32][33];
    const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + r, n = n0 + threadIdx.x;
        float v = 0.f;
        if (k < K && n < N)
            v = src_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(w)[(size_t)k * N + n])
                         : reinterpret_cast<const float*>(w)[(size_t)k * N + n];
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int n = n0 + r, k = k0 + threadIdx.x;
        if (k < K && n < N) wt[(size_t)n * K + k] = __float2bfloat16_rn(tile[threadIdx.x][r]);
    }
}

__global__ void __launch_bounds__(kGemmThreads, 1)
dia_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        float* y, const float* residual, int M, int N, int K) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* a_ring = smem;
    unsigned char* b_ring = smem + kASlots * kTileBytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(b_ring + kBSlots * kBTileBytes);
    uint64_t* a_empty = a_full + kASlots;
    uint64_t* b_full = a_empty + kASlots;
    uint64_t* b_empty = b_full + kBSlots;
    uint64_t* t_full = b_empty + kBSlots;                   // accumulator complete (MMA warp -> epilogue), one per accumulator
    uint64_t* t_empty = t_full + 2;                         // accumulator drained (epilogue warps -> MMA warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kb = K / BK;
    // tile t = (column block t / tiles_m, row block t % tiles_m): the CTAs of a wave share a few weight blocks and all of A
    const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN, n_tiles = tiles_m * tiles_n;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kASlots; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < kBSlots; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
        fence_mbar_init();
    }
    if (warp == 1) {                                        // one warp allocates (and later frees) the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            unsigned ia = 0, ib = 0;                        // ring positions, running across the tiles
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int n0 = (t / tiles_m) * BN, m0 = (t % tiles_m) * BM;
                for (int kb = 0; kb < n_kb; ++kb) {
                    {
                        const unsigned s = ib % kBSlots;
                        mbar_wait_bounded(&b_empty[s], ((ib / kBSlots) & 1u) ^ 1u);
                        mbar_arrive_expect_tx(&b_full[s], kBTileBytes);
                        tma_load_2d(smem_u32(b_ring + s * kBTileBytes), &map_b, smem_u32(&b_full[s]), kb * BK, n0);
                        ++ib;
                    }
                    for (int tm = 0; tm < kTerms; ++tm) {   // rows of term tm start at tm * M in the stacked matrix
                        const unsigned s = ia % kASlots;
                        mbar_wait_bounded(&a_empty[s], ((ia / kASlots) & 1u) ^ 1u);
                        mbar_arrive_expect_tx(&a_full[s], kTileBytes);
                        tma_load_2d(smem_u32(a_ring + s * kTileBytes), &map_a, smem_u32(&a_full[s]), kb * BK, tm * M + m0);
                        ++ia;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the k-blocks (uniform operands), one elected lane issues =====
        // instruction descriptor: fp32 accumulate, bf16 x bf16, both operands K-major, N = 256, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        unsigned ia = 0, ib = 0, it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const unsigned acc = it & 1u;
            mbar_wait_bounded(&t_empty[acc], ((it >> 1) & 1u) ^ 1u);      // the epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d = tmem_base + acc * BN;
            for (int kb = 0; kb < n_kb; ++kb) {
                const unsigned sb = ib % kBSlots;
                mbar_wait_bounded(&b_full[sb], (ib / kBSlots) & 1u);
                const uint64_t bdesc = umma_desc(smem_u32(b_ring + sb * kBTileBytes));
#pragma unroll
                for (int tm = 0; tm < kTerms; ++tm) {
                    const unsigned sa = ia % kASlots;
                    mbar_wait_bounded(&a_full[sa], (ia / kASlots) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t adesc = umma_desc(smem_u32(a_ring + sa * kTileBytes));
                    if (elect_one_sync()) {
#pragma unroll
                        for (int j = 0; j < BK / 16; ++j)       // advance 16 elements = 32 bytes = 2 descriptor units along K
                            umma_bf16(d, adesc + 2 * j, bdesc + 2 * j, idesc, (kb | tm | j) != 0 ? 1u : 0u);
                        umma_commit(&a_empty[sa]);              // the slot is free once these MMAs have read it
                        if (tm == kTerms - 1) umma_commit(&b_empty[sb]);
                    }
                    __syncwarp();
                    ++ia;
                }
                ++ib;
            }
            if (elect_one_sync()) umma_commit(&t_full[acc]);    // the accumulator is complete
            __syncwarp();
        }
    } else {
        // ===== epilogue: a warp may touch the 32 TMEM lanes of its quarter (warp id mod 4) =====
        const int q = warp & 3;
        unsigned it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int n0 = (t / tiles_m) * BN, m0 = (t % tiles_m) * BM;
            const unsigned acc = it & 1u;
            mbar_wait_bounded(&t_full[acc], (it >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = m0 + q * 32 + lane;
#pragma unroll 1
            for (int cb = 0; cb < BN / 32; ++cb) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + cb * 32;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < M) {
                    // residual add of the layer (dia/layers.py:555,574,582) fused: y may alias residual (each element is
                    // read and written by the same thread).  Column tail: N need not be a multiple of the tile (logits head).
                    const int c0 = n0 + cb * 32;
                    float* dst = y + (size_t)row * N + c0;
                    const float* res = residual ? residual + (size_t)row * N + c0 : nullptr;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (c0 + 4 * i + 4 <= N) {
                            float4 a = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                   __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                            if (res) {
                                const float4 r4 = *reinterpret_cast<const float4*>(res + 4 * i);
                                a.x += r4.x; a.y += r4.y; a.z += r4.z; a.w += r4.w;
                            }
                            *reinterpret_cast<float4*>(dst + 4 * i) = a;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[acc]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN));
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// a [rows][K] bf16 matrix, K contiguous, read in [box_rows x 64 k] boxes with the 128-byte swizzle
static bool make_map(CUtensorMap* map, const void* base, long long rows, int K, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t gemm_workspace_bytes(int M, int K) { return (size_t)kTerms * M * K * 2; }

cudaError_t launch_transpose_to_bf16(const void* w, int src_bf16, void* wt, int K, int N, cudaStream_t st) {
    dim3 grid((N + 31) / 32, (K + 31) / 32), block(32, 8);
    transpose_to_bf16_kernel<<<grid, block, 0, st>>>(w, src_bf16, reinterpret_cast<__nv_bfloat16*>(wt), K, N);
    return cudaGetLastError();
}

// returns cudaErrorNotSupported for shapes the tiling does not cover (the caller decides what to do)
cudaError_t launch_gemm_tcgen05(const float* x, const float* norm_w, float eps, const void* wt, const float* residual,
                                float* y, void* workspace, int M, int N, int K, cudaStream_t st) {
    if (M <= 0 || N % 4 || N <= 0 || K % BK || K < BK) return cudaErrorNotSupported;
    static bool attr[64] = {};                         // function attributes are per device
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess) return e0;
    if (dev < 0 || dev >= 64 || !attr[dev]) {
        cudaError_t e = cudaFuncSetAttribute(dia_gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr[dev] = true;
    }
    const long long n = (long long)M * K;
    if (norm_w)
        split3_norm_rows_kernel<<<M, 256, 0, st>>>(x, norm_w, eps, reinterpret_cast<__nv_bfloat16*>(workspace), M, K);
    else
        split3_rows_kernel<<<(int)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, st>>>(
            x, reinterpret_cast<__nv_bfloat16*>(workspace), n);
    CUtensorMap ma, mb;
    if (!make_map(&ma, workspace, (long long)kTerms * M, K, BM) || !make_map(&mb, wt, N, K, BN)) return cudaErrorNotSupported;
    static int n_sm[64] = {};
    if (dev < 0 || dev >= 64 || n_sm[dev] == 0) {
        int n = 0;
        cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64) n_sm[0] = n; else n_sm[dev] = n;
    }
    const int sms = (dev < 0 || dev >= 64) ? n_sm[0] : n_sm[dev];
    const int n_tiles = ((N + BN - 1) / BN) * ((M + BM - 1) / BM);
    dia_gemm_tcgen05_kernel<<<std::min(n_tiles, sms), kGemmThreads, kGemmSmem, st>>>(ma, mb, y, residual, M, N, K);
    return cudaGetLastError();
}

}  // namespace dia
