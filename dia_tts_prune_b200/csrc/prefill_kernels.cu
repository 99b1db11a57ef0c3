// prefill_kernels.cu - everything around the tcgen05 GEMMs of the T > 1 passes (encoder, cross-attention K/V
// precompute, prompt prefill: SURVEY.md 8(f) rank 1) that is not a dense contraction:
//
//   attention_rows_kernel   F.scaled_dot_product_attention call sites (dia/layers.py:329-337) for T > 1: causal GQA
//                           self-attention of the prompt prefill, full / pad-partitioned encoder self-attention,
//                           cross-attention over the valid text keys.  fp32 throughout (these K/V and outputs are
//                           re-read by every later greedy step): flash-style, one CTA per (batch row, query head,
//                           64 queries), K/V tiles of 64 keys staged in padded shared memory, online softmax.
//   rope_rows_kernel        RotaryEmbedding (dia/layers.py:108-173) on projected heads, fused with the layout change
//                           into the [B, H, T, 128] cache (KVCache.prefill / from_kv, dia/state.py:88-109)
//   rmsnorm_rows_kernel     torch.nn.RMSNorm over rows (dia/layers.py:462, encoder final norm)
//   silu_mul_kernel         MlpBlock gate: silu(gate) * up (dia/layers.py:95-101)
//   embed_rows_kernel       nn.Embedding row gather (dia/layers.py:445-447)
//
// Bound: attention by the fp32 FMA pipe (no tensor cores: the operands may not be rounded), the rest by HBM.
#include <algorithm>

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "common.cuh"
#include "engine_internal.h"

namespace dia {

namespace {

constexpr int kAQ = 64;               // queries per CTA
constexpr int kAK = 64;               // keys per tile
constexpr int kQStride = 132;         // padded row stride (floats) of the Q and K tiles: conflict-free float4 reads
constexpr int kPStride = 68;
constexpr int kAttnThreads = 256;
constexpr int kAttnSmem = (kAQ * kQStride + kAK * kQStride + kAK * kHeadDim + kAQ * kPStride) * 4;

}  // namespace

// mode 0: causal - query t attends keys [0, t]                               (prompt prefill, dia/layers.py:722-766)
// mode 1: partition - query t < n attends keys [0, n), t >= n keys [n, Tk)   (encoder mask, dia/state.py:24-31)
// mode 2: prefix - every query attends keys [0, n); n == 0 gives exact zeros (cross-attention)
// n = n_valid[b].  q / out: [B][Tq][Hq][128]; k / v: [B][Hkv][Tk_stride][128]; query head h reads kv head h / (Hq / Hkv).
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_rows_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      float* __restrict__ out, int Tq, int Tk, int Hq, int Hkv, int Tk_stride, int mode,
                      DelayArg n_valid) {
    extern __shared__ __align__(16) float smem_f[];
    float* Qs = smem_f;
    float* Ks = Qs + kAQ * kQStride;
    float* Vs = Ks + kAK * kQStride;
    float* Ps = Vs + kAK * kHeadDim;

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int q0 = blockIdx.x * kAQ, h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / (Hq / Hkv);
    const int n = n_valid.d[b];
    const float scale = 0.08838834764831845f;                 // 1/sqrt(128)
    const float* kb = k + ((size_t)(b * Hkv + kvh) * Tk_stride) * kHeadDim;
    const float* vb = v + ((size_t)(b * Hkv + kvh) * Tk_stride) * kHeadDim;

    // ---- the query tile (pre-scaled) ----
    for (int i = tid; i < kAQ * 32; i += kAttnThreads) {
        const int r = i >> 5, c4 = i & 31, t = q0 + r;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < Tq) {
            val = *reinterpret_cast<const float4*>(q + (((size_t)b * Tq + t) * Hq + h) * kHeadDim + c4 * 4);
            val.x *= scale; val.y *= scale; val.z *= scale; val.w *= scale;
        }
        *reinterpret_cast<float4*>(Qs + r * kQStride + c4 * 4) = val;
    }

    float m_run[4], l_run[4], o[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m_run[i] = -CUDART_INF_F; l_run[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[i][j] = 0.f;
    }

    // key range this query tile can see
    int k_begin = 0, k_end = Tk;
    if (mode == 0) k_end = min(Tk, q0 + kAQ);
    else if (mode == 2) k_end = min(Tk, n);
    else if (mode == 1) {
        if (q0 + kAQ <= n) k_end = min(Tk, n);               // all queries of the tile are valid text
        else if (q0 >= n) k_begin = (n / kAK) * kAK;         // all queries are padding
    }

    for (int kt0 = k_begin; kt0 < k_end; kt0 += kAK) {
        __syncthreads();                                      // the previous tile is consumed (and Qs is written)
        for (int i = tid; i < kAK * 32; i += kAttnThreads) {
            const int r = i >> 5, c4 = i & 31, key = kt0 + r;
            float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
            if (key < k_end) {
                kk = *reinterpret_cast<const float4*>(kb + (size_t)key * kHeadDim + c4 * 4);
                vv = *reinterpret_cast<const float4*>(vb + (size_t)key * kHeadDim + c4 * 4);
            }
            *reinterpret_cast<float4*>(Ks + r * kQStride + c4 * 4) = kk;
            *reinterpret_cast<float4*>(Vs + r * kHeadDim + c4 * 4) = vv;
        }
        __syncthreads();
        // ---- S = Q K^T: rows ty*4 + i, columns tx + 16 j ----
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
        for (int d4 = 0; d4 < 32; ++d4) {
            float4 qa[4], ka[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qa[i] = *reinterpret_cast<const float4*>(Qs + (ty * 4 + i) * kQStride + d4 * 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) ka[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * kQStride + d4 * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s[i][j] = fmaf(qa[i].x, ka[j].x, s[i][j]);
                    s[i][j] = fmaf(qa[i].y, ka[j].y, s[i][j]);
                    s[i][j] = fmaf(qa[i].z, ka[j].z, s[i][j]);
                    s[i][j] = fmaf(qa[i].w, ka[j].w, s[i][j]);
                }
        }
        // ---- mask, online softmax (row statistics across the 16 threads that share a row) ----
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = q0 + ty * 4 + i;
            float mx = -CUDART_INF_F;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key = kt0 + tx + 16 * j;
                bool ok = key < k_end && key >= k_begin;
                if (mode == 0) ok = ok && key <= t;
                else if (mode == 1) ok = ok && ((t < n) == (key < n));
                if (!ok) s[i][j] = -CUDART_INF_F;
                mx = fmaxf(mx, s[i][j]);
            }
#pragma unroll
            for (int w = 8; w >= 1; w >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, w));
            const float m_new = fmaxf(m_run[i], mx);
            float corr = 1.f, psum = 0.f;
            if (m_new == -CUDART_INF_F) {
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
            } else {
                corr = expf(m_run[i] - m_new);               // exp(-inf) = 0 on the first live tile
#pragma unroll
                for (int j = 0; j < 4; ++j) { s[i][j] = expf(s[i][j] - m_new); psum += s[i][j]; }
            }
#pragma unroll
            for (int w = 8; w >= 1; w >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, w);
            l_run[i] = l_run[i] * corr + psum;
            m_run[i] = m_new;
#pragma unroll
            for (int j = 0; j < 8; ++j) o[i][j] *= corr;
#pragma unroll
            for (int j = 0; j < 4; ++j) Ps[(ty * 4 + i) * kPStride + tx + 16 * j] = s[i][j];
        }
        __syncthreads();
        // ---- O += P V: rows ty*4 + i, output dims tx*4 .. +3 and 64 + tx*4 .. +3 ----
#pragma unroll 2
        for (int c4 = 0; c4 < kAK / 4; ++c4) {
            float4 pa[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pa[i] = *reinterpret_cast<const float4*>(Ps + (ty * 4 + i) * kPStride + c4 * 4);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const float4 v0 = *reinterpret_cast<const float4*>(Vs + (c4 * 4 + cc) * kHeadDim + tx * 4);
                const float4 v1 = *reinterpret_cast<const float4*>(Vs + (c4 * 4 + cc) * kHeadDim + 64 + tx * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float pv = cc == 0 ? pa[i].x : cc == 1 ? pa[i].y : cc == 2 ? pa[i].z : pa[i].w;
                    o[i][0] = fmaf(pv, v0.x, o[i][0]); o[i][1] = fmaf(pv, v0.y, o[i][1]);
                    o[i][2] = fmaf(pv, v0.z, o[i][2]); o[i][3] = fmaf(pv, v0.w, o[i][3]);
                    o[i][4] = fmaf(pv, v1.x, o[i][4]); o[i][5] = fmaf(pv, v1.y, o[i][5]);
                    o[i][6] = fmaf(pv, v1.z, o[i][6]); o[i][7] = fmaf(pv, v1.w, o[i][7]);
                }
            }
        }
    }
    // ---- normalise and store; a query without any allowed key gives exact zeros ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = q0 + ty * 4 + i;
        if (t >= Tq) continue;
        const float inv = l_run[i] > 0.f ? 1.0f / l_run[i] : 0.f;
        float* dst = out + (((size_t)b * Tq + t) * Hq + h) * kHeadDim;
        *reinterpret_cast<float4*>(dst + tx * 4) = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
        *reinterpret_cast<float4*>(dst + 64 + tx * 4) = make_float4(o[i][4] * inv, o[i][5] * inv, o[i][6] * inv, o[i][7] * inv);
    }
}

// src [B*T][H*128]; dst: in the same layout (to_cache == 0) or [B][H][dst_T][128] at slot dst_t0 + t.
// rotate: out[:64] = a cos - b sin, out[64:] = a sin + b cos with (sin, cos) = table[pos[row]] (dia/layers.py:161-173)
__global__ void rope_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, const float* __restrict__ sin_tab,
                                 const float* __restrict__ cos_tab, const int* __restrict__ pos, int B, int T, int H,
                                 int rotate, int to_cache, int dst_T, int dst_t0, int n_pos) {
    const long long total = (long long)B * T * H * 64;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i & 63);
        const long long rh = i >> 6;
        const int hh = (int)(rh % H);
        const long long row = rh / H;
        const int t = (int)(row % T), b = (int)(row / T);
        const float a = src[(row * H + hh) * kHeadDim + d], bb = src[(row * H + hh) * kHeadDim + d + 64];
        float o0 = a, o1 = bb;
        if (rotate) {
            const int p = min(max(pos[row], 0), n_pos - 1);
            const float sn = sin_tab[(size_t)p * 64 + d], cs = cos_tab[(size_t)p * 64 + d];
            o0 = __fsub_rn(__fmul_rn(a, cs), __fmul_rn(bb, sn));      // separate roundings, like the reference's
            o1 = __fadd_rn(__fmul_rn(a, sn), __fmul_rn(bb, cs));      // x1 * cos - x2 * sin (no fused multiply-add)
        }
        float* o = to_cache ? dst + (((size_t)b * H + hh) * dst_T + dst_t0 + t) * kHeadDim
                            : dst + (row * H + hh) * kHeadDim;
        o[d] = o0;
        o[d + 64] = o1;
    }
}

// y = x * rsqrt(mean(x^2) + eps) * w, one CTA per row (torch.nn.RMSNorm, fp32)
__global__ void __launch_bounds__(256) rmsnorm_rows_kernel(const float* __restrict__ x, const float* __restrict__ w, float eps,
                                                           float* __restrict__ y, int D) {
    __shared__ float part[8];
    const float* xr = x + (size_t)blockIdx.x * D;
    float ss = 0.f;
    for (int i = threadIdx.x; i < D; i += 256) { const float vv = xr[i]; ss = fmaf(vv, vv, ss); }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += part[i];
    const float inv = 1.0f / sqrtf(tot / (float)D + eps);
    for (int i = threadIdx.x; i < D; i += 256) y[(size_t)blockIdx.x * D + i] = (xr[i] * inv) * w[i];
}

// h[m][f] = silu(gu[m][0][f]) * gu[m][1][f]
__global__ void silu_mul_kernel(const float* __restrict__ gu, float* __restrict__ h, long long M, int F) {
    const long long total = M * F;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / F;
        const int f = (int)(i - m * F);
        const float g = gu[(m * 2) * F + f], u = gu[(m * 2 + 1) * F + f];
        h[i] = (g / (1.0f + expf(-g))) * u;
    }
}

__global__ void embed_rows_kernel(const float* __restrict__ table, const int* __restrict__ ids, float* __restrict__ out,
                                  int n_rows, int vocab, int D) {
    const long long total = (long long)n_rows * (D / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / (D / 4)), c4 = (int)(i % (D / 4));
        const int id = min(max(ids[r], 0), vocab - 1);
        reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(table + (size_t)id * D)[c4];
    }
}

// ---- launchers ------------------------------------------------------------------------------------------------
static int grid_for(long long n, int block) { return (int)std::min<long long>((n + block - 1) / block, 148 * 16); }

cudaError_t launch_attention_rows(const float* q, const float* k, const float* v, float* out, int B, int Tq, int Tk, int Hq,
                                  int Hkv, int Tk_stride, int mode, const int* n_valid_host, cudaStream_t st) {
    static bool attr[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !attr[dev]) {
        e = cudaFuncSetAttribute(attention_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr[dev] = true;
    }
    DelayArg nv{};
    for (int b = 0; b < B; ++b) nv.d[b] = n_valid_host ? n_valid_host[b] : Tk;
    dim3 grid((Tq + kAQ - 1) / kAQ, Hq, B);
    attention_rows_kernel<<<grid, kAttnThreads, kAttnSmem, st>>>(q, k, v, out, Tq, Tk, Hq, Hkv, Tk_stride, mode, nv);
    return cudaGetLastError();
}

cudaError_t launch_rope_rows(const float* src, float* dst, const float* sin_tab, const float* cos_tab, const int* pos, int B,
                             int T, int H, int rotate, int to_cache, int dst_T, int dst_t0, int n_pos, cudaStream_t st) {
    const long long total = (long long)B * T * H * 64;
    rope_rows_kernel<<<grid_for(total, 256), 256, 0, st>>>(src, dst, sin_tab, cos_tab, pos, B, T, H, rotate, to_cache, dst_T,
                                                           dst_t0, n_pos);
    return cudaGetLastError();
}

cudaError_t launch_rmsnorm_rows(const float* x, const float* w, float eps, float* y, int M, int D, cudaStream_t st) {
    rmsnorm_rows_kernel<<<M, 256, 0, st>>>(x, w, eps, y, D);
    return cudaGetLastError();
}

cudaError_t launch_silu_mul(const float* gu, float* h, int M, int F, cudaStream_t st) {
    silu_mul_kernel<<<grid_for((long long)M * F, 256), 256, 0, st>>>(gu, h, M, F);
    return cudaGetLastError();
}

cudaError_t launch_embed_rows(const float* table, const int* ids, float* out, int n_rows, int vocab, int D, cudaStream_t st) {
    embed_rows_kernel<<<grid_for((long long)n_rows * (D / 4), 256), 256, 0, st>>>(table, ids, out, n_rows, vocab, D);
    return cudaGetLastError();
}

}  // namespace dia
