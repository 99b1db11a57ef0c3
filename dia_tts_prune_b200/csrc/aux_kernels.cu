// aux_kernels.cu - one-time / per-utterance kernels around the step kernel:
// weight repack into the per-CTA bf16 stream, standalone embedding gather-sum, layout
// conversion at the API boundary, and the codebook delay / revert integer gathers.
#include <cuda_bf16.h>

#include "common.cuh"
#include "engine_internal.h"

namespace dia {

// ------------------------------------------------------------------------------------------
// weight repack: source DenseGeneral kernels [K][N] (N contiguous, dia/layers.py:47-53) ->
// per-CTA slabs [K][gc][8] bf16.  One thread writes one 16-byte (k, group) unit.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float load_src(const void* src, int bf16, size_t idx) {
    if (bf16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[idx]);
    return reinterpret_cast<const float*>(src)[idx];
}

__global__ void repack_dense_kernel(const RepackArgs a) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (g >= a.n_groups) return;
    const int cta = a.owner[g], gl = a.local[g];
    const CtaTable& t = a.tab[cta];
    const unsigned long long slab = t.stream_base +
        (a.gemm == G_LOGITS ? t.logits_off : (unsigned long long)a.layer * t.layer_bytes + t.slab_off[a.gemm]);
    unsigned char* dst = a.wstream + slab + ((size_t)gemm_row_position(k, t.gc[a.gemm], a.K) * t.gc[a.gemm] + gl) * 16;

    float v[8];
    if (a.gemm == G_QKV) {
        const int n = g * 8, nq = a.Hq * kHeadDim, nk = a.Hkv * kHeadDim;
        const void* s; int col, width;
        if (n < nq) { s = a.src[0]; col = n; width = nq; }
        else if (n < nq + nk) { s = a.src[1]; col = n - nq; width = nk; }
        else { s = a.src[2]; col = n - nq - nk; width = nk; }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = load_src(s, a.src_bf16, (size_t)k * width + col + i);
    } else if (a.gemm == G_WI) {
        // a group = the gate columns (0..3) and the up columns (4..7) of the same 4 hidden units; source is
        // [K][2][F] (dia/layers.py:77-82)
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = load_src(a.src[0], a.src_bf16, ((size_t)k * 2 + (i >> 2)) * a.F + g * 4 + (i & 3));
    } else if (a.gemm == G_LOGITS) {
        // every channel's 1028 columns are padded to Vpad (a multiple of 8) with zero weights
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = g * 8 + i, ch = n / a.Vpad, vv = n - ch * a.Vpad;
            v[i] = (ch < a.C && vv < a.V) ? load_src(a.src[0], a.src_bf16, ((size_t)k * a.C + ch) * a.V + vv) : 0.f;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = load_src(a.src[0], a.src_bf16, (size_t)k * a.N + g * 8 + i);
    }
    uint4 out;
    __nv_bfloat162 b0 = __floats2bfloat162_rn(v[0], v[1]), b1 = __floats2bfloat162_rn(v[2], v[3]),
                   b2 = __floats2bfloat162_rn(v[4], v[5]), b3 = __floats2bfloat162_rn(v[6], v[7]);
    out.x = *reinterpret_cast<uint32_t*>(&b0); out.y = *reinterpret_cast<uint32_t*>(&b1);
    out.z = *reinterpret_cast<uint32_t*>(&b2); out.w = *reinterpret_cast<uint32_t*>(&b3);
    *reinterpret_cast<uint4*>(dst) = out;
}

// source value of contraction row k, column i of 8-column group g (same addressing as repack_dense_kernel)
__device__ __forceinline__ float repack_src(const RepackArgs& a, int g, int k, int i) {
    if (a.gemm == G_QKV) {
        const int n = g * 8, nq = a.Hq * kHeadDim, nk = a.Hkv * kHeadDim;
        if (n < nq) return load_src(a.src[0], a.src_bf16, (size_t)k * nq + n + i);
        if (n < nq + nk) return load_src(a.src[1], a.src_bf16, (size_t)k * nk + n - nq + i);
        return load_src(a.src[2], a.src_bf16, (size_t)k * nk + n - nq - nk + i);
    }
    if (a.gemm == G_WI) return load_src(a.src[0], a.src_bf16, ((size_t)k * 2 + (i >> 2)) * a.F + g * 4 + (i & 3));
    if (a.gemm == G_LOGITS) {
        const int n = g * 8 + i, ch = n / a.Vpad, vv = n - ch * a.Vpad;
        return (ch < a.C && vv < a.V) ? load_src(a.src[0], a.src_bf16, ((size_t)k * a.C + ch) * a.V + vv) : 0.f;
    }
    return load_src(a.src[0], a.src_bf16, (size_t)k * a.N + g * 8 + i);
}

// Batched engine (batch_kernel.cu): the slab as K / 64 chunks of [gc*8 output columns][64 k] bf16, K-major, 128-byte rows
// whose 16-byte units are XOR-swizzled with (row % 8) - the shared-memory image tcgen05.mma reads as its M operand.
// One thread per (8 consecutive contraction rows, 8-column group): eight 16-byte units, one per column.
__global__ void repack_batch_kernel(const RepackArgs a) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int k0 = blockIdx.y * 8;
    if (g >= a.n_groups) return;
    const int cta = a.owner[g], gl = a.local[g];
    const CtaTable& t = a.tab[cta];
    const int gc = t.gc[a.gemm];
    const unsigned long long slab = t.stream_base +
        (a.gemm == G_LOGITS ? t.logits_off : (unsigned long long)a.layer * t.layer_bytes + t.slab_off[a.gemm]);
    unsigned char* chunk = a.wstream + slab + (size_t)bchunk_position(k0 >> 6, gc, a.K) * bchunk_bytes(gc);
    const int unit = (k0 >> 3) & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = repack_src(a, g, k0 + j, i);
        uint4 out;
        __nv_bfloat162 b0 = __floats2bfloat162_rn(v[0], v[1]), b1 = __floats2bfloat162_rn(v[2], v[3]),
                       b2 = __floats2bfloat162_rn(v[4], v[5]), b3 = __floats2bfloat162_rn(v[6], v[7]);
        out.x = *reinterpret_cast<uint32_t*>(&b0); out.y = *reinterpret_cast<uint32_t*>(&b1);
        out.z = *reinterpret_cast<uint32_t*>(&b2); out.w = *reinterpret_cast<uint32_t*>(&b3);
        const int m = gl * 8 + i;                              // row of the chunk = column of the slab
        *reinterpret_cast<uint4*>(chunk + (size_t)m * 128 + ((unit ^ i) << 4)) = out;
    }
}

// 2:4 slabs (engine_internal.h, gemm_slot_rows): one thread per (4 consecutive contraction rows, 8-column group).
// Per column it keeps the (at most two) non-zeros of the four rows in row order - padding with a zero entry of the
// group when there are fewer - writes them to the two compressed rows of the group and ORs the 4-bit index pair into
// the mma.sp metadata block of (16-column tile, 32-row block) at the word / nibble that the hardware reads for it
// (sp_meta_word / sp_meta_shift in common.cuh; layout measured with tools/microbench/mma_sp_probe.cu).
__global__ void repack_sparse24_kernel(const RepackArgs a) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int k0 = blockIdx.y * 4;
    if (g >= a.n_groups) return;
    const int cta = a.owner[g], gl = a.local[g];
    const CtaTable& t = a.tab[cta];
    const int gc = t.gc[a.gemm];
    const unsigned long long slab = t.stream_base +
        (a.gemm == G_LOGITS ? t.logits_off : (unsigned long long)a.layer * t.layer_bytes + t.slab_off[a.gemm]);
    int slot, rr;
    gemm_sparse_position(k0, gc, a.K, &slot, &rr);
    const int r = gemm_slot_rows(gc, a.K, 1);
    unsigned char* base = a.wstream + slab + (size_t)slot * gemm_slot_bytes(gc, a.K, 1);
    const int row_bytes = gc * 16, n_mt = (gc + 1) / 2;
    float e0[8], e1[8];
    uint32_t nib[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v[4];
        int cnt = 0, p0 = -1, p1 = -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = repack_src(a, g, k0 + j, i);
            if (v[j] != 0.f) { if (cnt == 0) p0 = j; else if (cnt == 1) p1 = j; ++cnt; }
        }
        if (cnt > 2) atomicAdd(a.violations, 1);
        if (cnt == 0) { p0 = 0; p1 = 1; }
        else if (cnt == 1) { if (p0 == 3) { p1 = 3; p0 = 0; } else p1 = 3; }      // ascending pair that contains the non-zero
        e0[i] = v[p0]; e1[i] = v[p1];
        nib[i] = (uint32_t)p0 | ((uint32_t)p1 << 2);
    }
    auto pack8 = [](const float (&x)[8]) {
        uint4 out;
        __nv_bfloat162 b0 = __floats2bfloat162_rn(x[0], x[1]), b1 = __floats2bfloat162_rn(x[2], x[3]),
                       b2 = __floats2bfloat162_rn(x[4], x[5]), b3 = __floats2bfloat162_rn(x[6], x[7]);
        out.x = *reinterpret_cast<uint32_t*>(&b0); out.y = *reinterpret_cast<uint32_t*>(&b1);
        out.z = *reinterpret_cast<uint32_t*>(&b2); out.w = *reinterpret_cast<uint32_t*>(&b3);
        return out;
    };
    const int crow = (rr >> 2) * 2;
    *reinterpret_cast<uint4*>(base + (size_t)crow * row_bytes + gl * 16) = pack8(e0);
    *reinterpret_cast<uint4*>(base + (size_t)(crow + 1) * row_bytes + gl * 16) = pack8(e1);
    uint32_t* meta = reinterpret_cast<uint32_t*>(base + (size_t)(r / 2) * row_bytes) +
                     ((size_t)(rr >> 5) * n_mt + (gl >> 1)) * 16;
    const int gq = (rr & 31) >> 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = (gl & 1) * 8 + i;                   // row of the 16 x 32 A tile = output column of the tile
        atomicOr(meta + sp_meta_word(n, gq), nib[i] << sp_meta_shift(n, gq));
    }
}

cudaError_t launch_repack(const RepackArgs& a, cudaStream_t st) {
    if (a.sparse) {
        dim3 grid((a.n_groups + 127) / 128, a.K / 4);
        repack_sparse24_kernel<<<grid, 128, 0, st>>>(a);
        return cudaGetLastError();
    }
    if (a.batch_format) {
        dim3 grid((a.n_groups + 127) / 128, a.K / 8);
        repack_batch_kernel<<<grid, 128, 0, st>>>(a);
        return cudaGetLastError();
    }
    dim3 grid((a.n_groups + 127) / 128, a.K);
    repack_dense_kernel<<<grid, 128, 0, st>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// standalone embedding gather-sum for n_rows token rows (prefill uses it with n_rows = 2*T)
// ------------------------------------------------------------------------------------------
__global__ void embed_sum_kernel(const float* __restrict__ emb, const int* __restrict__ tokens, int n_rows, int C,
                                 int V, int D, float* __restrict__ x, int* err) {
    const int row = blockIdx.y;
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D || row >= n_rows) return;
    float s = 0.f;
    for (int ch = 0; ch < C; ++ch) {
        int t = tokens[(size_t)row * C + ch];
        if (t < 0 || t >= V) { if (err) *err = kErrBadState; t = 0; }
        const float e = __ldg(emb + ((size_t)ch * V + t) * D + d);
        s = ch == 0 ? e : s + e;
    }
    x[(size_t)row * D + d] = s;
}

cudaError_t launch_embed_sum(const float* emb, const int* tokens, int n_rows, int C, int V, int D, float* x,
                             cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    dim3 grid((D + 255) / 256, n_rows);
    embed_sum_kernel<<<grid, 256, 0, st>>>(emb, tokens, n_rows, C, V, D, x, nullptr);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// [2][D] rows <-> interleaved [D][2]
// ------------------------------------------------------------------------------------------
__global__ void interleave_kernel(const float* __restrict__ rows, float2* __restrict__ il, int D) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < D) il[d] = make_float2(rows[d], rows[D + d]);
}
__global__ void deinterleave_kernel(const float2* __restrict__ il, float* __restrict__ rows, int D) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < D) { const float2 v = il[d]; rows[d] = v.x; rows[D + d] = v.y; }
}
cudaError_t launch_interleave(const float* x_rows, float2* x_il, int D, cudaStream_t st) {
    interleave_kernel<<<(D + 255) / 256, 256, 0, st>>>(x_rows, x_il, D);
    return cudaGetLastError();
}
cudaError_t launch_deinterleave(const float2* x_il, float* x_rows, int D, cudaStream_t st) {
    deinterleave_kernel<<<(D + 255) / 256, 256, 0, st>>>(x_il, x_rows, D);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// delay pattern gathers (dia/audio.py).  The reference materialises [B*T*C, 3] int64 index
// tensors and gathers through them; here t - delay[c] is computed in-register and the
// [B][T][C] grid is read and written fully coalesced (consecutive threads = consecutive c, t).
// ------------------------------------------------------------------------------------------
__global__ void delay_apply_kernel(const int* __restrict__ in, int* __restrict__ out, int B, int T, int C,
                                   DelayArg dl, int pad, int bos) {
    const long long n = (long long)B * T * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long bt = i / C;
        const int t = (int)(bt % T);
        const long long b = bt / T;
        const int ti = t - dl.d[c];
        int v;
        if (ti < 0) v = bos;
        else if (ti >= T) v = pad;
        else v = in[(b * T + ti) * C + c];
        out[i] = v;
    }
}
__global__ void delay_revert_kernel(const int* __restrict__ in, int* __restrict__ out, int B, int T, int C,
                                    DelayArg dl, int pad, int T_orig) {
    const long long n = (long long)B * T * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long bt = i / C;
        const int t = (int)(bt % T);
        const long long b = bt / T;
        int ti = t + dl.d[c];
        if (ti > T - 1) ti = T - 1;
        out[i] = ti >= T_orig ? pad : in[(b * T + ti) * C + c];
    }
}
// revert + drop the last max(delay) rows + zero codes outside [0, codebook) + transpose to [C][T']
__global__ void finalize_codes_kernel(const int* __restrict__ in, int* __restrict__ out, int T, int C, DelayArg dl,
                                      int pad, int codebook, int Tout) {
    const int n = C * Tout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = i / Tout, t = i - c * Tout;
        int ti = t + dl.d[c];
        if (ti > T - 1) ti = T - 1;
        int v = ti >= T ? pad : in[(size_t)ti * C + c];
        if (v < 0 || v > codebook - 1) v = 0;
        out[i] = v;
    }
}
__global__ void build_delay_indices_kernel(int* __restrict__ t_idx, long long* __restrict__ idx, int B, int T, int C,
                                           DelayArg dl) {
    const long long n = (long long)B * T * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long bt = i / C;
        const int t = (int)(bt % T);
        const long long b = bt / T;
        const int ti = t - dl.d[c];
        t_idx[i] = ti;
        idx[i * 3 + 0] = b;
        idx[i * 3 + 1] = ti < 0 ? 0 : (ti > T - 1 ? T - 1 : ti);
        idx[i * 3 + 2] = c;
    }
}
__global__ void build_revert_indices_kernel(long long* __restrict__ t_idx, long long* __restrict__ idx, int B, int T,
                                            int C, DelayArg dl) {
    const long long n = (long long)B * T * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long bt = i / C;
        const int t = (int)(bt % T);
        const long long b = bt / T;
        int ti = t + dl.d[c];
        if (ti > T - 1) ti = T - 1;
        t_idx[i] = ti;
        idx[i * 3 + 0] = b;
        idx[i * 3 + 1] = ti;
        idx[i * 3 + 2] = c;
    }
}

static DelayArg make_delay(const int* delay, int C) {
    DelayArg a;
    for (int i = 0; i < DIA_B200_MAX_CHANNELS; ++i) a.d[i] = i < C ? delay[i] : 0;
    return a;
}
static int grid_for(long long n) {
    long long g = (n + 255) / 256;
    if (g < 1) g = 1;
    if (g > 148 * 8) g = 148 * 8;       // grid-stride; a multiple of the SM count once the grid saturates
    return (int)g;
}

cudaError_t launch_delay_apply(const int* in, int* out, int B, int T, int C, const int* delay, int pad, int bos,
                               cudaStream_t st) {
    const long long n = (long long)B * T * C;
    if (n == 0) return cudaSuccess;
    delay_apply_kernel<<<grid_for(n), 256, 0, st>>>(in, out, B, T, C, make_delay(delay, C), pad, bos);
    return cudaGetLastError();
}
cudaError_t launch_delay_revert(const int* in, int* out, int B, int T, int C, const int* delay, int pad, int T_orig,
                                cudaStream_t st) {
    const long long n = (long long)B * T * C;
    if (n == 0) return cudaSuccess;
    delay_revert_kernel<<<grid_for(n), 256, 0, st>>>(in, out, B, T, C, make_delay(delay, C), pad, T_orig);
    return cudaGetLastError();
}
cudaError_t launch_finalize_codes(const int* in, int* out, int T, int C, const int* delay, int pad, int codebook,
                                  cudaStream_t st) {
    int dmax = 0;
    for (int i = 0; i < C; ++i) dmax = delay[i] > dmax ? delay[i] : dmax;
    const int Tout = T - dmax;
    if (Tout <= 0) return cudaSuccess;
    finalize_codes_kernel<<<grid_for((long long)C * Tout), 256, 0, st>>>(in, out, T, C, make_delay(delay, C), pad,
                                                                          codebook, Tout);
    return cudaGetLastError();
}
cudaError_t launch_build_delay_indices(int* t_idx, long long* idx, int B, int T, int C, const int* delay,
                                       cudaStream_t st) {
    const long long n = (long long)B * T * C;
    if (n == 0) return cudaSuccess;
    build_delay_indices_kernel<<<grid_for(n), 256, 0, st>>>(t_idx, idx, B, T, C, make_delay(delay, C));
    return cudaGetLastError();
}
cudaError_t launch_build_revert_indices(long long* t_idx, long long* idx, int B, int T, int C, const int* delay,
                                        cudaStream_t st) {
    const long long n = (long long)B * T * C;
    if (n == 0) return cudaSuccess;
    build_revert_indices_kernel<<<grid_for(n), 256, 0, st>>>(t_idx, idx, B, T, C, make_delay(delay, C));
    return cudaGetLastError();
}

}  // namespace dia
