// engine_internal.h - structures shared between the host engine and the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "dia_b200.h"

namespace dia {

// GEMM families inside one decoder layer (+ the logits head); index into CtaTable
enum GemmType : int { G_QKV = 0, G_SO = 1, G_CQ = 2, G_CO = 3, G_WI = 4, G_WO = 5, G_LOGITS = 6, G_COUNT = 7 };

// stage kinds of one decode step
enum StageKind : int {
    S_EMBED = 0,
    S_QKV, S_SATTN, S_SO, S_CQ, S_CATTN, S_CO, S_WI, S_WO,   // 8 per layer
    S_LOGITS, S_SAMPLE
};

// Per-CTA slice of every GEMM: a CTA owns `gc` groups of 8 output columns starting at group `g0`
// and streams its [K x gc*8] bf16 slab (k-major, 16 B per (k, group)) from `stream_base`.
struct CtaTable {
    unsigned long long stream_base;    // byte offset of this CTA's weight stream
    unsigned long long logits_off;     // byte offset of the logits slab inside the stream
    unsigned layer_bytes;              // bytes of one layer in this CTA's stream
    unsigned slab_off[G_COUNT];        // byte offset of each slab inside a layer block
    int g0[G_COUNT];
    int gc[G_COUNT];
};

// Rows of a slab in one ring slot.  Each of the 8 math warps owns the contiguous K/8 rows [w*K/8, (w+1)*K/8) of the
// contraction; a slot holds `r` of them, r a power of two that divides K/8 and fits 8 KB.  Stream position ci of a
// slab is slot j = ci / 8 of warp w = ci % 8, i.e. rows w*K/8 + j*r ..: ring order alternates between the warps
// while every warp walks a contiguous k-range (one fetch of its input words serves up to 256 / r slots).
//
// 2:4 slabs (`sparse`): a slot holds the same r LOGICAL rows as r/2 compressed rows - of every 4 consecutive logical
// rows the (at most) 2 non-zeros per output column, in row order - followed by the mma.sp metadata of its r/32
// 32-row blocks: 64 bytes per (block, 16-column tile).  0.5625 of the dense bytes.
__host__ __device__ inline int gemm_slot_rows(int gc, int K, int sparse = 0) {
    const int sl = K / 8;
    const int row_bytes = sparse ? gc * 8 + ((gc + 1) / 2) * 2 : gc * 16;     // bytes per logical row
    int cap = 8192 / row_bytes;
    if (cap > 256) cap = 256;
    if (cap > sl) cap = sl;
    int r = sparse ? 32 : 16;
    while (r * 2 <= cap && sl % (r * 2) == 0) r *= 2;
    return r;
}
__host__ __device__ inline unsigned gemm_slot_bytes(int gc, int K, int sparse = 0) {
    const unsigned r = (unsigned)gemm_slot_rows(gc, K, sparse);
    return sparse ? (r / 2) * (unsigned)gc * 16u + (r / 32) * (unsigned)((gc + 1) / 2) * 64u : r * (unsigned)gc * 16u;
}
__host__ __device__ inline unsigned long long gemm_slab_bytes(int gc, int K, int sparse = 0) {
    if (gc == 0) return 0ull;
    return (unsigned long long)(K / gemm_slot_rows(gc, K, sparse)) * gemm_slot_bytes(gc, K, sparse);
}
// position (row index inside the slab's stream) of contraction row k  [dense slabs]
__host__ __device__ inline int gemm_row_position(int k, int gc, int K) {
    const int sl = K / 8, r = gemm_slot_rows(gc, K);
    const int w = k / sl, within = k - w * sl, j = within / r, rr = within - j * r;
    return (w + 8 * j) * r + rr;
}
// slot (stream position) and row inside the slot of logical contraction row k  [2:4 slabs]
__host__ __device__ inline void gemm_sparse_position(int k, int gc, int K, int* slot, int* row_in_slot) {
    const int sl = K / 8, r = gemm_slot_rows(gc, K, 1);
    const int w = k / sl, within = k - w * sl, j = within / r;
    *slot = w + 8 * j;
    *row_in_slot = within - j * r;
}

struct GenState {                      // device-resident Dia.generate loop state (dia/model.py:736-807)
    int dec_step;
    int finished;
    int eos_detected;
    int eos_countdown;
    int bos_countdown;
    int steps_run;
    int device_error;
    int reserved;
};

struct StepParams {
    // geometry
    int L, D, F, Hq, Hkv, Hc, C, V, Vpad, Lmax, Smax;
    int Kdim[G_COUNT];                 // contraction length per GEMM family
    int tclass[G_COUNT];               // MMA tiles every CTA computes per k-block: 1, 2, 4 or 8 (>= its real tiles)
    float eps;
    int G;                             // CTAs the tables were built for
    int n_res;                         // CTAs [0, n_res) own columns of the residual stream (and publish sum(x^2))
    int sa_nsplit, ca_nsplit;          // max key splits per (row, kv head) / per cross head
    int sparse24;                      // weight slabs are 2:4-compressed (mma.sp), see gemm_slot_rows
    // K-row compaction: rowmap[gt] (device, [L][Kfull[gt]], logits [Kfull]) = position of input element k in the compacted
    // contraction of GEMM family gt, or -1 (dropped); nullptr = identity.  Producers write their words there.
    const int* rowmap[G_COUNT];
    int Kfull[G_COUNT];
    // weights
    const unsigned char* wstream;
    const CtaTable* cta_tab;
    const float* emb;                  // [C][V][D] fp32
    const float* norms;                // [L][3][D] then final [D]
    const float* rope_sin;             // [n_pos][64]
    const float* rope_cos;
    int n_pos;
    // caches (device arrays of L pointers)
    float* const* self_k;
    float* const* self_v;
    const float* const* cross_k;
    const float* const* cross_v;
    int text_len;
    // plain buffers at the launch boundary
    float2* x;                         // residual stream, interleaved [D][2] (read when stage_begin > 0, written if want_x)
    int want_x;                        // keep a plain copy of the residual stream (layer-wise API and debug launches)
    float* logits;                     // [2][C][V]
    // flag-in-data buffers (zeroed before every launch; see step_kernel.cu).  Activation vectors feeding a GEMM
    // are stored as [k/16][2 rows][16] words of (bf16 hi, lo, lo2, flag16); the rest as (fp32, flag32).
    unsigned long long* ll_x;          // RMSNorm-weighted residual stream for the next projection   [D]
    unsigned long long* ll_attn;       // self-attention output                                     [Hq*128]
    unsigned long long* ll_cattn;      // cross-attention output (row 0 = exact zeros)             [Hc*128]
    unsigned long long* ll_hidden;     // silu(gate) * up                                           [F]
    unsigned long long* ll_qkv;        // [(Hq + 2 Hkv) * 128][2]
    unsigned long long* ll_cq;         // [Hc * 128][2]
    unsigned long long* ll_ssq;        // [G][2] per-CTA partial sums of x^2 (for the consumer's RMSNorm)
    unsigned long long* ll_sa_part;    // [2*Hkv][sa_nsplit][4][132] split-KV partials (m, l, -, -, o[128])
    unsigned long long* ll_ca_part;    // [Hc][ca_nsplit][132]
    unsigned long long* ll_glog;       // [C][V] guided + masked logits
    unsigned long long* ll_pred;       // [C] raw prediction per channel
    unsigned long long* ll_tok;        // [C] input tokens of the next step
    int* err;
    // run control
    int stage_begin, stage_end;        // stages of a step to execute, [begin, end)
    int n_steps;
    int pos0, slot0;                   // RoPE position / cache slot of the first step
    const int* tokens;                 // [2][C] explicit input tokens (API mode) or nullptr (generate mode)
    // generate mode
    int* grid;                         // [Lmax][C] token grid
    GenState* gs;
    float cfg_scale, temperature, top_p;
    int top_k;
    int max_tokens;
    unsigned long long seed;
    unsigned long long draw0;          // RNG draw index of the first step of this launch
    int eos, pad, bos;
    int delay[DIA_B200_MAX_CHANNELS];
    int* pred_out;                     // [C] raw prediction of the last executed step
    int timing_cta;                    // the CTA whose thread 0 writes `timing`
    long long* timing;                 // optional [n_steps][S][8] clock64 stamps of CTA 0 (see tools/stage_profile.py)
    unsigned long long* cta_timing;    // optional [S][G] globaltimer at the end of each stage of step 1, per CTA
};

// ---- batched engine (N utterances per GPU, batch_kernel.cu) ---------------------------------------------------------
constexpr int kMaxUtt = 8;             // utterances per launch: 16 batch rows = N of the tcgen05 MMA
// activation buffers of the batched kernel (index into BatchParams::ctr)
enum ActBuf : int { A_XQ = 0, A_XC, A_XM, A_XL, A_ATTN, A_CATTN, A_HIDDEN, A_COUNT };
constexpr unsigned kArrivalsPerCta = 2;   // per generation and CTA (the embedding stage arrives from two warps)

// Batched weight stream: per CTA and GEMM the slab [K][gc*8 columns] is stored as K / 64 chunks, each the tile
// tcgen05.mma reads as its M operand: [gc*8 rows = output columns][64 k] bf16, K-major, rows of 128 bytes whose 16-byte
// units are XOR-swizzled with (row % 8) (SWIZZLE_128B).  A ring slot (16 KB) holds `bslot_chunks` consecutive chunks.
__host__ __device__ inline int bchunk_bytes(int gc) { return gc * 8 * 128; }
// Four warps issue the MMAs of a CTA, each on its own quarter of the contraction (kBIssuers); the stream interleaves their
// slots - stream slot 4 j + w is the j-th slot of issuer w - so that every ring slot has exactly one consumer.
constexpr int kBIssuers = 4;
__host__ __device__ inline int bslot_chunks(int gc, int K) {
    const int q = (K / 64) / kBIssuers;                 // chunks per issuer
    int n = 1;
    while (n * 2 * bchunk_bytes(gc) <= 16384 && n * 2 <= 8 && q % (n * 2) == 0) n *= 2;
    return n;
}
// position (in chunks) of k-chunk c of a slab inside the CTA's stream
__host__ __device__ inline int bchunk_position(int c, int gc, int K) {
    const int q = (K / 64) / kBIssuers, cps = bslot_chunks(gc, K);
    const int w = c / q, within = c - w * q, j = within / cps, r = within - j * cps;
    return (kBIssuers * j + w) * cps + r;
}
__host__ __device__ inline unsigned long long bslab_bytes(int gc, int K) {
    return (unsigned long long)(K / 64) * (unsigned long long)bchunk_bytes(gc);
}

struct UttParams {                     // one utterance of a batched launch
    int text_len;                      // valid keys of the conditional row of its cross caches
    int pos0, slot0;                   // RoPE position / self-cache slot of the first step of the launch
    int reserved;
    int* grid;                         // [Lmax][C] token grid (generate mode) or nullptr
    GenState* gs;
    unsigned long long seed;
};

struct BatchParams {
    int L, D, F, Hq, Hkv, Hc, C, V, Vpad, Lmax, Smax;
    int Kdim[G_COUNT];
    float eps;
    int G, U, R;                       // CTAs, utterances, batch rows (2 U: row 2u = unconditional, 2u + 1 = conditional)
    int mc;                            // 1: launched as clusters of 2 CTAs that share every activation stage (multicast halves)
    int sa_nsplit, ca_nsplit;
    const unsigned char* wstream;
    const CtaTable* cta_tab;
    const float* emb;
    const float* norms;
    const float* rope_sin;
    const float* rope_cos;
    int n_pos;
    float* const* self_k;              // device arrays of U * L pointers ([u * L + layer])
    float* const* self_v;
    const float* const* cross_k;
    const float* const* cross_v;
    UttParams utt[kMaxUtt];
    float* logits;                     // [R][C][V] raw logits of the last executed step (optional)
    // exchange buffers (zeroed before every launch).  act_*: the input vectors of the GEMM stages, stored as the very
    // shared-memory image tcgen05.mma reads as its N operand - per 64-k chunk a hi tile and a lo tile (x ~ hi + lo, both
    // bf16) of [16 rows][64 k], K-major, 128-byte rows, 16-byte units XOR-swizzled with (row % 8): 4 KB per chunk, so a
    // consumer stages 128 k of all rows with ONE 8 KB bulk copy.  A buffer is complete when its arrival counter
    // (`ctr`, one per buffer, every CTA adds kArrivalsPerCta per generation) has reached the generation's total.
    unsigned char* act_xq;             // normed residual stream for qkv / cross-q / mlp-in / the logits head: one buffer
    unsigned char* act_xc;             // per consumer (a buffer is rewritten only after every reader has passed a later
    unsigned char* act_xm;             // all-to-all stage)
    unsigned char* act_xl;
    unsigned char* act_attn;
    unsigned char* act_cattn;
    unsigned char* act_hidden;
    unsigned int* ctr;                 // [A_COUNT][32] arrival counters (one 128-byte line each)
    float* ssq;                        // [D / 8][16] sum(x^2) of every 8-column group of the residual stream, per row
    unsigned long long* ll_qkv;        // [(Hq + 2 Hkv) * 128][R]   (fp32, flag32) words as in the single-utterance kernel
    unsigned long long* ll_cq;         // [Hc * 128][R]
    unsigned long long* ll_sa_part;
    unsigned long long* ll_ca_part;
    unsigned long long* ll_glog;       // [U][C][V]
    unsigned long long* ll_pred;       // [kMaxUtt][16]
    unsigned long long* ll_tok;        // [kMaxUtt][16]
    int* err;
    int n_steps;
    int with_sample;                   // 1: embed .. sample (generate mode); 0: embed .. logits (operator boundary)
    const int* tokens;                 // [U][C] input tokens of the first step (operator boundary) or nullptr
    float cfg_scale, temperature, top_p;
    int top_k, max_tokens;
    unsigned long long draw0;
    int eos, pad, bos;
    int delay[DIA_B200_MAX_CHANNELS];
    int* pred_out;                     // [kMaxUtt][16]
    unsigned long long* prof;          // optional [32]: SM-clock totals of CTA 0 (MMA thread / math thread 0), see batch_kernel.cu
};

cudaError_t launch_batch_kernel(const BatchParams& p, cudaStream_t st);
size_t batch_ll_layout(const BatchParams& geom, BatchParams* out, unsigned char* base);

// ---- launchers (each returns the cudaError_t of the launch) --------------------------------
cudaError_t launch_step_kernel(const StepParams& p, bool cooperative, cudaStream_t st);
int step_kernel_smem_bytes();
cudaError_t launch_head_sample(const StepParams& p, const float* logits, unsigned long long draw, int* pred,
                               float* probs, cudaStream_t st);
size_t ll_layout(const StepParams& geom, StepParams* out, unsigned long long* base);   // carve the LL region
struct RepackArgs {
    const void* src[3];     // QKV: q, k, v kernels; otherwise src[0]
    int src_bf16;
    int K;
    int gemm;
    int layer;
    int n_groups;
    int Hq, Hkv, F, V, Vpad, C, N;
    const int* owner;       // [n_groups] CTA owning each column group
    const int* local;       // [n_groups] index of the group inside its CTA's slab
    const CtaTable* tab;
    unsigned char* wstream;
    int batch_format;       // write the batched engine's K-major swizzled chunks (bchunk_bytes) instead of [k][gc][8] slabs
    int sparse;             // write 2:4-compressed slabs (the stream must be zeroed first: metadata is OR-ed in)
    int* violations;        // sparse: counts 4-row groups with more than 2 non-zeros (the model is not 2:4)
};
cudaError_t launch_repack(const RepackArgs& a, cudaStream_t st);
cudaError_t launch_embed_sum(const float* emb, const int* tokens, int n_rows, int C, int V, int D, float* x,
                             cudaStream_t st);
cudaError_t launch_interleave(const float* x_rows, float2* x_il, int D, cudaStream_t st);
cudaError_t launch_deinterleave(const float2* x_il, float* x_rows, int D, cudaStream_t st);
cudaError_t launch_delay_apply(const int* in, int* out, int B, int T, int C, const int* delay, int pad, int bos,
                               cudaStream_t st);
cudaError_t launch_delay_revert(const int* in, int* out, int B, int T, int C, const int* delay, int pad, int T_orig,
                                cudaStream_t st);
cudaError_t launch_finalize_codes(const int* in, int* out, int T, int C, const int* delay, int pad, int codebook,
                                  cudaStream_t st);
cudaError_t launch_build_delay_indices(int* t_idx, long long* idx, int B, int T, int C, const int* delay,
                                       cudaStream_t st);
cudaError_t launch_build_revert_indices(long long* t_idx, long long* idx, int B, int T, int C, const int* delay,
                                        cudaStream_t st);

// tcgen05 GEMM for the T > 1 dense layers (gemm_tcgen05.cu)
size_t gemm_workspace_bytes(int M, int K);
cudaError_t launch_transpose_to_bf16(const void* w, int src_bf16, void* wt, int K, int N, cudaStream_t st);
// y = (residual ? residual : 0) + (norm_w ? rmsnorm(x; norm_w, eps) : x) . W        (y may alias residual)
cudaError_t launch_gemm_tcgen05(const float* x, const float* norm_w, float eps, const void* wt, const float* residual,
                                float* y, void* workspace, int M, int N, int K, cudaStream_t st);

// the rest of the T > 1 passes (prefill_kernels.cu)
cudaError_t launch_attention_rows(const float* q, const float* k, const float* v, float* out, int B, int Tq, int Tk, int Hq,
                                  int Hkv, int Tk_stride, int mode, const int* n_valid_host, cudaStream_t st);
cudaError_t launch_rope_rows(const float* src, float* dst, const float* sin_tab, const float* cos_tab, const int* pos, int B,
                             int T, int H, int rotate, int to_cache, int dst_T, int dst_t0, int n_pos, cudaStream_t st);
cudaError_t launch_rmsnorm_rows(const float* x, const float* w, float eps, float* y, int M, int D, cudaStream_t st);
cudaError_t launch_silu_mul(const float* gu, float* h, int M, int F, cudaStream_t st);
cudaError_t launch_embed_rows(const float* table, const int* ids, float* out, int n_rows, int vocab, int D, cudaStream_t st);

struct DelayArg { int d[DIA_B200_MAX_CHANNELS]; };

}  // namespace dia
