/* dia_b200.h - C ABI of the B200-native Dia decode path (libdia_b200.so).
 *
 * Plain C types only: device pointers, sizes, an explicit CUDA stream passed as
 * void* (cudaStream_t).  Every entry point returns 0 on success or a negative
 * DIA_B200_E* code; nothing throws across this boundary and the per-step calls
 * neither allocate nor synchronise (they are graph-capturable).  There is no
 * CPU fallback: every compute entry point launches sm_100a kernels.
 *
 * Each entry point names the reference interface it replaces
 * (babybirdprd/dia-tts-prune, paths relative to the reference root).
 *
 * Memory layouts
 *   token ids        int32
 *   logits           float32 [2][C][V]            (reference: [2,1,C,V], dia/layers.py:720)
 *   self  KV cache   float32 [2][kv_heads][max_audio_len][128]  per layer (dia/state.py:83-84)
 *   cross KV cache   float32 [2][cross_heads][max_text_len][128] per layer (dia/layers.py:659-663), contiguous
 *   activations x    float32 [2][d_model]          at the API; interleaved [d_model][2] inside the engine
 *   CFG rows         row 0 = unconditional, row 1 = conditional (dia/model.py:362,450-451)
 */
#ifndef DIA_B200_H
#define DIA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIA_B200_ABI_VERSION 1
#define DIA_B200_MAX_CHANNELS 16
#define DIA_B200_HEAD_DIM 128

enum dia_b200_status {
    DIA_B200_OK = 0,
    DIA_B200_EINVAL = -1,      /* bad argument / unsupported shape                      */
    DIA_B200_ECUDA = -2,       /* a CUDA runtime call failed (see dia_b200_last_cuda_error) */
    DIA_B200_ENOMEM = -3,      /* device or host allocation failed                      */
    DIA_B200_ESTATE = -4,      /* call sequence error (weights / caches not bound ...)   */
    DIA_B200_EUNSUPPORTED = -5 /* valid in the reference, not implemented on this path   */
};

/* Model geometry, filled from DiaConfig (dia/config.py:24-153). head_dim must be 128. */
typedef struct dia_b200_shape {
    int32_t n_layer;        /* model.decoder.n_layer            */
    int32_t d_model;        /* model.decoder.n_embd             */
    int32_t n_hidden;       /* model.decoder.n_hidden           */
    int32_t q_heads;        /* model.decoder.gqa_query_heads    */
    int32_t kv_heads;       /* model.decoder.kv_heads           */
    int32_t cross_heads;    /* model.decoder.cross_query_heads  */
    int32_t channels;       /* data.channels                    */
    int32_t vocab;          /* model.tgt_vocab_size             */
    int32_t max_audio_len;  /* data.audio_length                */
    int32_t max_text_len;   /* data.text_length                 */
    int32_t eos_value, pad_value, bos_value; /* data.audio_{eos,pad,bos}_value */
    int32_t delay_pattern[DIA_B200_MAX_CHANNELS]; /* data.delay_pattern */
    float norm_eps;         /* model.normalization_layer_epsilon */
    int32_t sparse24;       /* 1: every dense kernel is 2:4-sparse along its input axis (at most 2 non-zeros in each 4
                               consecutive K entries of a column - the 2:4 variant of offline_prune.py's checkpoints);
                               the engine streams compressed slabs and multiplies with mma.sp.  0: dense. */
    int32_t k_rows[7];      /* K-row compaction of a structurally pruned checkpoint (offline_prune.py --prune-dim 0 zeroes whole
                               INPUT rows of a kernel, dia/pruning_utils.py:64-119): contraction length of each GEMM family
                               (0 qkv, 1 self-o, 2 cross-q, 3 cross-o, 4 mlp-in, 5 mlp-out, 6 logits) after its all-zero rows
                               were dropped; 0 = not compacted.  The kernels handed to dia_b200_load_decoder_weights then
                               have that many rows, and dia_b200_set_row_map says where every input element goes. */
} dia_b200_shape;

/* Sampling / loop parameters of Dia.generate (dia/model.py:632-647). */
typedef struct dia_b200_gen_params {
    float cfg_scale;        /* default 3.0  */
    float temperature;      /* 0 => argmax (dia/model.py:38-40) */
    float top_p;            /* default 0.95 */
    int32_t top_k;          /* cfg_filter_top_k, default 35; <=0 disables */
    int32_t max_tokens;     /* loop bound, dia/model.py:702,748 */
    int32_t prefill_step;   /* DecoderOutput.prefill_step, dia/state.py:205-208 */
    int32_t first_slot;     /* KVCache.current_idx when the loop starts: 0 without an audio prompt,
                               prefill_step-2 after a prefill (dia/state.py:105-109) */
    int32_t reserved;
    uint64_t seed;          /* Philox key for the multinomial draw */
} dia_b200_gen_params;

/* Host-visible snapshot of the device-side loop state (dia/model.py:736-807). */
typedef struct dia_b200_gen_status {
    int32_t dec_step;       /* value of dec_step when the loop stopped / so far     */
    int32_t finished;       /* 1 once the reference loop would have exited          */
    int32_t eos_detected;
    int32_t eos_countdown;
    int32_t bos_countdown;
    int32_t steps_run;      /* decode steps executed since generate_begin           */
    int32_t device_error;   /* non-zero if a kernel watchdog fired                  */
    int32_t reserved;
} dia_b200_gen_status;

typedef struct dia_b200_engine dia_b200_engine;

/* ---- library ------------------------------------------------------------------------- */
int dia_b200_abi_version(void);
const char *dia_b200_error_string(int code);
const char *dia_b200_last_cuda_error(void);

/* ---- engine lifetime ------------------------------------------------------------------
 * One engine per GPU (per process in the one-process-per-GPU layout).  n_ctas = 0 picks
 * one persistent CTA per SM (148 on B200).  Replaces the implicit state behind
 * Dia.__init__ / DiaModel (dia/model.py:102-137, dia/layers.py:769-807) for the decode path. */
int dia_b200_engine_create(const dia_b200_shape *shape, int device, int n_ctas, dia_b200_engine **out);
int dia_b200_engine_destroy(dia_b200_engine *e);
int dia_b200_engine_num_ctas(const dia_b200_engine *e);
/* bytes of bf16 weights streamed per decode step (the roofline numerator's fixed part) */
int64_t dia_b200_engine_weight_stream_bytes(const dia_b200_engine *e);

/* Repack the decoder parameters into the per-CTA bf16 stream the step kernel reads.
 * `tensors` is a HOST array of DEVICE pointers in this order (state_dict names of
 * dia/layers.py):
 *   [0 .. C)                       decoder.embeddings.c.weight            (V, D)   always float32
 *   then per layer l, 11 tensors:  pre_sa_norm, pre_ca_norm, pre_mlp_norm (D,)     always float32
 *                                  self_attention.{q,k,v,o}_proj.weight
 *                                  cross_attention.{q,o}_proj.weight
 *                                  mlp.wi_fused.weight (D,2,F), mlp.wo.weight (F,D)
 *   then                           decoder.norm.weight (D,) float32, decoder.logits_dense.weight (D,C,V)
 * Dense kernels are [in..., out...] row-major, out-contiguous (DenseGeneral, dia/layers.py:47-53);
 * `dense_dtype` is 0 for float32 sources, 1 for bfloat16.  The sources may be freed afterwards. */
int dia_b200_load_decoder_weights(dia_b200_engine *e, const void *const *tensors, int n_tensors,
                                  int dense_dtype, void *stream);
/* 2:4 / structured pruning support: same call as above with masks already applied as zeros
 * (dia/pruning_utils.py:122-151 makes them permanent) - zeros stream like any other value. */

/* sin/cos of position*inv_freq for positions [0,n_pos), HOST float32 [n_pos][64] each
 * (RotaryEmbedding, dia/layers.py:126-132,161-169; computed by the caller so that it equals
 * the reference's own CPU values bit for bit). */
/* K-row compaction (shape.k_rows[gemm] > 0): map int32 [n_layer][K_full] (host; the logits head: [K_full]) gives, for every
 * element k of the input vector of that GEMM family in every layer, its row in the compacted kernel, or -1 if the kernel's
 * row k was all zeros and has been dropped.  Every row 0 .. k_rows[gemm]-1 must be hit exactly once per layer.  The stage
 * that PRODUCES the vector writes each element straight to its compacted position, so the consumer streams and multiplies
 * only the live rows - exact, because the dropped products are zeros. */
int dia_b200_set_row_map(dia_b200_engine *e, int gemm, const int32_t *map_host, void *stream);

int dia_b200_set_rope_table(dia_b200_engine *e, const float *sin_host, const float *cos_host, int n_pos);

/* Bind one utterance's caches: HOST arrays (n_layer entries) of DEVICE pointers.
 * Replaces DecoderInferenceState.self_attn_cache / cross_attn_cache (dia/state.py:112-162).
 * text_len = number of leading valid (non-pad) text positions of the conditional row; the
 * unconditional row attends nothing and contributes exact zeros. */
int dia_b200_bind_caches(dia_b200_engine *e, void *const *self_k, void *const *self_v,
                         const void *const *cross_k, const void *const *cross_v, int n_layer, int text_len,
                         void *stream);

/* ---- operator boundaries ---------------------------------------------------------------- */

/* Decoder.decode_step (dia/layers.py:671-720): tokens int32 [2][C] (device) -> logits float32
 * [2][C][V] (device); side effect: K/V of this step written at `slot` of every bound self cache
 * (KVCache.update, dia/state.py:99-103); attends slots [0, slot].  pos = RoPE position. */
int dia_b200_decode_step(dia_b200_engine *e, const int32_t *tokens, int pos, int slot, float *logits, void *stream);

/* DecoderLayer.forward at T=1 (dia/layers.py:530-584): x_in/x_out float32 [2][D] (device, may alias). */
int dia_b200_decoder_layer_step(dia_b200_engine *e, int layer, const float *x_in, float *x_out, int pos, int slot,
                                void *stream);

/* 9-codebook embedding gather-sum (dia/layers.py:691-696,737-742): tokens int32 [n_rows][C] ->
 * x float32 [n_rows][D]; sums in ascending channel order like the reference. */
int dia_b200_embed_sum(dia_b200_engine *e, const int32_t *tokens, int n_rows, float *x, void *stream);

/* Dia._decoder_step post-processing + _sample_next_token (dia/model.py:447-488, 32-82) on given
 * logits float32 [2][C][V]: CFG combine, EOS/PAD/BOS masks, argmax or temperature/top-k/top-p/
 * multinomial.  pred int32 [C] (device).  probs (optional, may be NULL) float32 [C][V] receives the
 * filtered probability vector the draw is made from (dia/model.py:73).  `draw` indexes the RNG stream. */
int dia_b200_head_sample(dia_b200_engine *e, const float *logits, float cfg_scale, float temperature, float top_p,
                         int top_k, uint64_t seed, uint64_t draw, int32_t *pred, float *probs, void *stream);

/* The Dia.generate hot loop (dia/model.py:736-815) as a device-resident state machine.
 * grid: int32 [max_audio_len][C] token grid (DecoderOutput.generated_tokens, dia/state.py:172-208),
 * already prefilled by the caller.  generate_steps runs up to n_steps loop iterations with no host
 * round trip (EOS countdown, BOS masking and forced EOS included); iterations after the loop has
 * finished are no-ops on the grid. */
int dia_b200_generate_begin(dia_b200_engine *e, int32_t *grid, const dia_b200_gen_params *p, void *stream);
int dia_b200_generate_steps(dia_b200_engine *e, int n_steps, void *stream);
/* asynchronous copy of the loop state into pinned host memory owned by the engine, then stream sync */
int dia_b200_generate_status(dia_b200_engine *e, dia_b200_gen_status *out, void *stream);

/* Codebook delay pattern (dia/audio.py:6-85) and its inverse (:88-163) as index-free gathers.
 * in/out int32 [B][T][C] (device, must not alias). */
int dia_b200_delay_apply_i32(const int32_t *in, int32_t *out, int B, int T, int C, const int32_t *delay_host,
                             int32_t pad_value, int32_t bos_value, void *stream);
int dia_b200_delay_revert_i32(const int32_t *in, int32_t *out, int B, int T, int C, const int32_t *delay_host,
                              int32_t pad_value, int T_orig, void *stream);
/* Token half of Dia._generate_output (dia/model.py:504-533): revert, drop the last max(delay) rows,
 * zero codes outside [0, codebook_size), transpose: in int32 [T][C] -> out int32 [C][T - max(delay)]. */
int dia_b200_finalize_codes_i32(const int32_t *in, int32_t *out, int T, int C, const int32_t *delay_host,
                                int32_t pad_value, int codebook_size, void *stream);
/* The index tensors build_delay_indices / build_revert_indices return (dia/audio.py:6-41, 88-122):
 * t_idx int32 (apply) or int64 (revert) [B][T][C], indices int64 [B*T*C][3]. */
int dia_b200_build_delay_indices(int32_t *t_idx, int64_t *indices, int B, int T, int C, const int32_t *delay_host,
                                 void *stream);
int dia_b200_build_revert_indices(int64_t *t_idx, int64_t *indices, int B, int T, int C, const int32_t *delay_host,
                                  void *stream);

/* ---- N utterances per GPU in one launch ------------------------------------------------
 * The reference hard-wires the CFG batch of 2 rows (dia/state.py:58-60,83-84,138-139); a batched engine decodes up to
 * DIA_B200_MAX_UTTERANCES independent utterances per launch - 2N batch rows share ONE pass over the weights (tcgen05
 * tensor cores, accumulators in tensor memory).  Every utterance keeps the reference's own state objects: its own
 * KVCache tensors [2][H][L][128] (dia/state.py:72-109), token grid (dia/state.py:172-208), text length, prompt length
 * (prefill_step / first_slot), RNG seed and EOS state machine.  One sampling configuration per launch.
 * dia_b200_load_decoder_weights and dia_b200_set_rope_table serve both kinds of engine. */
#define DIA_B200_MAX_UTTERANCES 8
int dia_b200_engine_create_batched(const dia_b200_shape *shape, int device, int n_ctas, int max_utterances,
                                   dia_b200_engine **out);
int dia_b200_engine_max_utterances(const dia_b200_engine *e);
/* caches of utterance `utterance` (same contract as dia_b200_bind_caches) */
int dia_b200_batch_bind_caches(dia_b200_engine *e, int utterance, void *const *self_k, void *const *self_v,
                               const void *const *cross_k, const void *const *cross_v, int n_layer, int text_len,
                               void *stream);
/* Decoder.decode_step (dia/layers.py:671-720) for n utterances at once: tokens int32 [n][C] (device; both CFG rows of
 * an utterance get the same tokens, dia/model.py:759), pos_host / slot_host int32 [n] (host) -> logits float32
 * [2n][C][V] (device; row 2u unconditional, 2u + 1 conditional of utterance u); appends K/V at slot_host[u]. */
int dia_b200_batch_decode_step(dia_b200_engine *e, int n_utterances, const int32_t *tokens, const int32_t *pos_host,
                               const int32_t *slot_host, float *logits, void *stream);
/* the Dia.generate loop (dia/model.py:736-815) of n utterances in lock step: grids = n device pointers (host array),
 * params = n entries (host array; cfg_scale, temperature, top_p, top_k, max_tokens must agree) */
int dia_b200_batch_generate_begin(dia_b200_engine *e, int n_utterances, int32_t *const *grids,
                                  const dia_b200_gen_params *params, void *stream);
int dia_b200_batch_generate_steps(dia_b200_engine *e, int n_steps, void *stream);
int dia_b200_batch_generate_status(dia_b200_engine *e, dia_b200_gen_status *out /* [n] */, void *stream);

/* ---- dense layers with more than one row (encoder, cross-attention K/V precompute, prompt prefill) ------------
 * DenseGeneral.forward (dia/layers.py:55-66) for M = B*T > 1 rows on the tcgen05 tensor cores:
 *   y[M][N] float32 = x[M][K] float32 . W[K][N]
 * with W given as the K-major bfloat16 copy wt[N][K] that dia_b200_dense_prepare_weight makes once per weight
 * (src_dtype 0 = float32 source, 1 = bfloat16).  x is split into three bf16 terms inside, so the product has
 * fp32-operand accuracy.  workspace: dia_b200_dense_workspace_bytes(M, K) bytes of device memory, 16-byte aligned.
 * Shapes need K % 64 == 0 and N % 4 == 0, otherwise DIA_B200_EUNSUPPORTED. */
int dia_b200_dense_prepare_weight(const void *w, int src_dtype, void *wt_bf16, int K, int N, void *stream);
size_t dia_b200_dense_workspace_bytes(int M, int K);
int dia_b200_dense_forward(const float *x, const void *wt_bf16, float *y, void *workspace, int M, int N, int K,
                           void *stream);
/* The same with the two neighbours of every projection fused in (DecoderLayer / EncoderLayer.forward,
 * dia/layers.py:384-416, 530-584): torch.nn.RMSNorm of the input rows in front (norm_weight float32 [K], may be NULL
 * = no norm) and the residual add behind (residual float32 [M][N], may be NULL; y may alias it):
 *   y = residual + rmsnorm(x; norm_weight, eps) . W */
int dia_b200_dense_forward_fused(const float *x, const float *norm_weight, float eps, const void *wt_bf16,
                                 const float *residual, float *y, void *workspace, int M, int N, int K, void *stream);

/* ---- the rest of the T > 1 passes (once per utterance): no library kernel on the path ---------------------------
 * F.scaled_dot_product_attention call sites (dia/layers.py:329-337) for more than one query row, fp32.
 *   q, out float32 [B][Tq][Hq][128] (q already rotated); k, v float32 [B][Hkv][Tk_stride][128] (KVCache layout,
 *   dia/state.py:72-86), the first Tk keys are used; query head h reads kv head h / (Hq / Hkv) (GQA, :319-320).
 *   mode 0  causal: query t attends keys [0, t]                                   prompt prefill (:722-766)
 *   mode 1  partition at n = n_valid_host[b]: queries < n attend keys [0, n), the others keys [n, Tk)
 *                                                                                 encoder mask (dia/state.py:24-31)
 *   mode 2  prefix: every query attends keys [0, n); n == 0 gives exact zeros     cross-attention (Appendix C Q7)
 * n_valid_host: int32 [B] in HOST memory (may be NULL = Tk).  B <= 16. */
int dia_b200_attention_rows(const float *q, const float *k, const float *v, float *out, int B, int Tq, int Tk, int Hq,
                            int Hkv, int Tk_stride, int mode, const int32_t *n_valid_host, void *stream);
/* RotaryEmbedding (dia/layers.py:108-173) on projected heads src float32 [B*T][H*128] with positions pos int32 [B*T]
 * (device) looked up in host-made sin/cos tables float32 [n_pos][64] (device; the reference's CPU values, bit for bit),
 * fused with the layout change: to_cache = 0 writes dst in the source layout (may alias src), to_cache = 1 writes
 * dst float32 [B][H][dst_T][128] at rows dst_t0 + t (KVCache.prefill / from_kv, dia/state.py:88-109).
 * rotate = 0 copies without rotating (the V projection). */
int dia_b200_rope_rows(const float *src, float *dst, const float *sin_tab, const float *cos_tab, const int32_t *pos,
                       int B, int T, int H, int rotate, int to_cache, int dst_T, int dst_t0, int n_pos, void *stream);
/* torch.nn.RMSNorm over M rows of length D (dia/layers.py:462,766): y = x * rsqrt(mean(x^2) + eps) * weight */
int dia_b200_rmsnorm_rows(const float *x, const float *weight, float eps, float *y, int M, int D, void *stream);
/* MlpBlock gate (dia/layers.py:95-101): gu float32 [M][2][F] (gate columns, then up columns) -> h[M][F] */
int dia_b200_silu_mul(const float *gu, float *h, int M, int F, void *stream);
/* nn.Embedding (dia/layers.py:445-447): out[r] = table[ids[r]], table float32 [vocab][D], ids int32 (device) */
int dia_b200_embed_rows(const float *table, const int32_t *ids, float *out, int n_rows, int vocab, int D, void *stream);

/* ---- introspection for tests and the bench ------------------------------------------------- */
enum dia_b200_buffer {
    DIA_B200_BUF_X = 0,      /* residual stream after the last executed stage, interleaved [D][2] */
    DIA_B200_BUF_LOGITS = 6, /* [2][C][V]                                */
    DIA_B200_BUF_PRED = 7,   /* int32 [C] raw prediction of the last step */
    DIA_B200_BUF_TIMING = 8, /* int64 [16][stages][16]: SM-clock stamps of CTA 0 inside each stage */
    DIA_B200_BUF_CTA_TIMING = 9 /* uint64 [stages][n_ctas]: %globaltimer (ns) at the end of each stage of step 1 */
};
/* stage ids inside one decode step: 0 = embed, 1+8*l+{0..7} = qkv, self-attn, self-o, cross-q,
 * cross-attn, cross-o, mlp-in, mlp-out of layer l, 1+8*L = logits, 2+8*L = sample.  A launch may begin at
 * the embedding, at a layer or at the logits head and end after any of those (the residual stream is handed
 * from launch to launch in DIA_B200_BUF_X); the `cooperative` argument is ignored (always cooperative). */
int dia_b200_debug_run_stages(dia_b200_engine *e, const int32_t *tokens, int stage_begin, int stage_end, int pos,
                              int slot, int cooperative, void *stream);
/* per-stage timestamps for launches of <= 16 steps (profiling aid; a persistent kernel is opaque to ncu);
 * enable = 0 switches them off, enable = 1 + c records CTA c */
int dia_b200_debug_enable_timing(dia_b200_engine *e, int enable);
int dia_b200_debug_read(dia_b200_engine *e, int which, void *host_dst, size_t nbytes, void *stream);
/* {code, block, thread, stage sequence number} written by a kernel watchdog into pinned host memory: readable
 * even after the launch was killed (codes: 1-4 ring / progress barriers, 5 bad state, 6 data-flag timeout) */
int dia_b200_debug_last_device_error(dia_b200_engine *e, int32_t *out, int n_words);
/* words [4..7]: site-specific detail; [16 + 2*(block*10 + warp)]: (site, info) of every warp that was waiting */
/* number of kernels this library has launched since load (for bench.py's gpu_launches) */
int64_t dia_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DIA_B200_H */
