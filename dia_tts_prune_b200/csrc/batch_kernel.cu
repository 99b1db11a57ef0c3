// batch_kernel.cu - the persistent decode-step kernel for N utterances per GPU.
//
// Same dataflow as step_kernel.cu - one CTA per SM, a producer warp streaming this CTA's weight slabs through a
// shared-memory ring with 1-D bulk copies, flag-in-data vectors between CTAs, no grid barrier, the loop of
// Dia.generate on the device - but the batch is R = 2N rows (CFG uncond / cond of N utterances, each with its own
// KV caches, text length, token grid, EOS state and RNG stream), so ONE pass over the 2.53 GB of weights produces N
// frames.  What changes with the rows:
//
//  * the GEMM stages run on the 5th-generation tensor cores: the CTA's slab is the M operand (<= 128 output columns,
//    K-major, 128-byte swizzle - the host repacks the weights into exactly the tiles tcgen05.mma reads, one 64-row
//    k-chunk after the other; M = 64 instructions for slabs of <= 64 columns), the activations are the N = 32 operand:
//    two bf16 terms (hi rows | lo rows) that one instruction accumulates into the same fp32 accumulator in tensor
//    memory.  FOUR issuers - math warps 4..7 - each own a quarter of the contraction (their own ring slots, activation
//    stages and accumulator: every barrier has one consumer) and issue from one elected lane; tcgen05.commit releases
//    the weight slot and the activation stage.  Warps 0..3 gather the RMSNorm sums meanwhile.  The epilogue (all eight
//    warps) reads the accumulators with tcgen05.ld: thread = output column, registers = 8 of the 16 rows, so RMSNorm
//    scaling, RoPE-ready q/k/v words, SiLU(gate) * up, the residual add (the stream lives in registers of its column's
//    thread) and the CFG combine are all local;
//  * the input vector of a GEMM stage crosses CTAs as the shared-memory image of the N operand itself (per 64-k chunk
//    a hi tile and a lo tile of [16 rows][64 k] bf16, swizzled): the producing epilogues store 2-byte terms straight
//    into that image in global memory, every CTA adds to the buffer's arrival counter with release semantics (once per
//    CTA and stage), and a dedicated lane of each consuming CTA polls the counter (acquire) and then streams the vector
//    into a 12-stage ring with 8 KB bulk copies (TMA engine) - no register staging, no per-word flags.  (The first
//    version moved (hi | lo | 1-bit flag) words through registers: 2 608 stage hand-offs per step between the math
//    warps and the MMA warp were 60 % of the step.)
//  * attention, embedding and sampling are the single-utterance stages indexed by (utterance, row): (row, kv head)
//    pairs x key splits over the CTAs, 9 sampler CTAs per utterance, one state machine per utterance.
//
// Reference semantics: dia/layers.py:671-720, :530-584, :238-346, :92-105; dia/model.py:429-488, 32-82, 748-807.
#include <cuda_bf16.h>

#include "common.cuh"
#include "engine_internal.h"
#include "sampler.cuh"

namespace dia {

typedef unsigned long long u64;

namespace {

constexpr int kBSlotBytes = 16384;          // one ring slot: up to 8 k-chunks of a slab, or a 16-key K / V tile
#ifndef DIA_B_SLOTS
#define DIA_B_SLOTS 8
#endif
#ifndef DIA_B_STAGES
#define DIA_B_STAGES 12
#endif
constexpr int kBNumSlots = DIA_B_SLOTS;
// The activation ring must hold a multiple of kBIssuers stages: stage index i belongs to issuer i % 4, and an issuer that
// waits for (or tests) the barrier of a stage relies on the PREVIOUS generation of that barrier being complete - the one-bit
// phase parity cannot tell "not filled yet" from "filled two generations ago".  With a multiple of 4 the previous generation
// is the issuer's own, consumed earlier; with 11 stages it was another issuer's, bulk copies may land out of order, and a
// rare false positive consumed a stage before it was filled (a hang after ~10 k stages, or a wrong token).
constexpr int kActStages = DIA_B_STAGES;
static_assert(kActStages % kBIssuers == 0, "see above");              // activation (B operand) staging ring
constexpr int kBTermBytes = 2048;           // [16 rows][64 k] bf16, K-major, 128-byte swizzle
constexpr int kStageChunks = 2;             // k-chunks (64 rows each) per activation stage: [chunk][hi tile | lo tile]
constexpr int kActStageBytes = kStageChunks * 2 * kBTermBytes;
constexpr int kScratchBytes = kActStages * kActStageBytes;   // B staging | attention scratch | sampler scratch (never live together)
static_assert(kScratchBytes >= 49152, "attention / sampler scratch");
constexpr int kBMiscBytes = 2048;
constexpr int kBSmem = kBNumSlots * kBSlotBytes + kScratchBytes + kBMiscBytes + 1024;   // + alignment slack
constexpr int kMmaWarp = 8;                 // allocates / frees the tensor memory (the MMAs are issued by math warps 4..7)
constexpr int kIssuerWarp0 = 4;             // math warps 4..7 issue the MMAs of a GEMM stage (each on its quarter of the contraction)
static_assert(kStageChunks == 2, "an issuer's stage is two chunks");
constexpr int kActWarp = 11;                // one lane: waits for the input buffer of every GEMM stage, streams it into the B ring
constexpr int kBThreads = 12 * 32;          // 8 math warps, TMEM warp, producer warp, (idle), activation warp
constexpr int kRows = 16;                   // batch rows of the B tiles (2 x kMaxUtt); N of the MMA = 2 kRows (hi rows, then lo rows)
#ifndef DIA_ACC_PER_ISSUER
#define DIA_ACC_PER_ISSUER 1
#endif
constexpr int kAccPerIssuer = DIA_ACC_PER_ISSUER;
// (back-to-back MMAs of ONE issuing thread on one accumulator serialise on its read-modify-write latency, ~115 cycles per
// instruction; with four issuers the pipe interleaves their instructions, and one accumulator each is 0.8 % faster than two:
// half the TMEM loads in the epilogue)
constexpr int kNumAcc = kAccPerIssuer * kBIssuers;
constexpr int kAccCols = kNumAcc * 2 * kRows;   // TMEM columns: accumulators x (hi | lo)
constexpr int kTmemCols = kAccCols;         // (one set: the warps that issue the MMAs of a stage also run its epilogue)

struct BMisc {
    uint64_t full[kBNumSlots], empty[kBNumSlots];
    uint64_t bfull[kActStages], bempty[kActStages];
    uint64_t acc_full;
    uint32_t tmem_base;
    int stages_done;
    int trace[kConsumerWarps][2];           // bring-up build: (marker, stage) of every math warp
    int ready_seq;                          // the input buffer of GEMM stage `ready_seq` (and of every earlier one) is complete
    CtaTable tab;
    float inv[kRows];                       // 1/rms of the stage input per row
    float ssq_part[kConsumerWarps][kRows];
    int toks[kMaxUtt][DIA_B200_MAX_CHANNELS];
};
static_assert(sizeof(BMisc) <= kBMiscBytes, "misc region too small");

struct BCtx {
    const BatchParams* p;
    unsigned char* ring;
    unsigned char* scratch;
    BMisc* misc;
    int tid, warp, lane;
    unsigned cbase;       // ring slot index at the start of the current stage
    unsigned bctr;        // activation stages consumed so far
    unsigned gctr;        // GEMM stages (with columns in this CTA) so far
    unsigned seq;         // sequence number of the current stage inside this launch (>= 1)
    int step;             // step index inside the launch
    float xres[kRows / 2];  // warps 0 and 4, lanes 0..15: this column's element of the residual stream, rows 0..7 / 8..15
    long long t_prof[8];  // CTA 0, thread 0 (p.prof): 0 wait for the stage input, 1 -, 2 rms gather, 3 accumulator wait,
                          // 4 epilogue, 5 end-of-stage barrier, 6 attention stages, 7 embed + sample
    bool prof;
    long long t_tmem;     // CTA 0, thread 0: the accumulator loads of the epilogues
    bool prof_issue;      // CTA 0, warp 4, lane 0: where the first MMA issuer spends its clocks (0 total, 1 bfull waits, 2 ring waits)
    long long t_issue[3];
};

__device__ __forceinline__ void decode_stage_b(int s, int L, int& kind, int& layer) {
    if (s == 0) { kind = S_EMBED; layer = 0; }
    else if (s <= 8 * L) { layer = (s - 1) >> 3; kind = S_QKV + ((s - 1) & 7); }
    else if (s == 8 * L + 1) { kind = S_LOGITS; layer = 0; }
    else { kind = S_SAMPLE; layer = 0; }
}
__device__ __forceinline__ int gemm_of_kind_b(int kind) {
    switch (kind) {
        case S_QKV: return G_QKV;
        case S_SO: return G_SO;
        case S_CQ: return G_CQ;
        case S_CO: return G_CO;
        case S_WI: return G_WI;
        case S_WO: return G_WO;
        case S_LOGITS: return G_LOGITS;
        default: return -1;
    }
}

// ---- activation buffers -----------------------------------------------------------------------------------------------------
// Every consumer GEMM has its own input buffer - xq (qkv), xc (cross-q), xm (mlp-in), xl (logits), attn (self-o), cattn
// (cross-o), hidden (mlp-out); each is written once per layer (xl once per step) and EVERY CTA adds kArrivalsPerCta to the
// buffer's counter per generation, whether it owns a piece of the vector or not, and only after it has passed every
// earlier stage.  A consumer of generation g waits for counter >= g * G * kArrivalsPerCta; every CTA waits at every GEMM
// stage (also one without columns there), so no CTA can arrive for generation g + 1 before all have arrived for g, and a
// buffer is only rewritten after all its readers have passed a later all-to-all stage.
__device__ __forceinline__ int act_of_gemm(int gt) {
    switch (gt) {
        case G_QKV: return A_XQ;
        case G_SO: return A_ATTN;
        case G_CQ: return A_XC;
        case G_CO: return A_CATTN;
        case G_WI: return A_XM;
        case G_WO: return A_HIDDEN;
        default: return A_XL;
    }
}
__device__ __forceinline__ unsigned act_generation(int gt, int L, int step, int layer) {
    return gt == G_LOGITS ? (unsigned)(step + 1) : (unsigned)(step * L + layer + 1);
}
__device__ __forceinline__ const unsigned char* act_buffer(const BatchParams& p, int a) {
    switch (a) {
        case A_XQ: return p.act_xq;
        case A_XC: return p.act_xc;
        case A_XM: return p.act_xm;
        case A_XL: return p.act_xl;
        case A_ATTN: return p.act_attn;
        case A_CATTN: return p.act_cattn;
        default: return p.act_hidden;
    }
}
__device__ __forceinline__ void act_arrive(const BatchParams& p, int a, unsigned n) { red_release_add_u32(p.ctr + a * 32, n); }

// x ~ hi + lo, both bf16, stored at the element's place in the swizzled tiles of its 64-k chunk (see engine_internal.h)
__device__ __forceinline__ void st_act(unsigned char* buf, int k, int r, float x) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    unsigned char* dst = buf + (size_t)(k >> 6) * (2 * kBTermBytes) + r * 128 + (((((k & 63) >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1));
    asm volatile("st.relaxed.gpu.global.b16 [%0], %1;" ::"l"(dst), "h"(__bfloat16_as_ushort(h)) : "memory");
    asm volatile("st.relaxed.gpu.global.b16 [%0], %1;" ::"l"(dst + kBTermBytes), "h"(__bfloat16_as_ushort(l)) : "memory");
}
__device__ __forceinline__ float4 ld_gpu_f4(const float* p) {
    float4 r;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_gpu_f(float* p, float v) {
    asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}

// latency-critical waits of the activation pipeline: poll without the suspend hint (a parked warp wakes up late)
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity, int* err, int code, unsigned info) {
    unsigned polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++polls > 200000000u) ll_timeout(err, code, info);
        ll_check_abort(err, polls, 100 + code, info);
    }
}

__device__ __forceinline__ void ring_wait_full_b(BMisc* misc, unsigned idx, int* err, unsigned info) {
    const unsigned slot = idx % kBNumSlots, parity = (idx / kBNumSlots) & 1u;
    mbar_wait(&misc->empty[slot], parity ^ 1u, err, kErrEmptyBarrierTimeout, info);     // see step_kernel.cu: parity alias
    mbar_wait(&misc->full[slot], parity, err, kErrFullBarrierTimeout, info);
}

// ---- tcgen05 ---------------------------------------------------------------------------------------------------
// K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (same encoding as gemm_tcgen05.cu)
__device__ __forceinline__ uint64_t umma_desc_b(uint32_t smem_addr) {
    const uint32_t lo = ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void umma_bf16_b(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same arrive delivered to the barrier at this offset in every CTA of `mask` (the CTA pair shares each activation stage)
__device__ __forceinline__ void umma_commit_mc_b(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_commit_b(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- work decomposition of the attention stages ---------------------------------------------------------------------
struct BAttnWork {
    int active, pair, split, k_lo, k_hi, n_active, has_new, u, row, head;
};
// self-attention: (row, kv head) pairs x key splits.  Old keys are the cache slots [0, slot_u); the key / value of THIS
// step comes from the qkv words and is handled (and appended) by the last active split.
__device__ __forceinline__ BAttnWork self_work_b(const BatchParams& p, int cta, int step) {
    BAttnWork w;
    const int nsplit = p.sa_nsplit, pairs = p.R * p.Hkv;
    w.pair = cta / nsplit;
    w.split = cta - w.pair * nsplit;
    w.row = w.pair / p.Hkv;
    w.head = w.pair - w.row * p.Hkv;
    w.u = w.row >> 1;
    // (an utterance that has used up its budget - a longer prompt than its neighbours' - idles at the last cache slot)
    const int n_old = w.pair < pairs ? min(p.utt[w.u].slot0 + step, p.Lmax - 1) : 0;
    int per = (n_old + nsplit - 1) / nsplit;
    per = (per + 15) & ~15;
    if (per < 64) per = 64;
    w.n_active = (n_old + per - 1) / per;
    if (w.n_active < 1) w.n_active = 1;
    w.active = (w.pair < pairs) && (w.split < w.n_active);
    w.k_lo = w.split * per;
    w.k_hi = min(n_old, w.k_lo + per);
    if (w.k_hi < w.k_lo) w.k_hi = w.k_lo;
    w.has_new = (w.split == w.n_active - 1);
    return w;
}
// cross-attention: conditional rows only, (utterance, head) pairs x key splits over the valid text keys
__device__ __forceinline__ BAttnWork cross_work_b(const BatchParams& p, int cta) {
    BAttnWork w;
    const int nsplit = p.ca_nsplit, pairs = p.U * p.Hc;
    w.pair = cta / nsplit;
    w.split = cta - w.pair * nsplit;
    w.u = w.pair / p.Hc;
    w.head = w.pair - w.u * p.Hc;
    w.row = 2 * w.u + 1;
    const int n = w.pair < pairs ? p.utt[w.u].text_len : 0;
    int per = (n + nsplit - 1) / nsplit;
    per = (per + 15) & ~15;
    if (per < 128) per = 128;
    w.n_active = (n + per - 1) / per;
    if (w.n_active < 1) w.n_active = 1;
    w.active = (w.pair < pairs) && (w.split < w.n_active);
    w.k_lo = w.split * per;
    w.k_hi = min(n, w.k_lo + per);
    if (w.k_hi < w.k_lo) w.k_hi = w.k_lo;
    w.has_new = 0;
    return w;
}
// ring slots of one attention stage of this CTA: one per 16-key tile (K tile in the first 8 KB, V tile in the second)
__device__ __forceinline__ int attn_slots_b(const BAttnWork& w) {
    return (!w.active || w.k_hi <= w.k_lo) ? 0 : ((w.k_hi - w.k_lo + 15) >> 4);
}

// ---- producer: walks this CTA's byte stream -----------------------------------------------------------------------------
__device__ void producer_loop_b(const BatchParams& p, unsigned char* ring, BMisc* misc) {
    unsigned pc = 0;
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    const CtaTable& tab = misc->tab;
    const int cta = blockIdx.x;
    const int S = 8 * p.L + 3, n_stage = p.with_sample ? S : S - 1;
    auto issue = [&](const void* src, uint32_t bytes, bool keep, unsigned seq) {
        const unsigned slot = pc % kBNumSlots, ph = (pc / kBNumSlots) & 1u;
        mbar_wait(&misc->empty[slot], ph ^ 1u, p.err, kErrEmptyBarrierTimeout, (seq << 8) | slot);
        mbar_arrive_expect_tx(&misc->full[slot], bytes);
        bulk_g2s_hint(ring + slot * kBSlotBytes, src, bytes, &misc->full[slot], keep ? pol_keep : pol_stream);
        pc++;
    };
    auto issue2 = [&](const void* src0, const void* src1, uint32_t bytes_each, bool keep, unsigned seq) {
        const unsigned slot = pc % kBNumSlots, ph = (pc / kBNumSlots) & 1u;
        mbar_wait(&misc->empty[slot], ph ^ 1u, p.err, kErrEmptyBarrierTimeout, (seq << 8) | slot);
        mbar_arrive_expect_tx(&misc->full[slot], 2 * bytes_each);
        bulk_g2s_hint(ring + slot * kBSlotBytes, src0, bytes_each, &misc->full[slot], keep ? pol_keep : pol_stream);
        bulk_g2s_hint(ring + slot * kBSlotBytes + kBSlotBytes / 2, src1, bytes_each, &misc->full[slot], keep ? pol_keep : pol_stream);
        pc++;
    };
#pragma unroll 1
    for (int n = 0; n < p.n_steps; ++n) {
#pragma unroll 1
        for (int s = 0; s < n_stage; ++s) {
            int kind, layer;
            decode_stage_b(s, p.L, kind, layer);
            const unsigned seq = 1u + (unsigned)(n * S + s);
            const int gt = gemm_of_kind_b(kind);
            if (gt >= 0) {
                const int gc = tab.gc[gt];
                if (gc == 0) continue;
                const int K = p.Kdim[gt], cps = bslot_chunks(gc, K);
                const uint32_t slot_bytes = (uint32_t)(cps * bchunk_bytes(gc));
                const int n_slots = (K / 64) / cps;
                const unsigned char* base = p.wstream + tab.stream_base +
                    (gt == G_LOGITS ? tab.logits_off : (unsigned long long)layer * tab.layer_bytes + tab.slab_off[gt]);
#pragma unroll 1
                for (int i = 0; i < n_slots; ++i) issue(base + (size_t)i * slot_bytes, slot_bytes, false, seq);
            } else if (kind == S_SATTN || kind == S_CATTN) {
                const bool self = kind == S_SATTN;
                const BAttnWork w = self ? self_work_b(p, cta, n) : cross_work_b(p, cta);
                if (attn_slots_b(w) == 0) continue;
                if (self && n > 0) {
                    // cache row slot-1 was appended by another CTA in this stage of step n-1 and fenced before its stage
                    // output; this CTA's math warps are past the cross-q stage of step n-1 once stages_done says so
                    const int need = (n - 1) * S + s + 3;
                    const unsigned long long t0 = clock64();
                    while (ld_acquire_cta_s32(&misc->stages_done) < need) {
                        if (clock64() - t0 > kWatchdogCycles) ll_timeout(p.err, kErrStepDoneTimeout, seq);
                    }
                    fence_proxy_async();
                }
                const float *kb, *vb;
                if (self) {
                    const size_t off = ((size_t)((w.row & 1) * p.Hkv + w.head) * p.Lmax) * kHeadDim;
                    kb = p.self_k[w.u * p.L + layer] + off;
                    vb = p.self_v[w.u * p.L + layer] + off;
                } else {
                    const size_t off = ((size_t)(p.Hc + w.head) * p.Smax) * kHeadDim;       // row 1 (cond) of the utterance
                    kb = p.cross_k[w.u * p.L + layer] + off;
                    vb = p.cross_v[w.u * p.L + layer] + off;
                }
#pragma unroll 1
                for (int k0 = w.k_lo; k0 < w.k_hi; k0 += 16) {
                    const int nk = min(16, w.k_hi - k0);
                    issue2(kb + (size_t)k0 * kHeadDim, vb + (size_t)k0 * kHeadDim, nk * kHeadDim * 4, !self, seq);
                }
            }
        }
    }
}

// ---- activation lane: for every GEMM stage, wait until the stage's input buffer is complete (all CTAs have arrived), tell
//      the math warps, then stream the vector into the B ring: one 8 KB bulk copy per 128 k of all 16 rows ------------------------
__device__ void act_loop_b(const BatchParams& p, unsigned char* scratch, BMisc* misc) {
    const CtaTable& tab = misc->tab;
    const int S = 8 * p.L + 3, n_stage = p.with_sample ? S : S - 1;
    unsigned bctr = 0;
    const bool prof = p.prof != nullptr && blockIdx.x == 0;
    long long t_ready = 0, t_bempty = 0, tq = 0;
    const uint32_t rank = p.mc ? cluster_ctarank() : 0u;
#pragma unroll 1
    for (int n = 0; n < p.n_steps; ++n) {
#pragma unroll 1
        for (int s = 0; s < n_stage; ++s) {
            int kind, layer;
            decode_stage_b(s, p.L, kind, layer);
            const int gt = gemm_of_kind_b(kind);
            if (gt < 0) continue;
            const unsigned seq = 1u + (unsigned)(n * S + s);
            const int a = act_of_gemm(gt);
            const unsigned need = act_generation(gt, p.L, n, layer) * (unsigned)p.G * kArrivalsPerCta;
            if (prof) tq = clock64();
            {
                const unsigned* ctr = p.ctr + a * 32;
                unsigned polls = 0;
                while (ld_acquire_u32(ctr) < need) {
#ifdef DIA_BATCH_TRACE
                    if (++polls > kMaxSpins || ((polls & 0x3fffu) == 0x3fffu && reinterpret_cast<volatile int*>(p.err)[0] != 0)) {
                        for (int i = 0; i < 100; ++i) __nanosleep(1000000);      // let the warps that can report do so
                        volatile int* e = reinterpret_cast<volatile int*>(p.err);
                        for (int w = 0; w < kConsumerWarps; ++w) {
                            const int sl = 16 + (blockIdx.x * 12 + w) * 2;
                            if (e[sl] == 0) { e[sl] = reinterpret_cast<volatile int*>(misc->trace[w])[0]; e[sl + 1] = reinterpret_cast<volatile int*>(misc->trace[w])[1]; }
                        }
                        __threadfence_system();
                    }
#endif
                    if (++polls > kMaxSpins) {
                        volatile int* e = reinterpret_cast<volatile int*>(p.err);
                        e[4] = (int)ld_relaxed_u32(ctr); e[5] = gt; e[6] = (int)need; e[7] = a;
                        ll_timeout(p.err, kErrFlagTimeout, seq * 16 + gt);
                    }
                    ll_check_abort(p.err, polls, 100 + kErrFlagTimeout, seq * 16 + gt);
                }
            }
            if (prof) t_ready += clock64() - tq;
            st_release_cta_s32(&misc->ready_seq, (int)seq);
            const int gc = tab.gc[gt];
            if (gc == 0) continue;
            fence_proxy_async();              // the vector was written through the generic proxy (by other SMs); bulk copies read through the async proxy
            const unsigned char* src = act_buffer(p, a);
            const int n_st = p.Kdim[gt] / (64 * kStageChunks);
            const int per_issuer = n_st / kBIssuers;          // stage 4 t + w of the ring is the t-th stage of issuer w's quarter
#pragma unroll 1
            for (int su = 0; su < n_st; ++su) {
                const int st = (su & (kBIssuers - 1)) * per_issuer + (su >> 2);
                const unsigned bi = bctr + (unsigned)su, bs = bi % kActStages;
                // The first kActStages copies of a GEMM stage need no wait: the stage's input is complete, so every CTA - this one
                // and its pair too - has finished the stage that produced it, i.e. all MMAs of its previous GEMM have retired and
                // every activation stage is free (a wait changes no barrier state; an mbarrier test costs ~120 cycles, and the four
                // issuers start with copies 0..3)
                // (not when a CTA pair shares its stages: the pair's multicast arrive may still be in flight)
                if (su >= kActStages || p.mc) {
                    if (prof) tq = clock64();
                    mbar_spin(&misc->bempty[bs], ((bi / kActStages) & 1u) ^ 1u, p.err, kErrEmptyBarrierTimeout, (seq << 8) | 0xc0 | bs);
                    if (prof) t_bempty += clock64() - tq;
                }
                mbar_arrive_expect_tx(&misc->bfull[bs], kActStageBytes);
                if (p.mc) {
                    // this CTA fetches its half of the stage (the chunk of its cluster rank) for both CTAs of the pair; the
                    // peer's half arrives the same way.  bempty[bs] has collected the commits of BOTH CTAs' MMA warps.
                    const uint32_t half = rank * (kActStageBytes / 2);
                    bulk_g2s_mc(scratch + bs * kActStageBytes + half, src + (size_t)st * kActStageBytes + half, kActStageBytes / 2,
                                &misc->bfull[bs], (uint16_t)3);
                } else {
                    bulk_g2s(scratch + bs * kActStageBytes, src + (size_t)st * kActStageBytes, kActStageBytes, &misc->bfull[bs]);
                }
            }
            bctr += (unsigned)n_st;
        }
    }
    if (prof) {
        p.prof[4] = (unsigned long long)t_ready;
        p.prof[5] = (unsigned long long)t_bempty;
    }
}

// sum v[i] over the 32 lanes for NV values at once (see step_kernel.cu)
template <int NV>
__device__ __forceinline__ void transpose_reduce_b(float (&v)[NV], int lane) {
    int n = NV;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        if (n > 1) {
            n >>= 1;
            const bool hi = (lane & m) != 0;
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) {
                if (i < n) {
                    const float a = v[i], b = v[i + n];
                    const float send = hi ? a : b;
                    const float keep = hi ? b : a;
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
                }
            }
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], m);
        }
    }
}

// ---- MMA issue: math warps 4..7, each on its own quarter of the contraction (its own ring slots, activation stages and two
//      accumulators).  The whole warp walks the loop (warp-uniform control flow and operands), one elected lane issues.
// Measured (tools/microbench/umma_issue_bench.cu): one of these M128 x N32 x K16 instructions costs its issuing warp ~43 cycles,
// a tcgen05.commit ~40, an mbarrier test ~120 before its predicate can be read - against a tensor-pipe floor of 16 cycles per
// MMA.  Hence four issuers that share nothing (a slot / stage barrier collects ONE commit), and the barrier of an issuer's next
// stage is tested before the MMAs of the current one are issued (the predicate, a PTX register declared at kernel scope, is
// read after them; only a miss falls back to the spin).
__device__ __forceinline__ void issue_gemm_b(BCtx& c, int gt, int K, int gc) {
    const BatchParams& p = *c.p;
    BMisc* misc = c.misc;
    const int wi = c.warp - kIssuerWarp0;
    const int n_chunks = K / 64, Q = n_chunks / kBIssuers, cps = bslot_chunks(gc, K), n_t = Q / kStageChunks;
    const uint32_t chunk_bytes = (uint32_t)bchunk_bytes(gc);
    const int cshift = __ffs(cps) - 1;
    // instruction descriptor: fp32 accumulate, bf16 x bf16, both operands K-major, M = 128, N = 32: the hi tile and the lo
    // tile of the activations are adjacent in shared memory and go through ONE instruction; columns 0..15 of an accumulator
    // are W.hi, 16..31 are W.lo.  Back-to-back MMAs on ONE accumulator serialise on its read-modify-write latency
    // (measured: ~115 cycles per instruction), when ONE thread issues them; see kNumAcc.
    // Slabs of up to 64 columns (everything but mlp-in) use M = 64: the tensor pipe is busy ~34 cycles per instruction
    // instead of ~40 (tools/microbench/umma_issue_bench.cu); row m of the slab then sits in TMEM lane 32 (m / 16) + m % 16.
    const uint32_t mdim = gc <= 8 ? 64u : 128u;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * kRows) >> 3) << 17) | ((mdim >> 4) << 24);
    const uint32_t ring_u32 = smem_u32(c.ring), stage_u32 = smem_u32(c.scratch);
    const uint32_t d0 = misc->tmem_base + (uint32_t)wi * (kAccPerIssuer * 2 * kRows);
    const bool prof = c.prof_issue;
    long long tq = 0;
    bool have = false;
    unsigned sidx = 0;                                   // ring slot under the current chunk
#pragma unroll 1
    for (int t = 0; t < n_t; ++t) {
        // the activation stage first: it is complete only after every CTA - this one too - has left the attention stage
        // before this GEMM, i.e. every earlier generation of the ring slots has been released (the one-bit phase parity of
        // the slot barriers cannot tell generations two apart)
        const unsigned bi = c.bctr + (unsigned)(kBIssuers * t + wi), bs = bi % kActStages, bpar = (bi / kActStages) & 1u;
        if (prof) tq = clock64();
        if (!have) mbar_spin(&misc->bfull[bs], bpar, p.err, kErrFullBarrierTimeout, (c.seq << 8) | 0x80 | bs);
        if (prof) c.t_issue[1] += clock64() - tq;
        const bool peek = t + 1 < n_t;
        if (peek) {
            const unsigned nbi = bi + kBIssuers, nbs = nbi % kActStages, npar = (nbi / kActStages) & 1u;
            asm volatile("mbarrier.test_wait.parity.shared::cta.b64 dia_pw_bfull, [%0], %1;" ::"r"(smem_u32(&misc->bfull[nbs])), "r"(npar) : "memory");
        }
#pragma unroll
        for (int jc = 0; jc < kStageChunks; ++jc) {
            const int cc = t * kStageChunks + jc, in_slot = cc & (cps - 1);
            if (in_slot == 0) {
                sidx = c.cbase + (unsigned)(kBIssuers * (cc >> cshift) + wi);
                if (prof) tq = clock64();
                ring_wait_full_b(misc, sidx, p.err, (c.seq << 8) | (sidx % kBNumSlots));
                if (prof) c.t_issue[2] += clock64() - tq;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned slot = sidx % kBNumSlots;
            const uint64_t adesc = umma_desc_b(ring_u32 + slot * kBSlotBytes + (uint32_t)in_slot * chunk_bytes);
            const uint64_t bdesc = umma_desc_b(stage_u32 + bs * kActStageBytes + (uint32_t)jc * (2 * kBTermBytes));
            if (elect_one_sync()) {
#pragma unroll
                for (int j = 0; j < 4; ++j)           // 16 elements along K = 32 bytes = 2 descriptor units
                    umma_bf16_b(d0 + (uint32_t)(j & (kAccPerIssuer - 1)) * (2 * kRows), adesc + 2 * j, bdesc + 2 * j, idesc,
                                (cc != 0 || j >= kAccPerIssuer) ? 1u : 0u);
                if (in_slot == cps - 1) umma_commit_b(&misc->empty[slot]);
                if (jc == kStageChunks - 1) {
                    if (p.mc) umma_commit_mc_b(&misc->bempty[bs], (uint16_t)3);
                    else umma_commit_b(&misc->bempty[bs]);
                }
            }
            __syncwarp();
        }
        have = false;
        if (peek) {
            uint32_t ok;
            asm volatile("selp.u32 %0, 1, 0, dia_pw_bfull;" : "=r"(ok));
            have = ok != 0;
        }
    }
    if (elect_one_sync()) umma_commit_b(&misc->acc_full);
    __syncwarp();
}

// bring-up build (-DDIA_BATCH_TRACE): every math warp leaves (marker, stage) in its slot of the watchdog record, so that a
// warp blocked in a block barrier (which cannot report by itself) can be located after a timeout
#ifdef DIA_BATCH_TRACE
#define BTRACE(c, code)                                                                                       \
    do {                                                                                                      \
        if ((c).lane == 0 && ((DIA_BATCH_TRACE >> ((code) & 31)) & 1)) {                                      \
            volatile int* t_ = reinterpret_cast<volatile int*>((c).misc->trace[(c).warp]);                    \
            t_[0] = (code);                                                                                   \
            t_[1] = (int)(c).seq;                                                                             \
        }                                                                                                     \
    } while (0)
#else
#define BTRACE(c, code) do { } while (0)
#endif

// ---- GEMM stage, math warps: the RMSNorm sums while the MMAs run, then the epilogue ----------------------------------------------
// (the activation lane streams the input vector, the MMA warps consume it)
__device__ __forceinline__ int out_act_of_gemm(int gt, bool last_layer) {       // the buffer a GEMM stage's epilogue writes, or -1
    return gt == G_SO ? A_XC : gt == G_CO ? A_XM : gt == G_WI ? A_HIDDEN : gt == G_WO ? (last_layer ? A_XL : A_XQ) : -1;
}
__device__ void gemm_stage_b(BCtx& c, int gt, int layer) {
    const BatchParams& p = *c.p;
    BMisc* misc = c.misc;
    const int gc = misc->tab.gc[gt], g0 = misc->tab.g0[gt];
    const int R = p.R, L = p.L, K = p.Kdim[gt], n_chunks = K / 64, NC = gc * 8;
    const int tid = c.tid, lane = c.lane, warp = c.warp;
    const bool resid = (gt == G_SO || gt == G_CO || gt == G_WO);
    const bool normed = !resid;
    const int out_act = out_act_of_gemm(gt, layer == L - 1);

    // every CTA waits for the stage input, with or without columns in this GEMM (see "activation buffers" above)
    {
        long long tq = c.prof ? clock64() : 0;
        // (every lane polls the same shared-memory word: the branch is warp-uniform, the warp reaches the block barriers
        // below converged - a lane-0-only spin loop left the warp split in two around `bar.sync`, each half counting as a
        // whole warp's arrival)
        unsigned polls = 0;
        while (ld_acquire_cta_s32(&misc->ready_seq) < (int)c.seq) {
            if (++polls > 400000000u) ll_timeout(p.err, kErrFlagTimeout + 1, c.seq * 16 + gt);
            ll_check_abort(p.err, polls, 100 + kErrFlagTimeout + 1, c.seq * 16 + gt);
        }
        __syncwarp();
        if (c.prof) c.t_prof[0] += clock64() - tq;
    }
    BTRACE(c, 1);
    if (gc == 0) {
        if (out_act >= 0 && tid == 0) act_arrive(p, out_act, kArrivalsPerCta);
        return;
    }
    // ---- main loop: warps 4..7 issue the MMAs; warps 0..3 meanwhile gather the RMSNorm sums ---------------------------------------
    long long tq2 = c.prof ? clock64() : 0;
    if (warp >= kIssuerWarp0) {
        const long long ti = c.prof_issue ? clock64() : 0;
        issue_gemm_b(c, gt, K, gc);
        if (c.prof_issue) c.t_issue[0] += clock64() - ti;
    } else if (normed) {
        // 1/rms per row from the producers' per-group sums (dia/layers.py:541,560,579,714): thread g takes the 8-column groups
        // g, g + 128, ...: the sums of all 16 rows (rows >= R are zero), every load in flight at once
        float sr[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) sr[r] = 0.f;
#pragma unroll 1
        for (int g = tid; g < (p.D >> 3); g += 4 * 32) {
            const float* base = p.ssq + (size_t)g * kRows;
            float4 q4[kRows / 4];
#pragma unroll
            for (int i = 0; i < kRows / 4; ++i) q4[i] = ld_gpu_f4(base + 4 * i);
#pragma unroll
            for (int i = 0; i < kRows / 4; ++i) {
                sr[4 * i] += q4[i].x; sr[4 * i + 1] += q4[i].y; sr[4 * i + 2] += q4[i].z; sr[4 * i + 3] += q4[i].w;
            }
        }
        BTRACE(c, 2);
        transpose_reduce_b<kRows>(sr, lane);                  // 31 shuffles instead of 16 x 5: lanes 2r, 2r + 1 hold row r
        if ((lane & 1) == 0) misc->ssq_part[warp][lane >> 1] = sr[0];
        BTRACE(c, 3);
        asm volatile("bar.sync 2, 128;" ::: "memory");         // warps 0..3 only
        if (tid < R) {
            float ss = 0.f;
#pragma unroll
            for (int ww = 0; ww < 4; ++ww) ss += misc->ssq_part[ww][tid];
            misc->inv[tid] = 1.0f / sqrtf(ss / (float)p.D + p.eps);
        }
    }
    c.cbase += (unsigned)(n_chunks / bslot_chunks(gc, K));
    c.bctr += (unsigned)(n_chunks / kStageChunks);
    BTRACE(c, 4);
    consumer_sync();            // misc->inv is written; the issuers have issued (and committed) everything
    BTRACE(c, 5);

    // ---- epilogue: thread = (output column of this CTA's slab = TMEM lane, half of the batch rows): warps 0..3 take rows
    //      0..7, warps 4..7 rows 8..15 of the same columns (a warp reads the TMEM lanes of its quadrant, warp % 4) ------------------
    constexpr int HR = kRows / 2;                               // rows per thread
    const int q = warp & 3, half = warp >> 2, r0 = half * HR;
    if (c.prof) { const long long t1 = clock64(); c.t_prof[2] += t1 - tq2; tq2 = t1; }
    {
        mbar_wait(&misc->acc_full, c.gctr & 1u, p.err, kErrGridBarrierTimeout, c.seq);
        if (c.prof) { const long long t1 = clock64(); c.t_prof[3] += t1 - tq2; tq2 = t1; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float acc[HR];
#pragma unroll
        for (int i = 0; i < HR; ++i) acc[i] = 0.f;
        if (r0 < R) {
            // accumulator j: columns 0..15 = W.hi of rows 0..15, 16..31 = W.lo; four accumulators' loads in flight per wait
#pragma unroll
            for (int jb = 0; jb < kNumAcc; jb += 4) {
                uint32_t v[4][2][HR];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const uint32_t taddr = misc->tmem_base + ((uint32_t)(q * 32) << 16) + (jb + j) * 2 * kRows + t * kRows + r0;
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                            : "=r"(v[j][t][0]), "=r"(v[j][t][1]), "=r"(v[j][t][2]), "=r"(v[j][t][3]),
                              "=r"(v[j][t][4]), "=r"(v[j][t][5]), "=r"(v[j][t][6]), "=r"(v[j][t][7])
                            : "r"(taddr)
                            : "memory");
                    }
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < HR; ++i) acc[i] += __uint_as_float(v[j][0][i]) + __uint_as_float(v[j][1][i]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (c.prof) c.t_tmem += clock64() - tq2;              // (part of t_prof[4]: the accumulator loads)
        BTRACE(c, 6);

        // column inside the slab: TMEM lane = column for M = 128; an M = 64 accumulator keeps 16 columns per lane quadrant
        const bool m64 = gc <= 8;
        const int m = m64 ? q * 16 + (lane & 15) : q * 32 + lane;
        const bool valid = m < NC && (!m64 || lane < 16);
        const int n = g0 * 8 + m;                               // column of the GEMM
        if (gt == G_QKV || gt == G_CQ) {
            u64* dst = gt == G_QKV ? p.ll_qkv : p.ll_cq;
            if (valid) {
#pragma unroll
                for (int i = 0; i < HR; ++i)
                    if (r0 + i < R) ll_st(dst + (size_t)n * R + r0 + i, __float_as_uint(acc[i] * misc->inv[r0 + i]), c.seq);
            }
        } else if (gt == G_WI) {
            // a group = gate columns 0..3 and up columns 4..7 of the same 4 hidden units: h = silu(gate) * up
            const int hn = (g0 + (m >> 3)) * 4 + (m & 3);
#pragma unroll
            for (int i = 0; i < HR; ++i) {
                if (r0 + i < R) {                               // (warp-uniform: every lane takes part in the shuffle)
                    const float y = acc[i] * misc->inv[r0 + i];
                    const float up = __shfl_down_sync(0xffffffffu, y, 4);
                    if (valid && (m & 4) == 0) st_act(p.act_hidden, hn, r0 + i, (y / (1.0f + expf(-y))) * up);
                }
            }
        } else if (gt == G_LOGITS) {
            const int ch = n / p.Vpad, vv = n - ch * p.Vpad;
            if (valid && ch < p.C && vv < p.V) {
#pragma unroll
                for (int i = 0; i < HR / 2; ++i) {
                    const int u = half * (HR / 2) + i;
                    if (u < p.U) {
                        const float un = acc[2 * i] * misc->inv[2 * u];
                        const float co = acc[2 * i + 1] * misc->inv[2 * u + 1];
                        if (p.logits != nullptr) {
                            p.logits[((size_t)(2 * u) * p.C + ch) * p.V + vv] = un;
                            p.logits[((size_t)(2 * u + 1) * p.C + ch) * p.V + vv] = co;
                        }
                        float gv = __fadd_rn(co, __fmul_rn(p.cfg_scale, __fsub_rn(co, un)));     // dia/model.py:450-457
                        if ((ch > 0 && vv == p.eos) || vv == p.pad || vv == p.bos) gv = -INFINITY;
                        ll_st(p.ll_glog + ((size_t)u * p.C + ch) * p.V + vv, __float_as_uint(gv), c.seq);
                    }
                }
            }
        } else {
            // residual add (dia/layers.py:555,574,582): the stream stays in this thread's registers; the new stream goes
            // out as the elements of x * w_norm for the next consumer, with sum(x^2) of this 8-column group per row
            // the next consumer: cross-q after self-o, mlp-in after cross-o, the next layer's qkv (or the logits head)
            const bool last = gt == G_WO && layer == L - 1;
            unsigned char* xdst = gt == G_SO ? p.act_xc : gt == G_CO ? p.act_xm : last ? p.act_xl : p.act_xq;
            const float* wn = gt == G_SO ? p.norms + ((size_t)layer * 3 + 1) * p.D
                            : gt == G_CO ? p.norms + ((size_t)layer * 3 + 2) * p.D
                                         : p.norms + ((size_t)(layer + 1) * 3) * p.D;
            const float wnv = valid ? __ldg(wn + n) : 0.f;
#pragma unroll
            for (int i = 0; i < HR; ++i) {
                if (r0 + i < R) {
                    float xn = 0.f;
                    if (valid) {
                        xn = c.xres[i] + acc[i];
                        c.xres[i] = xn;
                        st_act(xdst, n, r0 + i, xn * wnv);
                    }
                    float sq = xn * xn;
                    sq += __shfl_xor_sync(0xffffffffu, sq, 4);
                    sq += __shfl_xor_sync(0xffffffffu, sq, 2);
                    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                    if (valid && (m & 7) == 0) st_gpu_f(p.ssq + (size_t)(g0 + (m >> 3)) * kRows + r0 + i, sq);
                }
            }
        }
    }
    c.gctr++;
    if (c.prof) { const long long t1 = clock64(); c.t_prof[4] += t1 - tq2; tq2 = t1; }
    BTRACE(c, 7);
    consumer_sync();            // the accumulators are free for the next stage's issuers; the staging ring is scratch of the next stage
    if (c.prof) { const long long t1 = clock64(); c.t_prof[5] += t1 - tq2; tq2 = t1; }
    // the CTA's piece of the output vector is written: ONE release per CTA (148 instead of 1 184 reductions on the counter's
    // address per stage - the L2 serialises them; the barrier above makes the other warps' stores part of this release)
    if (out_act >= 0 && tid == 0) act_arrive(p, out_act, kArrivalsPerCta);
    if (c.prof) c.t_prof[1] += clock64() - tq2;
}

constexpr int HPKB = 4;          // query heads per KV tile in self-attention (GQA 4:1)

// ---- attention stage (the single-utterance stage of step_kernel.cu, indexed by utterance and row) ------------------------------
template <int NH>
__device__ void attn_body_b(BCtx& c, int layer) {
    constexpr bool self = NH == HPKB;
    const BatchParams& p = *c.p;
    const int cta = blockIdx.x, R = p.R;
    const BAttnWork w = self ? self_work_b(p, cta, c.step) : cross_work_b(p, cta);
    if (!w.active) return;
    const int r = w.row, u = w.u;
    const int kvh = w.head;
    const int head0 = self ? kvh * HPKB : w.head;
    const int nh = self ? HPKB : 1;
    const int nsplit = self ? p.sa_nsplit : p.ca_nsplit;
    const uint32_t fprev = c.seq - 1;
    const bool has_new = self && w.has_new;
    const int pos = p.utt[u].pos0 + c.step, slot = min(p.utt[u].slot0 + c.step, p.Lmax - 1);

    float* qs = reinterpret_cast<float*>(c.scratch);      // [HPKB][128] rotated, pre-scaled queries
    float* kn = qs + HPKB * kHeadDim;                      // [128] rotated key of this step
    float* vn = kn + kHeadDim;                             // [128] value of this step
    float* psm = vn + kHeadDim + c.warp * 64;              // [warps][16 keys][HPKB] probabilities of the current tile
    float* wstat = vn + kHeadDim + kConsumerWarps * 64;    // [warps][8]: m[HPKB], l[HPKB] of each warp
    float* cw = wstat + kConsumerWarps * 8;                // split-combine staging
    float* racc = reinterpret_cast<float*>(c.scratch) + 5120;   // [warps][HPKB][128]
    const u64* qsrc = self ? p.ll_qkv : p.ll_cq;

    {   // ---- inputs: q (and k, v of this step), RoPE, scale -----------------------------------------------------------------
        const int pclamp = min(pos, p.n_pos - 1);
        const int d = c.tid & 63;
        const float sn = __ldg(p.rope_sin + (size_t)pclamp * 64 + d), cs = __ldg(p.rope_cos + (size_t)pclamp * 64 + d);
        const float scale = 0.08838834764831845f;           // 1/sqrt(128)
        if ((c.warp >> 1) < nh) {
            if (c.lane == 0)
                ll_wait32(qsrc + ((size_t)(head0 + (c.warp >> 1)) * kHeadDim + (c.warp & 1) * 32) * R + r, fprev, p.err);
            __syncwarp();
        }
        const int n_items = (HPKB + 2) * 64;
#pragma unroll 1
        for (int i = c.tid; i < n_items; i += kConsumerThreads) {
            const int hh = i >> 6;
            float o0 = 0.f, o1 = 0.f;
            if (hh < nh || (hh >= HPKB && has_new)) {
                const int col = hh < HPKB ? (head0 + hh) * kHeadDim
                                          : (hh == HPKB ? (p.Hq + kvh) * kHeadDim : (p.Hq + p.Hkv + kvh) * kHeadDim);
                const u64* pa = qsrc + ((size_t)col + d) * R + r;
                const u64* pb = pa + (size_t)64 * R;
                uint2 wa = ll_ld(pa), wb = ll_ld(pb);
                if (wa.y != fprev) wa.x = ll_wait32(pa, fprev, p.err);
                if (wb.y != fprev) wb.x = ll_wait32(pb, fprev, p.err);
                const float a = __uint_as_float(wa.x), b = __uint_as_float(wb.x);
                o0 = a; o1 = b;
                if (hh <= HPKB) {                            // RotaryEmbedding (dia/layers.py:161-173)
                    o0 = a * cs - b * sn;
                    o1 = a * sn + b * cs;
                    if (hh < HPKB) { o0 *= scale; o1 *= scale; }
                }
            }
            qs[hh * kHeadDim + d] = o0;
            qs[hh * kHeadDim + d + 64] = o1;
        }
    }
    consumer_sync();
    float* const* ck = p.self_k;
    float* const* cv = p.self_v;
    if (has_new && c.tid < kHeadDim) {
        // KVCache.update (dia/state.py:99-103): append this step's K/V at `slot` of this utterance's cache
        const size_t row = ((size_t)((r & 1) * p.Hkv + kvh) * p.Lmax + slot) * kHeadDim;
        ck[u * p.L + layer][row + c.tid] = kn[c.tid];
        cv[u * p.L + layer][row + c.tid] = vn[c.tid];
    }
    float4 q[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) q[h] = reinterpret_cast<const float4*>(qs + h * kHeadDim)[c.lane];

    const int nk = w.k_hi - w.k_lo;
    const int nkc = (nk + 15) >> 4;
    const int nvc = nkc + (has_new ? 1 : 0);
    float m_run = -INFINITY, l_run = 0.f;
    float4 acc[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) acc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
    // Warp w takes the tiles w, w + 8, ...  With a ring of 8 slots (= the warps) these are successive generations of ONE
    // slot: the warp consumed the previous one itself, so the one-bit phase parity of the slot barriers is never
    // ambiguous and the warps need no barrier between the rounds - they drift apart, the slots are released and
    // refilled one by one instead of all eight at once (with the barrier a round waited for a whole ring of copies,
    // then computed, then waited again: 37 GB/s per CTA).  Any other ring size needs the round barrier.
    constexpr bool kRoundBarrier = (kBNumSlots % kConsumerWarps) != 0;
#pragma unroll 1
    for (int c0 = 0; c0 < nvc; c0 += kConsumerWarps) {
        if (kRoundBarrier && c0 > 0) consumer_sync();
        const int ci = c0 + c.warp;
        if (ci >= nvc) continue;
        const bool in_ring = ci < nkc;
        unsigned sl = 0;
        const float4* kt = reinterpret_cast<const float4*>(kn) + c.lane;
        const float4* vt = reinterpret_cast<const float4*>(vn) + c.lane;
        int keys_in = 1;
        if (in_ring) {
            const unsigned idx = c.cbase + ci;
            sl = idx % kBNumSlots;
            ring_wait_full_b(c.misc, idx, p.err, (c.seq << 8) | 0x40 | sl);
            kt = reinterpret_cast<const float4*>(c.ring + sl * kBSlotBytes) + c.lane;
            vt = reinterpret_cast<const float4*>(c.ring + sl * kBSlotBytes + kBSlotBytes / 2) + c.lane;
            keys_in = min(16, nk - ci * 16);
        }
        float s2[2];
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float v[8 * NH];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 kv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (half * 8 + i < keys_in) kv = kt[(half * 8 + i) * 32];
#pragma unroll
                for (int h = 0; h < NH; ++h)
                    v[i * NH + h] = kv.x * q[h].x + kv.y * q[h].y + kv.z * q[h].z + kv.w * q[h].w;
            }
            transpose_reduce_b<8 * NH>(v, c.lane);
            const float sv = (half * 8 + (c.lane >> 2)) < keys_in ? v[0] : -INFINITY;
            if (half == 0) s2[0] = sv; else s2[1] = sv;
        }
        float mt = fmaxf(s2[0], s2[1]);
        mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 4));
        mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 8));
        mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 16));
        const float m_new = fmaxf(m_run, mt);
        const float scale = expf(m_run - m_new);
        const float p0 = expf(s2[0] - m_new), p1 = expf(s2[1] - m_new);
        float lt = p0 + p1;
        lt += __shfl_xor_sync(0xffffffffu, lt, 4);
        lt += __shfl_xor_sync(0xffffffffu, lt, 8);
        lt += __shfl_xor_sync(0xffffffffu, lt, 16);
        l_run = l_run * scale + lt;
        m_run = m_new;
        psm[c.lane] = p0;
        psm[32 + c.lane] = p1;
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const float sh = __shfl_sync(0xffffffffu, scale, h);
            acc[h].x *= sh; acc[h].y *= sh; acc[h].z *= sh; acc[h].w *= sh;
        }
        __syncwarp();
#pragma unroll 2
        for (int key = 0; key < keys_in; ++key) {
            const float4 vv = vt[key * 32];
            const float4 pr = *reinterpret_cast<const float4*>(psm + key * HPKB);
            const float prh[4] = {pr.x, pr.y, pr.z, pr.w};
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                acc[h].x = fmaf(prh[h], vv.x, acc[h].x); acc[h].y = fmaf(prh[h], vv.y, acc[h].y);
                acc[h].z = fmaf(prh[h], vv.z, acc[h].z); acc[h].w = fmaf(prh[h], vv.w, acc[h].w);
            }
        }
        __syncwarp();
        if (in_ring && c.lane == 0) mbar_arrive(&c.misc->empty[sl]);
    }
    c.cbase += nkc;
    if (c.lane < HPKB) { wstat[c.warp * 8 + c.lane] = m_run; wstat[c.warp * 8 + 4 + c.lane] = l_run; }
#pragma unroll
    for (int h = 0; h < NH; ++h)
        reinterpret_cast<float4*>(racc + ((size_t)c.warp * HPKB + h) * kHeadDim)[c.lane] = acc[h];
    consumer_sync();

    unsigned char* oparts = self ? p.act_attn : p.act_cattn;
    u64* part = (self ? p.ll_sa_part : p.ll_ca_part) + ((size_t)w.pair * nsplit) * (nh * 132);
#pragma unroll 1
    for (int i = c.tid; i < nh * kHeadDim; i += kConsumerThreads) {
        const int h = i >> 7, d = i & 127;
        float M = -INFINITY;
#pragma unroll
        for (int ww = 0; ww < kConsumerWarps; ++ww) M = fmaxf(M, wstat[ww * 8 + h]);
        float o = 0.f, l = 0.f;
#pragma unroll
        for (int ww = 0; ww < kConsumerWarps; ++ww) {
            const float f = expf(wstat[ww * 8 + h] - M);
            l = fmaf(wstat[ww * 8 + 4 + h], f, l);
            o = fmaf(racc[((size_t)ww * HPKB + h) * kHeadDim + d], f, o);
        }
        if (w.n_active == 1) {
            const float val = l > 0.f ? o / l : 0.f;
            const int k = (head0 + h) * kHeadDim + d;
            st_act(oparts, k, r, val);                  // (the unconditional row of cross-attention attends nothing: its
                                                        // elements stay the zeros the launch starts with)
        } else {
            u64* pp = part + ((size_t)w.split * nh + h) * 132;
            ll_st(pp + 4 + d, __float_as_uint(o), c.seq);
            if (d == 0) { ll_st(pp, __float_as_uint(M), c.seq); ll_st(pp + 1, __float_as_uint(l), c.seq); }
        }
    }
    consumer_sync();
    if (w.n_active == 1) {
        if (has_new && c.tid < kHeadDim) __threadfence();
        return;
    }
    // ---- every split combines a slice of the outputs from all splits' partials (fixed order) ----------------------------------
    const int E = nh * kHeadDim;
    const int per = (E + w.n_active - 1) / w.n_active;
    const int e0 = w.split * per, e1 = min(E, e0 + per);
    if (e0 >= e1) {
        if (has_new && c.tid < kHeadDim) __threadfence();
        return;
    }
    const int ne = e1 - e0, na = w.n_active;
    const int h_lo = e0 >> 7, nhh = ((e1 - 1) >> 7) - h_lo + 1;
    float* cml = cw + ne * na;
    const int n_items = ne * na + nhh * na * 2;
    auto item_src = [&](int i) -> const u64* {
        if (i < ne * na) {
            const int e = e0 + i / na, s = i - (i / na) * na;
            return part + ((size_t)s * nh + (e >> 7)) * 132 + 4 + (e & 127);
        }
        const int j = i - ne * na;
        const int hh = j / (na * 2), rem = j - hh * (na * 2);
        return part + ((size_t)(rem >> 1) * nh + h_lo + hh) * 132 + (rem & 1);
    };
#pragma unroll 1
    for (int i = c.tid; i < n_items; i += kConsumerThreads)
        cw[i] = __uint_as_float(ll_wait32(item_src(i), c.seq, p.err));
    consumer_sync();
    float* cf = cml + nhh * na * 2;
    if (c.warp < nhh) {
        const float* ml = cml + (size_t)c.warp * na * 2;
        const float m = c.lane < na ? ml[2 * c.lane] : -INFINITY;
        const float l = c.lane < na ? ml[2 * c.lane + 1] : 0.f;
        const float M = warp_max(m);
        const float f = c.lane < na ? expf(m - M) : 0.f;
        const float Ls = warp_sum(l * f);
        if (c.lane < na) cf[c.warp * na + c.lane] = f;
        if (c.lane == 0) cf[nhh * na + c.warp] = Ls;
    }
    consumer_sync();
#pragma unroll 1
    for (int i = c.tid; i < ne; i += kConsumerThreads) {
        const int e = e0 + i, h = e >> 7, d = e & 127;
        const float* fw = cf + (size_t)(h - h_lo) * na;
        float O = 0.f;
#pragma unroll 1
        for (int s2 = 0; s2 < na; ++s2) O = fmaf(cw[i * na + s2], fw[s2], O);
        const float Lsum = cf[nhh * na + (h - h_lo)];
        const float val = Lsum > 0.f ? O / Lsum : 0.f;
        const int k = (head0 + h) * kHeadDim + d;
        st_act(oparts, k, r, val);
    }
    if (has_new && c.tid < kHeadDim) __threadfence();
    consumer_sync();
}

template <int NH>
__device__ void attn_stage_b(BCtx& c, int layer) {
    BTRACE(c, 20);
    attn_body_b<NH>(c, layer);
    BTRACE(c, 21);
    // every CTA arrives, active or not, once all its threads are done with the stage (and with its shared-memory scratch,
    // which the activation lane overwrites as soon as the next GEMM's input is complete)
    consumer_sync();
    if (c.tid == 0) act_arrive(*c.p, NH == HPKB ? A_ATTN : A_CATTN, kArrivalsPerCta);
}

// ---- residual stream entry: embedding gather-sum (dia/layers.py:691-696: x = ((e0 + e1) + e2) ... + e8) ------------------------------
__device__ void embed_stage_b(BCtx& c) {
    const BatchParams& p = *c.p;
    BMisc* misc = c.misc;
    const int gc = misc->tab.gc[G_SO], g0 = misc->tab.g0[G_SO];
    const int tid = c.tid;
    if (gc == 0) {                                           // this CTA owns no residual columns
        if (tid == 0) act_arrive(p, A_XQ, kArrivalsPerCta);
        return;
    }
    if (tid < p.U * p.C) {
        const int u = tid / p.C, ch = tid - u * p.C;
        int t;
        if (p.tokens != nullptr && c.step == 0) t = ldcg_i(p.tokens + u * p.C + ch);
        else if (c.step == 0) {
            // the rows of a finished utterance idle on token 0 (its grid has no row at pos0 - 1 any more)
            const bool done = p.utt[u].gs != nullptr && ldcg_i(&p.utt[u].gs->finished) != 0;
            t = done ? 0 : ldcg_i(p.utt[u].grid + (size_t)(min(p.utt[u].pos0, p.Lmax) - 1) * p.C + ch);
        }
        else t = (int)ll_wait32(p.ll_tok + u * DIA_B200_MAX_CHANNELS + ch, c.seq - 1, p.err);
        if (t < 0 || t >= p.V) { *p.err = kErrBadState; t = 0; }
        misc->toks[u][ch] = t;
    }
    consumer_sync();
    if ((c.warp & 3) != 0) return;
    // warp 0: utterances 0..3 (rows 0..7), warp 4: utterances 4..7 (rows 8..15) - the threads that own these elements of the
    // residual stream in the GEMM epilogues
    const int half = c.warp >> 2;
    const int m = c.lane;
    const bool valid = m < gc * 8;
    const int n = g0 * 8 + m;
    const float wnv = valid ? __ldg(p.norms + n) : 0.f;
#pragma unroll
    for (int i = 0; i < kMaxUtt / 2; ++i) {
        const int u = half * (kMaxUtt / 2) + i;
        if (u < p.U) {
            float x = 0.f;
            if (valid) {
                float e[DIA_B200_MAX_CHANNELS];
#pragma unroll
                for (int ch = 0; ch < DIA_B200_MAX_CHANNELS; ++ch)
                    if (ch < p.C) e[ch] = __ldg(p.emb + ((size_t)ch * p.V + misc->toks[u][ch]) * p.D + n);
                x = e[0];
#pragma unroll
                for (int ch = 1; ch < DIA_B200_MAX_CHANNELS; ++ch)
                    if (ch < p.C) x += e[ch];
                c.xres[2 * i] = x;
                c.xres[2 * i + 1] = x;
                st_act(p.act_xq, n, 2 * u, x * wnv);
                st_act(p.act_xq, n, 2 * u + 1, x * wnv);
            }
            float sq = x * x;
            sq += __shfl_xor_sync(0xffffffffu, sq, 4);
            sq += __shfl_xor_sync(0xffffffffu, sq, 2);
            sq += __shfl_xor_sync(0xffffffffu, sq, 1);
            if (valid && (m & 7) == 0) {
                st_gpu_f(p.ssq + (size_t)(g0 + (m >> 3)) * kRows + 2 * u, sq);
                st_gpu_f(p.ssq + (size_t)(g0 + (m >> 3)) * kRows + 2 * u + 1, sq);
            }
        }
    }
    __syncwarp();
    if (c.lane == 0) act_arrive(p, A_XQ, kArrivalsPerCta / 2);        // warps 0 and 4
}

// ---- sampling: CTA u * C + ch draws channel ch of utterance u; the CTA of channel 0 then runs the body of the reference's
//      while loop after _decoder_step (dia/model.py:771-807) for its utterance and publishes the next step's tokens -------------------
__device__ void sample_stage_b(BCtx& c) {
    const BatchParams& p = *c.p;
    const int b = blockIdx.x;
    if (b >= p.U * p.C) return;
    const int u = b / p.C, ch = b - u * p.C;
    SampleSmem* sm = reinterpret_cast<SampleSmem*>(c.scratch);
    const u64* src = p.ll_glog + ((size_t)u * p.C + ch) * p.V;
    float g[kPerThread];
    {
        uint2 wv[kPerThread];
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            const int idx = c.tid + kConsumerThreads * i;
            wv[i] = idx < p.V ? ll_ld(src + idx) : make_uint2(0xff800000u, c.seq - 1);
        }
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            if (wv[i].y != c.seq - 1) wv[i].x = ll_wait32(src + c.tid + kConsumerThreads * i, c.seq - 1, p.err);
            g[i] = __uint_as_float(wv[i].x);
        }
    }
    const UttParams& up = p.utt[u];
    const int tok = sample_channel_cta(g, p.V, p.temperature, p.top_p, p.top_k, up.seed, p.draw0 + c.step, ch, nullptr, sm, c.tid);
    u64* preds = p.ll_pred + u * DIA_B200_MAX_CHANNELS;
    if (c.tid == 0) ll_st(preds + ch, (uint32_t)tok, c.seq);
    if (ch != 0 || c.warp != 0) return;

    const int lane = c.lane;
    int pr = 0;
    if (lane < p.C) {
        pr = (int)ll_wait32(preds + lane, c.seq, p.err);
        p.pred_out[u * DIA_B200_MAX_CHANNELS + lane] = pr;
    }
    GenState* gs = up.gs;
    int next_tok = 0;
    if (gs != nullptr && up.grid != nullptr) {
        int dec_step = gs->dec_step, finished = gs->finished, eos_detected = gs->eos_detected;
        int eos_cd = gs->eos_countdown, bos_cd = gs->bos_countdown, steps_run = gs->steps_run;
        __syncwarp();
        if (!finished) {
            if (dec_step >= p.max_tokens - 1) {
                finished = 1;
            } else {
                int dmax = 0;
#pragma unroll 1
                for (int i = 0; i < p.C; ++i) dmax = max(dmax, p.delay[i]);
                const int cur = dec_step + 1;
                if (cur != up.pos0 + c.step && lane == 0) *p.err = kErrBadState;
                const int pr0 = __shfl_sync(0xffffffffu, pr, 0);
                if (!eos_detected && pr0 == p.eos) { eos_detected = 1; eos_cd = dmax; }
                if (eos_cd > 0) {
                    const int s = dmax - eos_cd;
                    if (lane < p.C) {
                        if (s == p.delay[lane]) pr = p.eos;
                        else if (s > p.delay[lane] && pr != p.eos) pr = p.pad;
                    }
                    eos_cd -= 1;
                }
                bos_cd = max(0, bos_cd - 1);
                if (lane < p.C) {
                    int* cell = up.grid + (size_t)cur * p.C + lane;
                    if (bos_cd > 0) {                                  // update_one(apply_mask=True)
                        const int old = ldcg_i(cell);
                        if (old == -1) *cell = pr; else pr = old;
                    } else {
                        *cell = pr;
                    }
                    next_tok = pr;
                }
                if (eos_cd == 0) {
                    finished = 1;                                   // break: dec_step is NOT advanced
                } else {
                    if (cur >= p.max_tokens - dmax - 1 && !eos_detected) { eos_detected = 1; eos_cd = dmax; }
                    dec_step += 1;
                    if (dec_step >= p.max_tokens - 1) finished = 1;
                }
                steps_run += 1;
            }
        }
        if (lane == 0) {
            gs->dec_step = dec_step; gs->finished = finished; gs->eos_detected = eos_detected;
            gs->eos_countdown = eos_cd; gs->bos_countdown = bos_cd; gs->steps_run = steps_run;
        }
        if (finished) next_tok = 0;                                 // the rows of a finished utterance idle on token 0
    }
    if (lane < p.C) ll_st(p.ll_tok + u * DIA_B200_MAX_CHANNELS + lane, (uint32_t)next_tok, c.seq);
}

}  // namespace

// ---- the kernel ----------------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(kBThreads, 1) dia_batch_step_kernel(const __grid_constant__ BatchParams p) {
    extern __shared__ unsigned char smem_raw_b[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw_b) + 1023) & ~(uintptr_t)1023);
    unsigned char* ring = smem;
    unsigned char* scratch = smem + kBNumSlots * kBSlotBytes;
    BMisc* misc = reinterpret_cast<BMisc*>(scratch + kScratchBytes);
    const int tid = threadIdx.x, warp = tid >> 5;

    // a launch queued behind the one that finished every utterance is a no-op (uniform: `finished` only changes at the
    // end of a step, which needs every CTA to have taken part in that step's stages)
    if (p.with_sample) {
        bool all_done = true;
        for (int u = 0; u < p.U; ++u) all_done = all_done && p.utt[u].gs != nullptr && ldcg_i(&p.utt[u].gs->finished) != 0;
        if (all_done) return;
    }
    if (tid == 0) {
        // every ring slot and activation stage has ONE consumer (the issuer whose quarter of the contraction it belongs to)
        for (int i = 0; i < kBNumSlots; ++i) { mbar_init(&misc->full[i], 1); mbar_init(&misc->empty[i], 1); }
        // a stage is filled by one bulk copy (or two multicast halves: one arrive.expect_tx for both) and released by
        // its issuer's commit - of both CTAs when a pair shares its activation stages
        for (int i = 0; i < kActStages; ++i) { mbar_init(&misc->bfull[i], 1); mbar_init(&misc->bempty[i], p.mc ? 2 : 1); }
        mbar_init(&misc->acc_full, kBIssuers);
        misc->stages_done = 0;
        misc->ready_seq = 0;
        fence_mbar_init();
    }
    {
        const int* src = reinterpret_cast<const int*>(p.cta_tab + blockIdx.x);
        int* dst = reinterpret_cast<int*>(&misc->tab);
        for (int i = tid; i < (int)(sizeof(CtaTable) / 4); i += kBThreads) dst[i] = src[i];
    }
    // (rows >= R of the activation tiles are zeros in global memory: the launch starts from a zeroed exchange region)
    if (warp == kMmaWarp) {                                  // one warp allocates (and later frees) the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&misc->tmem_base)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (p.mc) cluster_sync_all();          // the peer's barriers exist before anything is multicast to them

    const int S = 8 * p.L + 3, n_stage = p.with_sample ? S : S - 1;
    asm volatile(".reg .pred dia_pw_bfull;" ::);       // mma_loop_b: result of the early test of the next activation stage's barrier
    if (warp == kProducerWarp) {
        if (tid == kProducerWarp * 32) producer_loop_b(p, ring, misc);
    } else if (warp == kActWarp) {
        if (tid == kActWarp * 32) act_loop_b(p, scratch, misc);
    } else if (warp < kConsumerWarps) {
        BCtx c;
        c.p = &p; c.ring = ring; c.scratch = scratch; c.misc = misc;
        c.tid = tid; c.warp = warp; c.lane = tid & 31;
        c.cbase = 0; c.bctr = 0; c.gctr = 0; c.seq = 0; c.step = 0;
        c.prof = p.prof != nullptr && blockIdx.x == 0 && tid == 0;
        c.prof_issue = p.prof != nullptr && blockIdx.x == 0 && tid == kIssuerWarp0 * 32;
        for (int i = 0; i < 3; ++i) c.t_issue[i] = 0;
        c.t_tmem = 0;
        for (int i = 0; i < 8; ++i) c.t_prof[i] = 0;
        const long long t_begin = c.prof ? clock64() : 0;
#pragma unroll
        for (int r = 0; r < kRows / 2; ++r) c.xres[r] = 0.f;
#pragma unroll 1
        for (int n = 0; n < p.n_steps; ++n) {
            c.step = n;
#pragma unroll 1
            for (int s = 0; s < n_stage; ++s) {
                int kind, layer;
                decode_stage_b(s, p.L, kind, layer);
                c.seq = 1u + (unsigned)(n * S + s);
                const long long ts0 = c.prof ? clock64() : 0;
                switch (kind) {
                    case S_EMBED: embed_stage_b(c); break;
                    case S_SATTN: attn_stage_b<HPKB>(c, layer); break;
                    case S_CATTN: attn_stage_b<1>(c, layer); break;
                    case S_SAMPLE: sample_stage_b(c); break;
                    default: gemm_stage_b(c, gemm_of_kind_b(kind), layer); break;
                }
                if (c.prof) {
                    if (kind == S_SATTN || kind == S_CATTN) c.t_prof[6] += clock64() - ts0;
                    else if (kind == S_EMBED || kind == S_SAMPLE) c.t_prof[7] += clock64() - ts0;
                }
                if (tid == 0) st_release_cta_s32(&misc->stages_done, n * S + s + 1);
            }
        }
        if (c.prof_issue) {
            p.prof[0] = (unsigned long long)c.t_issue[0];
            p.prof[1] = (unsigned long long)c.t_issue[1];
            p.prof[2] = (unsigned long long)c.t_issue[2];
        }
        if (c.prof) {
            p.prof[8] = (unsigned long long)(clock64() - t_begin);
            for (int i = 0; i < 8; ++i) p.prof[9 + i] = (unsigned long long)c.t_prof[i];
            p.prof[17] = (unsigned long long)c.t_tmem;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (p.mc) cluster_sync_all();          // no CTA leaves while its peer may still multicast into its shared memory
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(misc->tmem_base), "r"(kTmemCols));
    }
}

cudaError_t launch_batch_kernel(const BatchParams& p, cudaStream_t st) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        e = cudaFuncSetAttribute(dia_batch_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBSmem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.G);
    cfg.blockDim = dim3(kBThreads);
    cfg.dynamicSmemBytes = kBSmem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = 2; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = p.mc ? 2 : 1;           // CTA pairs (the two SMs of a TPC) when the pair shares its activation stages
    return cudaLaunchKernelEx(&cfg, dia_batch_step_kernel, p);
}

// carve the exchange region of the batched kernel (`base` may be null to only size it); returns bytes
size_t batch_ll_layout(const BatchParams& g, BatchParams* out, unsigned char* base) {
    size_t off = 0;
    auto take = [&](void** ptr, size_t bytes) {
        if (out) *ptr = base ? base + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
    };
    BatchParams scratch;
    BatchParams& o = out ? *out : scratch;
    const size_t R = (size_t)g.R;
    const size_t nq = (size_t)g.Hq * kHeadDim, nkv = (size_t)g.Hkv * kHeadDim, nc = (size_t)g.Hc * kHeadDim;
    auto act_bytes = [](size_t K) { return (K / 64) * (size_t)(2 * kBTermBytes); };
    take(reinterpret_cast<void**>(&o.act_xq), act_bytes(g.D));
    take(reinterpret_cast<void**>(&o.act_xc), act_bytes(g.D));
    take(reinterpret_cast<void**>(&o.act_xm), act_bytes(g.D));
    take(reinterpret_cast<void**>(&o.act_xl), act_bytes(g.D));
    take(reinterpret_cast<void**>(&o.act_attn), act_bytes(nq));
    take(reinterpret_cast<void**>(&o.act_cattn), act_bytes(nc));
    take(reinterpret_cast<void**>(&o.act_hidden), act_bytes(g.F));
    take(reinterpret_cast<void**>(&o.ctr), (size_t)A_COUNT * 32 * 4);
    take(reinterpret_cast<void**>(&o.ssq), (size_t)(g.D / 8) * kRows * 4);
    take(reinterpret_cast<void**>(&o.ll_qkv), (nq + 2 * nkv) * R * 8);
    take(reinterpret_cast<void**>(&o.ll_cq), nc * R * 8);
    take(reinterpret_cast<void**>(&o.ll_sa_part), (size_t)g.G * 4 * 132 * 8);
    take(reinterpret_cast<void**>(&o.ll_ca_part), (size_t)g.G * 132 * 8);
    take(reinterpret_cast<void**>(&o.ll_glog), (size_t)g.U * g.C * g.V * 8);
    take(reinterpret_cast<void**>(&o.ll_pred), (size_t)kMaxUtt * DIA_B200_MAX_CHANNELS * 8);
    take(reinterpret_cast<void**>(&o.ll_tok), (size_t)kMaxUtt * DIA_B200_MAX_CHANNELS * 8);
    return off;
}

}  // namespace dia
