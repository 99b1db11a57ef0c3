// step_kernel.cu - the persistent Dia decode-step kernel for sm_100a.
//
// One CTA per SM (148 on B200), launched cooperatively (co-residency is required: CTAs wait on each
// other's data).  A decode step is a chain of ~147 strictly dependent stages (embed, 8 per decoder
// layer, logits, sample) over a batch of 2 rows (CFG uncond/cond) - 2 FLOP per weight byte, i.e. bound
// by streaming 2.53 GB of bf16 weights from HBM3e, and after that by the LATENCY of the stage chain
// (386 us of HBM time per step leaves 2.6 us per stage).  The design follows from those two facts:
//
//  * every CTA owns a fixed slice of the output columns of every GEMM, and the host repacks the
//    weights ONCE into one contiguous bf16 byte stream per CTA in exactly the order that CTA
//    consumes them;
//  * a producer warp (one elected lane) walks that stream with 1-D bulk async copies
//    (cp.async.bulk, the TMA engine) into a 20 x 8 KB shared-memory ring guarded by full/empty
//    mbarriers.  It never waits for the math warps' dependencies, so it keeps prefetching the NEXT
//    stages' weights while they sit in a latency-bound stage - HBM stays busy across stage boundaries;
//  * there is NO grid barrier.  Every vector that crosses CTAs (normalised residual stream, q/k/v,
//    attention outputs, MLP hidden, split-KV partials, logits, tokens) is written as 8-byte words that
//    carry their own sequence flag ("LL" words: one strong store, one strong load, the reader re-polls
//    until the flag is the producing stage's number).  A stage boundary therefore costs one L2 write
//    plus one L2 read instead of fence + atomic + poll + read;
//  * the residual stream never leaves the SM: the three residual GEMMs (self-o, cross-o, mlp-out) share
//    one column partition, so each element of x lives in a register of the thread that owns its column;
//  * 8 math warps consume ring slots (one slot per warp at a time) on the tensor cores (mma.sync
//    m16n8k16, swap-AB: 16 output columns x the 2 batch rows split into three bf16 terms, exact to
//    24 bits, so the product has fp32-activation accuracy - SURVEY.md 8(c)); residual stream, norms,
//    softmax and the KV cache stay fp32, reductions run in a fixed order (deterministic).  2:4-pruned
//    checkpoints stream compressed slabs (2 of every 4 K entries + metadata, 0.5625 of the bytes) and
//    use mma.sp m16n8k32 on the same operand plumbing;
//  * a norm warp gathers the per-CTA partial sums of x^2 behind every RMSNorm while the math warps run
//    their MMAs, and hands 1/rms to the epilogue through a shared-memory flag;
//  * the code is deliberately compact and rolled: the L1.5 instruction cache is 32 KB and every stage
//    runs once per layer, so straight-line code is fetched cold from L2 at ~30 cycles/instruction;
//  * self-attention streams its K/V tiles through the same ring (split-KV over CTAs, 4 query heads
//    share each KV tile, RoPE fused, K/V append fused), the splits exchange partials as LL words and
//    each combines a slice; cross-attention reads only the valid keys of the conditional row;
//  * sampling runs on 9 CTAs (one codebook each: CFG combine fused into the logits epilogue, radix
//    select for top-k, top-p, Philox draw), the EOS state machine on CTA 0; a whole run of steps needs
//    no host sync.
//
// Reference semantics: dia/layers.py:671-720 (decode_step), :530-584 (DecoderLayer),
// :238-346 (Attention), :92-105 (MlpBlock); dia/model.py:429-488, 32-82, 748-807.
#include <cuda_bf16.h>

#include "common.cuh"
#include "engine_internal.h"
#include "sampler.cuh"

namespace dia {

typedef unsigned long long u64;

// per-GEMM constants of this CTA, computed once per launch (no integer divisions in the stage loop)
struct GemmCfg {
    int gc, g0, K;
    int row_bytes;     // gc * 16
    int rpc;           // rows of the slab per ring slot (gemm_slot_rows)
    int n_chunks;      // ring slots of the slab = K / rpc
    int sl;            // contraction rows owned by one warp = K / 8
    int urows;         // rows of one input fetch unit = min(sl, 256)
    int cpu;           // ring slots per fetch unit = urows / rpc
    int n_mt;          // 16-column MMA tiles = ceil(gc / 2)
    int slot_bytes;    // bytes of one ring slot of the slab (dense: rpc * row_bytes; 2:4: values + metadata)
    int meta_off;      // 2:4: byte offset of the metadata inside a slot = (rpc / 2) * row_bytes
};

struct SharedMisc {
    uint64_t full[kNumSlots];
    uint64_t empty[kNumSlots];
    CtaTable tab;
    GemmCfg gcfg[G_COUNT];
    float inv_rms[2][2];                 // [stage parity][batch row]: 1/rms of the stage input
    unsigned inv_seq[2];                 // stage sequence number for which inv_rms[parity] is valid
    float stat[16];
    int bcast[8];
    int stages_done;                     // consumer -> producer progress (global stage index + 1)
};
static_assert(sizeof(SharedMisc) <= kMiscBytes, "misc region too small");

struct Ctx {
    const StepParams* p;
    unsigned char* ring;
    unsigned char* xs;
    float* red;
    SharedMisc* misc;
    int tid, warp, lane;
    unsigned cbase;      // ring chunk index at the start of the current stage
    unsigned seq;        // sequence number of the current stage inside this launch (>= 1)
    float xres;          // warp 0: this lane's element of the residual stream (lane = row * 16 + column)
    long long* ts;       // debug: 8 clock64 slots of the current stage (CTA 0, thread 0 only)
};

// A math warp waits for generation g of a ring slot.  Successive generations of one slot are consumed by
// DIFFERENT warps, so the waiting warp has not observed generation g-1 complete and the one-bit phase parity
// would alias (a wait for g passes while g-1 is still in flight).  Waiting first for the RELEASE of g-1 (the
// empty barrier, which needs g-1 filled and consumed) pins the full barrier to generation g.
__device__ __forceinline__ void ring_wait_full(SharedMisc* misc, unsigned slot, unsigned parity, int* err, unsigned info) {
    mbar_wait(&misc->empty[slot], parity ^ 1u, err, kErrEmptyBarrierTimeout, info);
    mbar_wait(&misc->full[slot], parity, err, kErrFullBarrierTimeout, info);
}

__device__ __forceinline__ void decode_stage(int s, int L, int& kind, int& layer) {
    if (s == 0) { kind = S_EMBED; layer = 0; }
    else if (s <= 8 * L) { layer = (s - 1) >> 3; kind = S_QKV + ((s - 1) & 7); }
    else if (s == 8 * L + 1) { kind = S_LOGITS; layer = 0; }
    else { kind = S_SAMPLE; layer = 0; }
}
__device__ __forceinline__ int gemm_of_kind(int kind) {
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

struct AttnWork {
    int active, pair, split, k_lo, k_hi, n_active, has_new;
};
// self-attention: (row, kv head) pairs x key splits over the CTAs.  Old keys are cache
// slots [0, slot); the key/value of THIS step is taken from the qkv words by the last active split.
__device__ __forceinline__ AttnWork self_attn_work(const StepParams& p, int cta, int slot) {
    AttnWork w;
    const int nsplit = p.sa_nsplit, pairs = 2 * p.Hkv, n_old = slot;
    w.pair = cta / nsplit;
    w.split = cta - w.pair * nsplit;
    int per = (n_old + nsplit - 1) / nsplit;
    per = (per + 15) & ~15;
    if (per < 64) per = 64;
    w.n_active = (n_old + per - 1) / per;
    if (w.n_active < 1) w.n_active = 1;
    w.active = (w.pair < pairs) && (w.split < w.n_active);
    w.k_lo = w.split * per;
    w.k_hi = min(n_old, w.k_lo + per);
    if (w.k_hi < w.k_lo) w.k_hi = w.k_lo;
    w.has_new = (w.split == w.n_active - 1);      // the last split has the fewest old keys (per is rounded up)
    return w;
}
// cross-attention: conditional row only, one query head per KV head, keys [0, text_len); a CTA takes
// at least 128 keys so that typical transcripts need no split at all
__device__ __forceinline__ AttnWork cross_attn_work(const StepParams& p, int cta) {
    AttnWork w;
    const int nsplit = p.ca_nsplit, n = p.text_len;
    w.pair = cta / nsplit;
    w.split = cta - w.pair * nsplit;
    int per = (n + nsplit - 1) / nsplit;
    per = (per + 15) & ~15;
    if (per < 128) per = 128;
    w.n_active = (n + per - 1) / per;
    if (w.n_active < 1) w.n_active = 1;
    w.active = (w.pair < p.Hc) && (w.split < w.n_active);
    w.k_lo = w.split * per;
    w.k_hi = min(n, w.k_lo + per);
    if (w.k_hi < w.k_lo) w.k_hi = w.k_lo;
    w.has_new = 0;
    return w;
}

// ------------------------------------------------------------------------------------------
// producer: walks this CTA's byte stream
// ------------------------------------------------------------------------------------------
struct Producer {
    unsigned char* ring;
    SharedMisc* misc;
    int* err;
    unsigned pc;
    unsigned seq;
    uint64_t pol_stream, pol_keep;
    __device__ __forceinline__ void issue(const void* src, uint32_t bytes, bool keep) {
        const unsigned slot = pc % kNumSlots;
        const unsigned ph = (pc / kNumSlots) & 1u;
        mbar_wait(&misc->empty[slot], ph ^ 1u, err, kErrEmptyBarrierTimeout, (seq << 8) | slot);
        mbar_arrive_expect_tx(&misc->full[slot], bytes);
        bulk_g2s_hint(ring + slot * kSlotBytes, src, bytes, &misc->full[slot], keep ? pol_keep : pol_stream);
        pc++;
    }
};

__device__ void producer_loop(const StepParams& p, unsigned char* ring, SharedMisc* misc) {
    Producer pr;
    pr.ring = ring; pr.misc = misc; pr.err = p.err; pr.pc = 0;
    pr.pol_stream = l2_policy_evict_first();
    pr.pol_keep = l2_policy_evict_last();
    const CtaTable& tab = misc->tab;
    const int cta = blockIdx.x;
    const int S = 8 * p.L + 3;
#pragma unroll 1
    for (int n = 0; n < p.n_steps; ++n) {
        const int slot = p.slot0 + n;
#pragma unroll 1
        for (int s = p.stage_begin; s < p.stage_end; ++s) {
            int kind, layer;
            decode_stage(s, p.L, kind, layer);
            pr.seq = 1u + (unsigned)(n * S + s);
            const int gt = gemm_of_kind(kind);
            if (gt >= 0) {
                const GemmCfg& g = misc->gcfg[gt];
                if (g.gc == 0) continue;
                const unsigned char* base = p.wstream + tab.stream_base +
                    (gt == G_LOGITS ? tab.logits_off
                                    : (unsigned long long)layer * tab.layer_bytes + tab.slab_off[gt]);
#pragma unroll 1
                for (int ci = 0; ci < g.n_chunks; ++ci)
                    pr.issue(base + (size_t)ci * g.slot_bytes, g.slot_bytes, false);
            } else if (kind == S_SATTN || kind == S_CATTN) {
                const bool self = kind == S_SATTN;
                const AttnWork w = self ? self_attn_work(p, cta, slot) : cross_attn_work(p, cta);
                if (!w.active || w.k_hi <= w.k_lo) continue;
                if (self && n > 0) {
                    // row slot-1 was appended by another CTA during this stage of step n-1; that CTA fenced the
                    // append before its self-o epilogue, and this CTA's math warps consumed every self-o output
                    // once they are past the cross-q stage of step n-1.  The ring holds a fraction of one layer,
                    // so this never blocks.
                    const int need = (n - 1) * S + s + 3;
                    if (ld_acquire_cta_s32(&misc->stages_done) < need) {
                        const unsigned long long t0 = clock64();
                        while (ld_acquire_cta_s32(&misc->stages_done) < need) {
                            if (clock64() - t0 > kWatchdogCycles) ll_timeout(p.err, kErrStepDoneTimeout);
                        }
                    }
                    fence_proxy_async();
                }
                const float* kb;
                const float* vb;
                if (self) {
                    const size_t off = ((size_t)w.pair * p.Lmax) * kHeadDim;      // pair = row*Hkv + kvh
                    kb = p.self_k[layer] + off;
                    vb = p.self_v[layer] + off;
                } else {
                    const size_t off = ((size_t)(p.Hc + w.pair) * p.Smax) * kHeadDim;   // row 1 (cond)
                    kb = p.cross_k[layer] + off;
                    vb = p.cross_v[layer] + off;
                }
#pragma unroll 1
                for (int k0 = w.k_lo; k0 < w.k_hi; k0 += 16) {        // K tile, then the V tile of the same keys
                    const int nk = min(16, w.k_hi - k0);
                    pr.issue(kb + (size_t)k0 * kHeadDim, nk * kHeadDim * 4, !self);
                    pr.issue(vb + (size_t)k0 * kHeadDim, nk * kHeadDim * 4, !self);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// norm warp: the one all-CTA dependency of a normed projection (qkv, cross-q, mlp-in, logits) is 1/rms of its input,
// i.e. the per-CTA partial sums of x^2 that the producing residual stage published next to x.  A warp of its own
// gathers them (fixed order), while the math warps are in the MMA loop, and hands 1/rms over through a
// shared-memory flag that the epilogue waits on.
// ------------------------------------------------------------------------------------------
constexpr int kSsqPerLane = 5;                     // <= 160 residual-owning CTAs

__device__ void norm_warp_loop(const StepParams& p, SharedMisc* misc) {
    const int lane = threadIdx.x & 31;
    const int S = 8 * p.L + 3;
#pragma unroll 1
    for (int n = 0; n < p.n_steps; ++n) {
#pragma unroll 1
        for (int s = p.stage_begin; s < p.stage_end; ++s) {
            int kind, layer;
            decode_stage(s, p.L, kind, layer);
            if (kind != S_QKV && kind != S_CQ && kind != S_WI && kind != S_LOGITS) continue;
            if (misc->gcfg[gemm_of_kind(kind)].gc == 0) continue;         // this CTA has no columns of the GEMM
            const unsigned seq = 1u + (unsigned)(n * S + s), fprev = seq - 1;
            // start polling when this CTA's math warps enter the stage: the partials appear together with x
            if (s > p.stage_begin || n > 0) {
                const int need = s > p.stage_begin ? n * S + s : (n - 1) * S + p.stage_end;
                // sleep between polls: a busy spin here took a tenth of the SM's issue slots, on the scheduler that
                // also runs warp 0 (the residual-stream epilogues)
                unsigned polls = 0;
                while (ld_acquire_cta_s32(&misc->stages_done) < need) {
                    __nanosleep(200);
                    if (++polls > 20000000u) ll_timeout(p.err, kErrStepDoneTimeout, seq);
                }
            }
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int i = 0; i < kSsqPerLane; ++i) {
                const int j = lane + 32 * i;
                if (j < p.n_res) {
                    uint4 q4 = ll_ld2(p.ll_ssq + 2 * j);
                    if (q4.y != fprev) q4.x = ll_wait32(p.ll_ssq + 2 * j, fprev, p.err);
                    if (q4.w != fprev) q4.z = ll_wait32(p.ll_ssq + 2 * j + 1, fprev, p.err);
                    s0 += __uint_as_float(q4.x);
                    s1 += __uint_as_float(q4.z);
                }
            }
            s0 = warp_sum(s0);
            s1 = warp_sum(s1);
            if (lane == 0) {   // torch.nn.RMSNorm: x * rsqrt(mean(x^2) + eps) * w
                const int par = seq & 1;
                volatile float* ir = misc->inv_rms[par];
                ir[0] = 1.0f / sqrtf(s0 / (float)p.D + p.eps);
                ir[1] = 1.0f / sqrtf(s1 / (float)p.D + p.eps);
                __threadfence_block();
                *reinterpret_cast<volatile unsigned*>(&misc->inv_seq[par]) = seq;
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------
// activation vectors as LL words
// ------------------------------------------------------------------------------------------
// x = hi + lo + lo2 with each term bf16 (round-to-nearest at every level; the residuals are exact)
__device__ __forceinline__ void ll_store_parts(u64* dst, int k, int r, float x, uint32_t f16) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h);
    const __nv_bfloat16 l = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(l);
    const __nv_bfloat16 l2 = __float2bfloat16_rn(r2);
    const uint32_t lo = (uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(l) << 16);
    const uint32_t hi = (uint32_t)__bfloat16_as_ushort(l2) | (f16 << 16);
    ll_st(dst + (size_t)(k >> 4) * 32 + r * 16 + (k & 15), lo, hi);
}

// position of input element k in the (K-row compacted) contraction of GEMM family gt in `layer`; -1 = dropped
__device__ __forceinline__ int row_pos(const StepParams& p, int gt, int layer, int k) {
    const int* m = p.rowmap[gt];
    return m == nullptr ? k : __ldg(m + (size_t)(gt == G_LOGITS ? 0 : layer) * p.Kfull[gt] + k);
}

// ------------------------------------------------------------------------------------------
// GEMM stage: y[2][N_cta] = x[2][K] . W_slab on the tensor cores (mma.sync m16n8k16, bf16 x bf16 ->
// fp32).  Swap-AB: the 16 rows of the MMA are 16 OUTPUT COLUMNS (two 8-column groups of the slab,
// fed from the ring with ldmatrix.trans - the slab is [k][n], n contiguous), the 8 columns of the MMA
// carry the two batch rows, each split into three bf16 terms  x = hi + lo + lo2:
//   MMA column n: 0,1,2 = row 0 (hi, lo, lo2)   4,5,6 = row 1 (hi, lo, lo2)   3,7 = zero
// The k index of the MMA is a free permutation as long as A and B agree; it is chosen so that the four
// k values of lane q are the four CONSECUTIVE slab rows 4q..4q+3 (see lm_off below).
// The input vector arrives as LL words [k/16][row][16] = (hi, lo, lo2, flag16) written by the epilogue
// that produced it (pre-multiplied by the consumer's RMSNorm weight).  A warp fetches the words of the
// k-range of its ring slot with up to 8 16-byte loads per lane, re-polls until every flag carries the
// producing stage's number, and transposes them into B fragments in its private staging buffer.  The
// 1/rms factor commutes with the GEMM and is applied in the epilogue from per-CTA partial sums.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t smem_addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_addr)
                 : "memory");
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

constexpr int kMaxTiles = 8;     // <= 16 column groups per CTA and GEMM
constexpr int kMaxKb = 16;       // k-blocks per ring slot (slot rows are capped at 256)
constexpr int kLLW = 8;          // 16-byte LL loads per lane for one slot: 16 k-blocks x 2 rows x 16 words / 2 / 32

// the LL words of rows [row0, row0 + 16 * nkb) of `src`: lane takes the 16-byte pairs lane, lane + 32, ...
__device__ __forceinline__ void ll_fetch(uint4 (&w)[kLLW], const u64* src, int row0, int nkb, int lane, uint32_t f16) {
    const unsigned char* base = reinterpret_cast<const unsigned char*>(src) + (size_t)row0 * 16 + lane * 16;
    const int pairs = nkb * 16;
#pragma unroll
    for (int i = 0; i < kLLW; ++i) {
        if (lane + 32 * i < pairs) w[i] = ll_ld2(base + i * 512);
        else w[i] = make_uint4(0u, f16 << 16, 0u, f16 << 16);
    }
}
__device__ __forceinline__ bool ll_stale(const uint4 (&w)[kLLW], uint32_t f16) {
    uint32_t bad = 0;
#pragma unroll
    for (int i = 0; i < kLLW; ++i) bad |= ((w[i].y >> 16) ^ f16) | ((w[i].w >> 16) ^ f16);
    return bad != 0;
}

// The MMAs of one ring slot for a compile-time tile count NT (a CTA with fewer real tiles computes the padding
// tiles on whatever follows its columns in the slot; their accumulators are never read).  Few tiles -> several
// k-blocks in flight on independent accumulator chains (the legacy HMMA path has ~21 cycles of latency).
template <int NT>
__device__ __forceinline__ void mma_chunk(float (&acc)[kMaxTiles][4], uint32_t a_addr, uint32_t kb_bytes, int nkb,
                                          uint32_t bst, uint32_t b_off0, uint32_t b_off1, uint32_t bmask) {
    constexpr int U = NT == 1 ? 4 : (NT == 2 ? 2 : 1);       // the host makes every slot a multiple of U k-blocks
#pragma unroll 1
    for (int kb = 0; kb < nkb; kb += U) {
        uint2 b[U];
        uint32_t a[U][NT][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {                       // all operand loads first: they are independent
            const int k = kb + u;
            const uint32_t boff = U == 1 ? ((k & 1) ? b_off1 : b_off0) : ((u & 1) ? b_off1 : b_off0);
            b[u] = lds_u2(bst + (k >> 1) * 512 + boff);
#pragma unroll
            for (int mt = 0; mt < NT; ++mt) ldmatrix_x4_trans(a[u][mt], a_addr + k * kb_bytes + mt * 32);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            b[u].x &= bmask; b[u].y &= bmask;
#pragma unroll
            for (int mt = 0; mt < NT; ++mt) mma_bf16_16816(acc[U > 1 ? u * NT + mt : mt], a[u][mt], b[u].x, b[u].y);
        }
    }
}

// 2:4 weights: mma.sp m16n8k32 - A = 16 output columns x 32 contraction rows held as the 16 compressed rows of the
// block (same ldmatrix pattern as the dense path on half the rows), B = the input fragments of two dense k-blocks,
// metadata = one word per lane (lanes with lane % 4 < 2, selector 0) from the slot's metadata area.
__device__ __forceinline__ void mma_sp_bf16_16832(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1,
                                                  uint32_t b2, uint32_t b3, uint32_t e) {
    asm("mma.sp::ordered_metadata.sync.aligned.m16n8k32.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
        "{%8, %9, %10, %11}, {%0, %1, %2, %3}, %12, 0x0;"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(b2), "r"(b3), "r"(e));
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
template <int NT>
__device__ __forceinline__ void mma_chunk_sp(float (&acc)[kMaxTiles][4], uint32_t a_addr, uint32_t kb_bytes, int n32,
                                             uint32_t bst, uint32_t b_off0, uint32_t b_off1, uint32_t bmask,
                                             uint32_t meta_addr, int n_mt) {
    constexpr int U = NT == 1 ? 2 : 1;                       // the host makes every slot a multiple of U 32-row blocks
#pragma unroll 1
    for (int jb = 0; jb < n32; jb += U) {
        uint2 blo[U], bhi[U];
        uint32_t a[U][NT][4], e[U][NT];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = jb + u;
            blo[u] = lds_u2(bst + j * 512 + b_off0);
            bhi[u] = lds_u2(bst + j * 512 + b_off1);
#pragma unroll
            for (int mt = 0; mt < NT; ++mt) {
                ldmatrix_x4_trans(a[u][mt], a_addr + j * kb_bytes + mt * 32);
                e[u][mt] = lds_u32(meta_addr + (uint32_t)(j * n_mt + mt) * 64u);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            blo[u].x &= bmask; blo[u].y &= bmask; bhi[u].x &= bmask; bhi[u].y &= bmask;
#pragma unroll
            for (int mt = 0; mt < NT; ++mt)
                mma_sp_bf16_16832(acc[U > 1 ? u * NT + mt : mt], a[u][mt], blo[u].x, blo[u].y, bhi[u].x, bhi[u].y, e[u][mt]);
        }
    }
}

__device__ void gemm_stage(Ctx& c, int gt, int layer) {
    const StepParams& p = *c.p;
    const GemmCfg& g = c.misc->gcfg[gt];
    const int gc = g.gc, g0 = g.g0, n_mt = g.n_mt;
    if (gc == 0) return;
    const int lane = c.lane, warp = c.warp;
    const bool resid = (gt == G_SO || gt == G_CO || gt == G_WO);
    const bool normed = !resid;
    const uint32_t fprev = c.seq - 1;                  // the producing stage
    const uint32_t f16 = fprev & 0xffffu;
    const int par = c.seq & 1;
    float* red = c.red + par * (kRedBytes / 4);

    const u64* src = gt == G_SO ? p.ll_attn : gt == G_CO ? p.ll_cattn : gt == G_WO ? p.ll_hidden : p.ll_x;
    const int n_chunks = g.n_chunks, rpc = g.rpc;

    uint4 w[kLLW];

    // epilogue role of this thread: (tile, row, column-in-tile)
    const int e_mt = warp, e_r = lane >> 4, e_m = lane & 15;
    const int e_group = 2 * e_mt + (e_m >> 3);
    const bool e_valid = e_mt < n_mt && e_group < gc;
    const int e_n = (g0 + e_group) * 8 + (e_m & 7);
    float wn_v = 0.f;
    int x_pos = -1;
    if (resid && e_valid) {
        // weight of the RMSNorm that consumes the new stream: pre_ca / pre_mlp of this layer, or the next
        // layer's pre_sa (the final norm after the last layer) - a cold line, so it is fetched up front
        const float* wn = gt == G_SO ? p.norms + ((size_t)layer * 3 + 1) * p.D
                        : gt == G_CO ? p.norms + ((size_t)layer * 3 + 2) * p.D
                                     : p.norms + ((size_t)(layer + 1) * 3) * p.D;
        wn_v = __ldg(wn + e_n);
        // ... and where this element sits in the contraction of the consumer (cross-q after self-o, mlp-in after cross-o,
        // the next layer's qkv or the logits head after mlp-out)
        const bool last = gt == G_WO && layer == p.L - 1;
        x_pos = row_pos(p, gt == G_SO ? G_CQ : gt == G_CO ? G_WI : last ? G_LOGITS : G_QKV, gt == G_WO ? layer + 1 : layer, e_n);
    }

    float acc[kMaxTiles][4];
#pragma unroll
    for (int i = 0; i < kMaxTiles; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

    // ---- lane constants ---------------------------------------------------------------------------------
    // A (ldmatrix.trans): matrix i = lane / 8 covers k rows 8*(i/2).. and columns 8*(i%2).. of the k-block
    const uint32_t lm_off = (uint32_t)(((lane & 7) + 8 * (lane >> 4)) * g.row_bytes + ((lane >> 3) & 1) * 16);
    const uint32_t ring_base = smem_u32(c.ring) + lm_off;
    const uint32_t kb_bytes = 16u * g.row_bytes;
    // 2:4: this lane's metadata word inside a 64-byte (tile, 32-row block) record - lanes with lane % 4 >= 2 load a
    // neighbour's word, which the instruction ignores (selector 0)
    const uint32_t meta_base = smem_u32(c.ring) + (uint32_t)g.meta_off + (uint32_t)(((lane >> 2) * 2 + (lane & 1)) * 4);
    // B: lane (n = lane / 4, q = lane % 4) holds rows {2q, 2q+1} and {2q+8, 2q+9} of MMA column n = 4 * row + term.
    // Staging: per k-block 64 words [row r][term t ^ s][q][2], s = 2 * (kb & 1) + r (bank swizzle)
    const uint32_t bst = smem_u32(c.xs) + warp * kBStageBytes;
    const int bq = lane & 3, bt = (lane >> 2) & 3, br = lane >> 4;
    const uint32_t b_off0 = (uint32_t)(br * 128 + (((bt ^ br) << 3) + bq * 2) * 4);
    const uint32_t b_off1 = (uint32_t)(256 + br * 128 + (((bt ^ (2 + br)) << 3) + bq * 2) * 4);
    const uint32_t bmask = bt == 3 ? 0u : 0xffffffffu;
    // store side: word pair p = lane + 32 i is (k-block (lane / 16) + 2 i, row (lane / 8) % 2, rows 2 kp, 2 kp + 1
    // with kp = lane % 8), i.e. register kp / 4 of the lanes with q = kp % 4
    const int ss = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
    const uint32_t s_base = bst + (uint32_t)((lane >> 4) * 256 + ((lane >> 3) & 1) * 128 +
                                             ((lane & 3) * 2 + ((lane >> 2) & 1)) * 4);
    const uint32_t s_t0 = s_base + ((0 ^ ss) << 5), s_t1 = s_base + ((1 ^ ss) << 5), s_t2 = s_base + ((2 ^ ss) << 5);
    const int tclass = p.tclass[gt];
    if (c.ts && c.tid == 0) c.ts[1] = clock64();

    // This warp owns contraction rows [warp * sl, (warp + 1) * sl), in fetch units of up to 256 rows (the staging
    // buffer).  Software pipeline: the input words of unit u + 1 are requested before the MMAs of unit u; the cpu ring
    // slots of a unit are stream positions warp + 8 * (u * cpu + jc).
    const int units = g.sl / g.urows, cpu = g.cpu, nkb_u = g.urows >> 4, nkb = rpc >> 4;
    const int row_w = warp * g.sl;
    ll_fetch(w, src, row_w, nkb_u, lane, f16);
#pragma unroll 1
    for (int u = 0; u < units; ++u) {
        const int row0 = row_w + u * g.urows;
        // ---- wait until the producing stage has written all input words of this unit
        unsigned spins = 0;
        while (__any_sync(0xffffffffu, ll_stale(w, f16))) {
            // 32 sentinels per warp: the first and the last word (row 0) of each 16-row k-block of the unit - a k-block
            // comes from at most two producer CTAs - polled with ONE warp-wide load per round trip; the full re-fetch
            // is issued when all of them carry the flag, so it is (almost always) the last one.
            {
                const bool live = (lane & 15) < nkb_u;
                const u64* sp = src + ((size_t)(row0 >> 4) + (lane & 15)) * 32 + ((lane >> 4) ? 15 : 0);
                bool ok = !live || (ll_ld(sp).y >> 16) == f16;
                while (!__all_sync(0xffffffffu, ok)) {
                    if (++spins > kMaxSpins) {
                        volatile int* e = reinterpret_cast<volatile int*>(p.err);
                        e[4] = 1 + row0 * 2; e[5] = (int)(ll_ld(sp).y >> 16); e[6] = (int)f16; e[7] = u;
                        ll_timeout(p.err, kErrFlagTimeout, c.seq * 16 + gt);
                    }
                    ll_check_abort(p.err, spins, 100 + kErrFlagTimeout, c.seq * 16 + gt);
                    if (!ok) ok = (ll_ld(sp).y >> 16) == f16;
                }
            }
            if (++spins > kMaxSpins) ll_timeout(p.err, kErrFlagTimeout + 1, c.seq * 16 + gt);
            ll_check_abort(p.err, spins, 100 + kErrFlagTimeout + 1, c.seq * 16 + gt);
            ll_fetch(w, src, row0, nkb_u, lane, f16);
        }
        if (c.ts && c.tid == 0 && u == 0) c.ts[6] = clock64();
        // ---- transpose into B fragments: word pair (k, k+1) of one row -> (hi|hi), (lo|lo), (lo2|lo2)
        const int pairs = nkb_u * 16;
#pragma unroll
        for (int i = 0; i < kLLW; ++i) {
            if (lane + 32 * i < pairs) {
                sts_u32(s_t0 + i * 512, __byte_perm(w[i].x, w[i].z, 0x5410));
                sts_u32(s_t1 + i * 512, __byte_perm(w[i].x, w[i].z, 0x7632));
                sts_u32(s_t2 + i * 512, __byte_perm(w[i].y, w[i].w, 0x5410));
            }
        }
        if (u + 1 < units) ll_fetch(w, src, row0 + g.urows, nkb_u, lane, f16);
        __syncwarp();
#pragma unroll 1
        for (int jc = 0; jc < cpu; ++jc) {
            const unsigned idx = c.cbase + warp + kConsumerWarps * (u * cpu + jc);
            const unsigned slot = idx % kNumSlots;
            ring_wait_full(c.misc, slot, (idx / kNumSlots) & 1u, p.err, (c.seq << 8) | slot);
            const uint32_t a_addr = ring_base + slot * kSlotBytes;
            const uint32_t b_base = bst + (uint32_t)(jc * nkb) * 256u;      // nkb is even: the parity swizzle lines up
            if (p.sparse24) {
                const uint32_t m_addr = meta_base + slot * kSlotBytes;
                const int n32 = nkb >> 1;
                if (tclass == 1) mma_chunk_sp<1>(acc, a_addr, kb_bytes, n32, b_base, b_off0, b_off1, bmask, m_addr, n_mt);
                else if (tclass == 2) mma_chunk_sp<2>(acc, a_addr, kb_bytes, n32, b_base, b_off0, b_off1, bmask, m_addr, n_mt);
                else if (tclass == 4) mma_chunk_sp<4>(acc, a_addr, kb_bytes, n32, b_base, b_off0, b_off1, bmask, m_addr, n_mt);
                else mma_chunk_sp<8>(acc, a_addr, kb_bytes, n32, b_base, b_off0, b_off1, bmask, m_addr, n_mt);
            }
            else if (tclass == 1) mma_chunk<1>(acc, a_addr, kb_bytes, nkb, b_base, b_off0, b_off1, bmask);
            else if (tclass == 2) mma_chunk<2>(acc, a_addr, kb_bytes, nkb, b_base, b_off0, b_off1, bmask);
            else if (tclass == 4) mma_chunk<4>(acc, a_addr, kb_bytes, nkb, b_base, b_off0, b_off1, bmask);
            else mma_chunk<8>(acc, a_addr, kb_bytes, nkb, b_base, b_off0, b_off1, bmask);
            __syncwarp();
            if (lane == 0) mbar_arrive(&c.misc->empty[slot]);
        }
        __syncwarp();                                        // the staging buffer is rewritten by the next unit
    }
    // fold the independent accumulator chains of the small tile classes (fixed order)
    if (tclass == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[0][j] = (acc[0][j] + acc[1][j]) + (acc[2][j] + acc[3][j]);
    } else if (tclass == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[0][j] += acc[2][j]; acc[1][j] += acc[3][j]; }
    }
    c.cbase += n_chunks;
    if (c.ts && c.tid == 0) c.ts[2] = clock64();

    if (c.ts && c.tid == 224) c.ts[8] = clock64();

    // ---- sum the three bf16 terms (MMA columns) of each batch row, then the 8 warps through smem ----
    // C fragment: c0,c1 = D[m][2q], D[m][2q+1]; c2,c3 = D[m+8][..] with m = lane/4, q = lane%4.
    // q = 0,1 hold row 0 (hi+lo | lo2+0), q = 2,3 hold row 1.
    {
        const int q = lane & 3, m = lane >> 2;
#pragma unroll
        for (int mt = 0; mt < kMaxTiles; ++mt) {
            if (mt < n_mt && mt < tclass) {               // (mt < tclass: uniform, skips the dead slots in one branch)
                float lo = acc[mt][0] + acc[mt][1], hi = acc[mt][2] + acc[mt][3];
                lo += __shfl_xor_sync(0xffffffffu, lo, 1);
                hi += __shfl_xor_sync(0xffffffffu, hi, 1);
                if ((q & 1) == 0) {
                    float* r = red + ((size_t)warp * n_mt + mt) * 32 + (q >> 1) * 16;
                    r[m] = lo;
                    r[m + 8] = hi;
                }
            }
        }
    }
    if (c.ts && c.tid == 224) c.ts[10] = clock64();
    consumer_sync();
    if (c.ts && c.tid == 0) c.ts[3] = clock64();
    if (e_mt >= n_mt) return;                          // whole warps without a tile are done

    float y = 0.f, y8 = 0.f;                           // column e_m and (mlp-in gate columns only) its up column e_m + 4
    {
        const float* rb = red + (size_t)e_mt * 32 + e_r * 16 + e_m;
#pragma unroll
        for (int ww = 0; ww < kConsumerWarps; ++ww) {
            y += rb[(size_t)ww * n_mt * 32];
            if (gt == G_WI && (e_m & 4) == 0) y8 += rb[(size_t)ww * n_mt * 32 + 4];     // the up column of this gate column
        }
    }
    const uint32_t f16n = c.seq & 0xffffu;
    if (c.ts && c.tid == 0) c.ts[5] = clock64();
    float inv = 1.0f;
    if (normed) {
        // 1/rms of the stage input: gathered from the producers' partial sums by the norm warp while the MMAs ran
        const volatile unsigned* fl = reinterpret_cast<const volatile unsigned*>(&c.misc->inv_seq[par]);
        unsigned spins = 0;
        while (*fl != c.seq) {
            if (++spins > (kMaxSpins << 3)) ll_timeout(p.err, kErrFlagTimeout + 4, c.seq);
        }
        __threadfence_block();
        inv = reinterpret_cast<const volatile float*>(c.misc->inv_rms[par])[e_r];
    }
    if (gt == G_WI) {
        // a group = gate columns 0..3 and up columns 4..7 of the same 4 hidden units: h = silu(gate) * up
        // (dia/layers.py:95-101)
        if (!e_valid || (e_m & 4) != 0) return;
        const float gate = y * inv, up = y8 * inv;
        const float h = (gate / (1.0f + expf(-gate))) * up;
        const int n = (g0 + e_group) * 4 + (e_m & 3);
        ll_store_parts(p.ll_hidden, n, e_r, h, f16n);
        return;
    }
    y *= inv;
    if (gt == G_QKV) {
        if (e_valid) ll_st(p.ll_qkv + (size_t)e_n * 2 + e_r, __float_as_uint(y), c.seq);
        return;
    }
    if (gt == G_CQ) {
        if (e_valid) ll_st(p.ll_cq + (size_t)e_n * 2 + e_r, __float_as_uint(y), c.seq);
        return;
    }
    if (gt == G_LOGITS) {
        // logits [2][C][V] (dia/layers.py:717-720) and, fused, the guided + masked logits the sampler draws from:
        // cond + s * (cond - uncond), then the -inf masks (dia/model.py:450-478).  Lanes 16..31 hold the cond row.
        const float un = __shfl_xor_sync(0xffffffffu, y, 16);
        const int ch = e_n / p.Vpad, vv = e_n - ch * p.Vpad;
        if (e_valid && ch < p.C && vv < p.V) {
            p.logits[((size_t)e_r * p.C + ch) * p.V + vv] = y;
            if (e_r == 1) {
                float gv = __fadd_rn(y, __fmul_rn(p.cfg_scale, __fsub_rn(y, un)));
                if ((ch > 0 && vv == p.eos) || vv == p.pad || vv == p.bos) gv = -INFINITY;
                ll_st(p.ll_glog + (size_t)ch * p.V + vv, __float_as_uint(gv), c.seq);
            }
        }
        return;
    }
    // residual add (dia/layers.py:555,574,582): the stream stays in this lane's register; the new stream is
    // published as the words of x * w_norm for the NEXT consumer, with this CTA's share of sum(x^2)
    float xn = 0.f;
    if (e_valid) {
        xn = c.xres + y;
        c.xres = xn;
        if (p.want_x) reinterpret_cast<float*>(p.x)[(size_t)e_n * 2 + e_r] = xn;
        if (x_pos >= 0) ll_store_parts(p.ll_x, x_pos, e_r, xn * wn_v, f16n);
    }
    if (c.ts && c.tid == 0) c.ts[7] = clock64();
    float sqv = xn * xn;                                               // half-warps = batch rows
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) sqv += __shfl_xor_sync(0xffffffffu, sqv, o);
    if ((lane & 15) == 0) ll_st(p.ll_ssq + 2 * blockIdx.x + e_r, __float_as_uint(sqv), c.seq);
}

// ------------------------------------------------------------------------------------------
// attention stages
// ------------------------------------------------------------------------------------------
// sum v[i] over the 32 lanes for NV values at once; afterwards lane l holds the total of
// value (l * NV / 32) [+ ...] - see callers for the index map
template <int NV>
__device__ __forceinline__ void transpose_reduce(float (&v)[NV], int lane) {
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

// Self-attention carries 4 query heads per KV tile (GQA 4:1); cross-attention has one query head per KV head.
// The shared-memory layout (queries, probabilities, per-warp partials) is the 4-head one for both.
constexpr int HPK = 4;

// NH = query heads that share a K/V tile: HPK for self-attention (GQA 4:1), 1 for cross-attention (a shared
// instantiation with three zero query heads cost cross-attention 4x the FMAs and shuffles: ~1 us per layer).
template <int NH>
__device__ void attn_stage(Ctx& c, int layer, int pos, int slot) {
    constexpr bool self = NH == HPK;
    const StepParams& p = *c.p;
    const int cta = blockIdx.x;
    const AttnWork w = self ? self_attn_work(p, cta, slot) : cross_attn_work(p, cta);
    if (!w.active) return;
    const int r = self ? w.pair / p.Hkv : 1;
    const int kvh = self ? w.pair - r * p.Hkv : w.pair;
    const int head0 = self ? kvh * HPK : w.pair;
    const int nh = self ? HPK : 1;                       // real query heads
    const int nsplit = self ? p.sa_nsplit : p.ca_nsplit;
    const uint32_t fprev = c.seq - 1;
    const bool has_new = self && w.has_new;

    float* qs = reinterpret_cast<float*>(c.xs);          // [HPK][128] rotated, pre-scaled queries
    float* kn = qs + HPK * kHeadDim;                     // [128] rotated key of this step
    float* vn = kn + kHeadDim;                           // [128] value of this step
    float* psm = vn + kHeadDim + c.warp * 64;            // [warps][16 keys][HPK] probabilities of the current tile
    float* wstat = vn + kHeadDim + kConsumerWarps * 64;  // [warps][8]: m[HPK], l[HPK] of each warp
    float* cw = wstat + kConsumerWarps * 8;              // split-combine staging
    float* racc = reinterpret_cast<float*>(c.xs) + 5120; // [warps][HPK][128]
    const u64* qsrc = self ? p.ll_qkv : p.ll_cq;

    // ---- inputs: wait for the projection that produced q (and k, v of this step), RoPE, scale ---------
    {
        const int pclamp = min(pos, p.n_pos - 1);
        const int d = c.tid & 63;                            // the same rotation pair in every iteration (256 % 64 == 0)
        const float sn = __ldg(p.rope_sin + (size_t)pclamp * 64 + d), cs = __ldg(p.rope_cos + (size_t)pclamp * 64 + d);
        const float scale = 0.08838834764831845f;           // 1/sqrt(128)
        // one lane per warp polls the first word the warp needs: a whole CTA spinning on its 768 input words would
        // cost the L2 too much; once that word is there the rest arrive within a round trip
        if ((c.warp >> 1) < nh) {
            if (c.lane == 0)
                ll_wait32(qsrc + ((size_t)(head0 + (c.warp >> 1)) * kHeadDim + (c.warp & 1) * 32) * 2 + r, fprev, p.err);
            __syncwarp();
        }
        const int n_items = (HPK + 2) * 64;                  // 4 query heads, then k, then v of this step
#pragma unroll 1
        for (int i = c.tid; i < n_items; i += kConsumerThreads) {
            const int hh = i >> 6;
            float o0 = 0.f, o1 = 0.f;
            if (hh < nh || (hh >= HPK && has_new)) {
                const int col = hh < HPK ? (head0 + hh) * kHeadDim
                                         : (hh == HPK ? (p.Hq + kvh) * kHeadDim : (p.Hq + p.Hkv + kvh) * kHeadDim);
                const u64* pa = qsrc + ((size_t)col + d) * 2 + r;
                uint2 wa = ll_ld(pa), wb = ll_ld(pa + 128);
                if (wa.y != fprev) wa.x = ll_wait32(pa, fprev, p.err);
                if (wb.y != fprev) wb.x = ll_wait32(pa + 128, fprev, p.err);
                const float a = __uint_as_float(wa.x), b = __uint_as_float(wb.x);
                o0 = a; o1 = b;
                if (hh <= HPK) {                             // RotaryEmbedding (dia/layers.py:161-173)
                    o0 = a * cs - b * sn;
                    o1 = a * sn + b * cs;
                    if (hh < HPK) { o0 *= scale; o1 *= scale; }
                }
            }
            qs[hh * kHeadDim + d] = o0;
            qs[hh * kHeadDim + d + 64] = o1;
        }
    }
    consumer_sync();
    if (c.ts && c.tid == 0) c.ts[1] = clock64();
    if (has_new && c.tid < kHeadDim) {
        // KVCache.update (dia/state.py:99-103): append this step's K/V at `slot`.  The row must be visible device-wide
        // before this CTA publishes its share of the stage OUTPUT (which is what lets, one step later, another CTA's
        // copy engine read it): the fence sits right before that publish, off the path of the other splits.
        const size_t row = ((size_t)w.pair * p.Lmax + slot) * kHeadDim;
        p.self_k[layer][row + c.tid] = kn[c.tid];
        p.self_v[layer][row + c.tid] = vn[c.tid];
    }
    float4 q[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) q[h] = reinterpret_cast<const float4*>(qs + h * kHeadDim)[c.lane];

    const int nk = w.k_hi - w.k_lo;
    const int nkc = (nk + 15) >> 4;
    const int nvc = nkc + (has_new ? 1 : 0);             // the key of this step is one more (virtual) tile

    // ---- one pass over the tiles of this warp (flash-style): scores of a K tile, running max / sum per head,
    //      then the V tile of the same keys.  Lane l owns score (key l/4 [+8], head l%4) and the running (m, l)
    //      of head l%4; every lane accumulates output dims 4*lane..4*lane+3 of all four heads.
    float m_run = -INFINITY, l_run = 0.f;
    float4 acc[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) acc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int ci = c.warp; ci < nvc; ci += kConsumerWarps) {
        const bool in_ring = ci < nkc;
        unsigned sl = 0;
        const float4* kt = reinterpret_cast<const float4*>(kn) + c.lane;
        const float4* vt = reinterpret_cast<const float4*>(vn) + c.lane;
        int keys_in = 1;
        if (in_ring) {
            const unsigned idx = c.cbase + 2 * ci;
            sl = idx % kNumSlots;
            ring_wait_full(c.misc, sl, (idx / kNumSlots) & 1u, p.err, (c.seq << 8) | 0x40 | sl);
            kt = reinterpret_cast<const float4*>(c.ring + sl * kSlotBytes) + c.lane;
            keys_in = min(16, nk - ci * 16);
        }
        if (c.ts && c.tid == 0 && ci == 0) c.ts[11] = clock64();
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
            transpose_reduce<8 * NH>(v, c.lane);             // lane l now holds (key l/4, head l%4 [head 0 if NH == 1])
            const float sv = (half * 8 + (c.lane >> 2)) < keys_in ? v[0] : -INFINITY;
            if (half == 0) s2[0] = sv; else s2[1] = sv;
        }
        if (in_ring) {
            __syncwarp();
            if (c.lane == 0) mbar_arrive(&c.misc->empty[sl]);
        }
        if (c.ts && c.tid == 0 && ci == 0) c.ts[12] = clock64();
        float mt = fmaxf(s2[0], s2[1]);                      // tile max / sum per head: over the 8 lanes of a head
        mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 4));
        mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 8));
        mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 16));
        const float m_new = fmaxf(m_run, mt);
        const float scale = expf(m_run - m_new);             // exp(-inf) = 0 on the first tile
        const float p0 = expf(s2[0] - m_new), p1 = expf(s2[1] - m_new);
        float lt = p0 + p1;
        lt += __shfl_xor_sync(0xffffffffu, lt, 4);
        lt += __shfl_xor_sync(0xffffffffu, lt, 8);
        lt += __shfl_xor_sync(0xffffffffu, lt, 16);
        l_run = l_run * scale + lt;
        m_run = m_new;
        psm[c.lane] = p0;                                    // index = key * 4 + head = lane (keys 0..7), +32 (8..15)
        psm[32 + c.lane] = p1;
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const float sh = __shfl_sync(0xffffffffu, scale, h);
            acc[h].x *= sh; acc[h].y *= sh; acc[h].z *= sh; acc[h].w *= sh;
        }
        __syncwarp();
        if (in_ring) {
            const unsigned idx = c.cbase + 2 * ci + 1;
            sl = idx % kNumSlots;
            ring_wait_full(c.misc, sl, (idx / kNumSlots) & 1u, p.err, (c.seq << 8) | 0x80 | sl);
            vt = reinterpret_cast<const float4*>(c.ring + sl * kSlotBytes) + c.lane;
        }
        if (c.ts && c.tid == 0 && ci == 0) c.ts[13] = clock64();
#pragma unroll 2
        for (int key = 0; key < keys_in; ++key) {
            const float4 vv = vt[key * 32];
            const float4 pr = *reinterpret_cast<const float4*>(psm + key * HPK);
            const float prh[4] = {pr.x, pr.y, pr.z, pr.w};
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                acc[h].x = fmaf(prh[h], vv.x, acc[h].x); acc[h].y = fmaf(prh[h], vv.y, acc[h].y);
                acc[h].z = fmaf(prh[h], vv.z, acc[h].z); acc[h].w = fmaf(prh[h], vv.w, acc[h].w);
            }
        }
        if (c.ts && c.tid == 0 && ci == 0) c.ts[14] = clock64();
        __syncwarp();                                        // psm is rewritten by the next tile
        if (in_ring && c.lane == 0) mbar_arrive(&c.misc->empty[sl]);
    }
    c.cbase += 2 * nkc;
    if (c.lane < HPK) { wstat[c.warp * 8 + c.lane] = m_run; wstat[c.warp * 8 + 4 + c.lane] = l_run; }
#pragma unroll
    for (int h = 0; h < NH; ++h)
        reinterpret_cast<float4*>(racc + ((size_t)c.warp * HPK + h) * kHeadDim)[c.lane] = acc[h];
    consumer_sync();
    if (c.ts && c.tid == 0) c.ts[2] = clock64();

    u64* oparts = self ? p.ll_attn : p.ll_cattn;
    u64* part = (self ? p.ll_sa_part : p.ll_ca_part) + ((size_t)w.pair * nsplit) * (nh * 132);
    const uint32_t f16n = c.seq & 0xffffu;
#pragma unroll 1
    for (int i = c.tid; i < nh * kHeadDim; i += kConsumerThreads) {
        const int h = i >> 7, d = i & 127;
        float M = -INFINITY;                              // merge the warps' (m, l, o) in warp order
#pragma unroll
        for (int ww = 0; ww < kConsumerWarps; ++ww) M = fmaxf(M, wstat[ww * 8 + h]);
        float o = 0.f, l = 0.f;
#pragma unroll
        for (int ww = 0; ww < kConsumerWarps; ++ww) {
            const float f = expf(wstat[ww * 8 + h] - M);
            l = fmaf(wstat[ww * 8 + 4 + h], f, l);
            o = fmaf(racc[((size_t)ww * HPK + h) * kHeadDim + d], f, o);
        }
        if (w.n_active == 1) {
            const float val = l > 0.f ? o / l : 0.f;
            const int k = row_pos(p, self ? G_SO : G_CO, layer, (head0 + h) * kHeadDim + d);
            if (k >= 0) {
                ll_store_parts(oparts, k, r, val, f16n);
                if (!self) ll_store_parts(oparts, k, 0, 0.f, f16n);   // the unconditional row attends nothing
            }
        } else {
            u64* pp = part + ((size_t)w.split * nh + h) * 132;
            ll_st(pp + 4 + d, __float_as_uint(o), c.seq);
            if (d == 0) { ll_st(pp, __float_as_uint(M), c.seq); ll_st(pp + 1, __float_as_uint(l), c.seq); }
        }
    }
    if (c.ts && c.tid == 0) c.ts[15] = clock64();
    consumer_sync();                                      // the scratch is reused by the next stage
    // The appended K/V row must be visible device-wide before another CTA's copy engine reads it one step later.  That
    // producer waits until its own math warps are past the cross-q stage of this step, which consumed the self-o
    // outputs of every CTA - stores this CTA issues after the fence below (see producer_loop).
    if (w.n_active == 1) {
        if (has_new && c.tid < kHeadDim) __threadfence();
        return;
    }

    // ---- every split combines a slice of the outputs from all splits' partials (fixed order) -------------
    const int E = nh * kHeadDim;
    const int per = (E + w.n_active - 1) / w.n_active;
    const int e0 = w.split * per, e1 = min(E, e0 + per);
    if (e0 >= e1) {                                       // (CTA-uniform)
        if (has_new && c.tid < kHeadDim) __threadfence();
        return;
    }
    const int ne = e1 - e0, na = w.n_active;
    const int h_lo = e0 >> 7, nhh = ((e1 - 1) >> 7) - h_lo + 1;
    float* cml = cw + ne * na;                            // cw: [ne][na] outputs, then [nhh][na][2] (m, l)
    const int n_items = ne * na + nhh * na * 2;           // <= 512 + na + 6 na  (na <= 18)
    auto item_src = [&](int i) -> const u64* {
        if (i < ne * na) {
            const int e = e0 + i / na, s = i - (i / na) * na;
            return part + ((size_t)s * nh + (e >> 7)) * 132 + 4 + (e & 127);
        }
        const int j = i - ne * na;
        const int hh = j / (na * 2), rem = j - hh * (na * 2);
        return part + ((size_t)(rem >> 1) * nh + h_lo + hh) * 132 + (rem & 1);
    };
    {   // every load of a thread is in flight before its first wait: one L2 round trip instead of three
        constexpr int kIt = 3;
        const u64* srcs[kIt];
        uint2 v[kIt];
#pragma unroll
        for (int it = 0; it < kIt; ++it) {
            const int i = c.tid + kConsumerThreads * it;
            if (i < n_items) { srcs[it] = item_src(i); v[it] = ll_ld(srcs[it]); }
        }
#pragma unroll
        for (int it = 0; it < kIt; ++it) {
            const int i = c.tid + kConsumerThreads * it;
            if (i < n_items) {
                unsigned spins = 0;
                while (v[it].y != c.seq) {
                    if (++spins > kMaxSpins) ll_timeout(p.err, kErrFlagTimeout + 2, c.seq);
                    ll_check_abort(p.err, spins, 100 + kErrFlagTimeout + 2, c.seq);
                    v[it] = ll_ld(srcs[it]);
                }
                cw[i] = __uint_as_float(v[it].x);
            }
        }
#pragma unroll 1
        for (int i = c.tid + kConsumerThreads * kIt; i < n_items; i += kConsumerThreads)
            cw[i] = __uint_as_float(ll_wait32(item_src(i), c.seq, p.err));
    }
    if (c.ts && c.tid == 0) c.ts[5] = clock64();
    consumer_sync();
    if (c.ts && c.tid == 0) c.ts[6] = clock64();
    // weights of the splits, exp(m_s - M), and sum(l_s * weight): warp hh takes head hh of the slice, lane = split
    float* cf = cml + nhh * na * 2;                        // [nhh][na] weights, then [nhh] normalisers
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
    {   // tpe threads per output element, each over a strided subset of the splits, then a butterfly (fixed order)
        int tpe = 8, sh = 3;
        while (tpe > 1 && ne * tpe > kConsumerThreads) { tpe >>= 1; --sh; }
        const int i = c.tid >> sh, j = c.tid & (tpe - 1);
        const bool live = i < ne;
        const int e = e0 + (live ? i : 0), h = e >> 7, d = e & 127;
        float O = 0.f;
        if (live) {
            const float* fw = cf + (size_t)(h - h_lo) * na;
#pragma unroll 1
            for (int s2 = j; s2 < na; s2 += tpe) O = fmaf(cw[i * na + s2], fw[s2], O);
        }
#pragma unroll
        for (int m = 4; m >= 1; m >>= 1)
            if (m < tpe) O += __shfl_xor_sync(0xffffffffu, O, m);
        if (live && j == 0) {
            const float Lsum = cf[nhh * na + (h - h_lo)];
            const float val = Lsum > 0.f ? O / Lsum : 0.f;
            const int k = row_pos(p, self ? G_SO : G_CO, layer, (head0 + h) * kHeadDim + d);
            if (k >= 0) {
                ll_store_parts(oparts, k, r, val, f16n);
                if (!self) ll_store_parts(oparts, k, 0, 0.f, f16n);
            }
        }
    }
    if (c.ts && c.tid == 0) c.ts[7] = clock64();
    if (has_new && c.tid < kHeadDim) __threadfence();
    consumer_sync();
}

// ------------------------------------------------------------------------------------------
// residual stream entry: embedding gather-sum (dia/layers.py:691-696: x = ((e0 + e1) + e2) ... + e8, both
// CFG rows), or - when a launch starts in the middle of a step - the stream a previous launch left in p.x.
// Warp 0 only: lane = row * 16 + column of this CTA's residual columns.
// ------------------------------------------------------------------------------------------
__device__ __noinline__ float enter_stream(const StepParams& p, unsigned char* xs, const SharedMisc* misc, int tid,
                                           bool embed, int pos, int step, const float* wnorm, uint32_t seq_out, int layer) {
    const GemmCfg& g = misc->gcfg[G_SO];
    if (g.gc == 0) return 0.f;                                // this CTA owns no residual columns
    const int lane = tid & 31, warp = tid >> 5, e_r = lane >> 4, e_m = lane & 15;
    const bool valid = (e_m >> 3) < g.gc;
    const int n = (g.g0 + (e_m >> 3)) * 8 + (e_m & 7);
    float x = 0.f;
    if (embed) {
        int* toks = reinterpret_cast<int*>(xs);             // [2][16] tokens, then [16][32] embedding rows
        float* part = reinterpret_cast<float*>(xs) + 32;
        if (warp == 0 && lane < p.C) {
            int t0, t1;
            if (p.tokens != nullptr && step == 0) { t0 = ldcg_i(p.tokens + lane); t1 = ldcg_i(p.tokens + p.C + lane); }
            else if (step == 0) { t0 = t1 = ldcg_i(p.grid + (size_t)(pos - 1) * p.C + lane); }
            else { t0 = t1 = (int)ll_wait32(p.ll_tok + lane, seq_out - 1, p.err); }
            if (t0 < 0 || t0 >= p.V || t1 < 0 || t1 >= p.V) { *p.err = kErrBadState; t0 = 0; t1 = 0; }
            toks[lane] = t0;
            toks[16 + lane] = t1;
        }
        consumer_sync();
        // warp w gathers channels w, w + 8: all rows of the tables are in flight at once
#pragma unroll 1
        for (int ch = warp; ch < p.C; ch += kConsumerWarps)
            part[ch * 32 + lane] = valid ? __ldg(p.emb + ((size_t)ch * p.V + toks[e_r * 16 + ch]) * p.D + n) : 0.f;
        consumer_sync();
        if (warp != 0) return 0.f;
        x = part[lane];
#pragma unroll 1
        for (int ch = 1; ch < p.C; ++ch) x += part[ch * 32 + lane];
    } else {
        if (warp != 0) return 0.f;
        if (valid) x = ldcg_f(reinterpret_cast<const float*>(p.x) + (size_t)n * 2 + e_r);
    }
    if (valid) {
        if (p.want_x) reinterpret_cast<float*>(p.x)[(size_t)n * 2 + e_r] = x;
        const int kp = row_pos(p, layer >= p.L ? G_LOGITS : G_QKV, layer, n);   // consumer: qkv of `layer` (or the logits head)
        if (kp >= 0) ll_store_parts(p.ll_x, kp, e_r, x * __ldg(wnorm + n), seq_out & 0xffffu);
    }
    float sqv = valid ? x * x : 0.f;
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) sqv += __shfl_xor_sync(0xffffffffu, sqv, o);
    if ((lane & 15) == 0) ll_st(p.ll_ssq + 2 * blockIdx.x + e_r, __float_as_uint(sqv), seq_out);
    return x;
}

// the sampling stage: CTA ch < C draws channel ch from the guided logits; CTA 0 then runs the body of the
// reference's while loop after _decoder_step (dia/model.py:771-807) and publishes the next step's tokens
__device__ void sample_stage(Ctx& c, int step_index, int pos) {
    const StepParams& p = *c.p;
    const int ch = blockIdx.x;
    if (ch >= p.C) return;
    SampleSmem* sm = reinterpret_cast<SampleSmem*>(c.xs);
    const u64* src = p.ll_glog + (size_t)ch * p.V;
    float g[kPerThread];
    {
        uint2 wv[kPerThread];
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            const int idx = c.tid + kConsumerThreads * i;
            wv[i] = idx < p.V ? ll_ld(src + idx) : make_uint2(0xff800000u, c.seq - 1);      // -inf past V
        }
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            if (wv[i].y != c.seq - 1) wv[i].x = ll_wait32(src + c.tid + kConsumerThreads * i, c.seq - 1, p.err);
            g[i] = __uint_as_float(wv[i].x);
        }
    }
    if (c.ts && c.tid == 0) c.ts[1] = clock64();
    const int tok = sample_channel_cta(g, p.V, p.temperature, p.top_p, p.top_k, p.seed, p.draw0 + step_index, ch,
                                       nullptr, sm, c.tid, c.ts);
    if (c.tid == 0) ll_st(p.ll_pred + ch, (uint32_t)tok, c.seq);
    if (c.ts && c.tid == 0) c.ts[3] = clock64();
    if (ch != 0 || c.warp != 0) return;

    // ---- CTA 0, warp 0: lane = channel ------------------------------------------------------------------
    const int lane = c.lane;
    int pr = 0;
    if (lane < p.C) {
        pr = (int)ll_wait32(p.ll_pred + lane, c.seq, p.err);
        p.pred_out[lane] = pr;
    }
    if (c.ts && c.tid == 0) c.ts[5] = clock64();
    GenState* gs = p.gs;
    int next_tok = 0;
    if (gs != nullptr && p.grid != nullptr) {
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
                if (cur != pos && lane == 0) *p.err = kErrBadState;
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
                    int* cell = p.grid + (size_t)cur * p.C + lane;
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
            }
        }
        steps_run += 1;
        if (lane == 0) {
            gs->dec_step = dec_step; gs->finished = finished; gs->eos_detected = eos_detected;
            gs->eos_countdown = eos_cd; gs->bos_countdown = bos_cd; gs->steps_run = steps_run;
        }
        if (finished) next_tok = 0;                                 // later steps of this launch are dead
    }
    if (lane < p.C) ll_st(p.ll_tok + lane, (uint32_t)next_tok, c.seq);
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(kThreads, 1) dia_step_kernel(const __grid_constant__ StepParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* ring = smem;
    unsigned char* xs = smem + kNumSlots * kSlotBytes;
    float* red = reinterpret_cast<float*>(smem + kNumSlots * kSlotBytes + kXsBytes);
    SharedMisc* misc = reinterpret_cast<SharedMisc*>(smem + kNumSlots * kSlotBytes + kXsBytes + 2 * kRedBytes);

    const int tid = threadIdx.x;
    // a launch queued behind the one that finished the utterance is a no-op.  Uniform: `finished` is only
    // written at the end of a step, which needs every CTA to have taken part in that step's stages.
    if (p.gs != nullptr && p.grid != nullptr && ldcg_i(&p.gs->finished) != 0) return;
    if (tid == 0) {
#pragma unroll 1
        for (int i = 0; i < kNumSlots; ++i) { mbar_init(&misc->full[i], 1); mbar_init(&misc->empty[i], 1); }
        misc->stages_done = 0;
        misc->inv_seq[0] = 0; misc->inv_seq[1] = 0;
        fence_mbar_init();
    }
    {   // copy this CTA's table
        const int* src = reinterpret_cast<const int*>(p.cta_tab + blockIdx.x);
        int* dst = reinterpret_cast<int*>(&misc->tab);
#pragma unroll 1
        for (int i = tid; i < (int)(sizeof(CtaTable) / 4); i += kThreads) dst[i] = src[i];
    }
    __syncthreads();
    if (tid < G_COUNT) {
        GemmCfg& g = misc->gcfg[tid];
        g.gc = misc->tab.gc[tid]; g.g0 = misc->tab.g0[tid]; g.K = p.Kdim[tid];
        g.row_bytes = g.gc * 16;
        g.rpc = g.gc > 0 ? gemm_slot_rows(g.gc, g.K, p.sparse24) : 16;
        g.slot_bytes = g.gc > 0 ? (int)gemm_slot_bytes(g.gc, g.K, p.sparse24) : 0;
        g.meta_off = (g.rpc / 2) * g.gc * 16;
        g.n_chunks = g.gc > 0 ? g.K / g.rpc : 0;
        g.sl = g.K / kConsumerWarps;
        g.urows = min(g.sl, kMaxKb * 16);
        g.cpu = g.urows / g.rpc;
        g.n_mt = (g.gc + 1) >> 1;
    }
    __syncthreads();

    if (tid >= kConsumerThreads) {
        if (tid == kProducerWarp * 32) producer_loop(p, ring, misc);
        else if ((tid >> 5) == kNormWarp) norm_warp_loop(p, misc);
        return;
    }

    Ctx c;
    c.p = &p; c.ring = ring; c.xs = xs; c.red = red; c.misc = misc;
    c.tid = tid; c.warp = tid >> 5; c.lane = tid & 31;
    c.cbase = 0; c.seq = 0; c.xres = 0.f; c.ts = nullptr;
    const int S = 8 * p.L + 3;

    if (p.stage_begin > 0) {
        // the stream a previous launch left in p.x, published as if stage_begin - 1 had just produced it
        const int layer = (p.stage_begin - 1) >> 3;
        c.seq = (unsigned)p.stage_begin;
        c.xres = enter_stream(p, xs, misc, tid, false, 0, 0, p.norms + (size_t)layer * 3 * p.D, c.seq, layer);
    }
#pragma unroll 1
    for (int n = 0; n < p.n_steps; ++n) {
        const int pos = p.pos0 + n, slot = p.slot0 + n;
#pragma unroll 1
        for (int s = p.stage_begin; s < p.stage_end; ++s) {
            int kind, layer;
            decode_stage(s, p.L, kind, layer);
            c.seq = 1u + (unsigned)(n * S + s);
            c.ts = (p.timing != nullptr && (int)blockIdx.x == p.timing_cta && (tid == 0 || tid == 224)) ? p.timing + ((size_t)n * S + s) * 16 : nullptr;
            if (c.ts && tid == 0) c.ts[0] = clock64();
            switch (kind) {
                case S_EMBED: c.xres = enter_stream(p, xs, misc, tid, true, pos, n, p.norms, c.seq, 0); break;
                case S_SATTN: attn_stage<HPK>(c, layer, pos, slot); break;
                case S_CATTN: attn_stage<1>(c, layer, pos, slot); break;
                case S_SAMPLE: sample_stage(c, n, pos); break;
                default: gemm_stage(c, gemm_of_kind(kind), layer); break;
            }
            if (c.ts && tid == 0) c.ts[4] = clock64();
            if (p.cta_timing != nullptr && n == 1 && tid == 0) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
                p.cta_timing[(size_t)s * gridDim.x + blockIdx.x] = t;
            }
            if (tid == 0) st_release_cta_s32(&misc->stages_done, n * S + s + 1);
        }
    }
}

// standalone head: CFG + masks + sampling on caller-provided logits (dia/model.py:447-488); one CTA per channel
extern "C" __global__ void __launch_bounds__(kConsumerThreads, 1)
dia_head_sample_kernel(const __grid_constant__ StepParams p, const float* logits, unsigned long long draw, int* pred,
                       float* probs) {
    __shared__ SampleSmem sm;
    const int ch = blockIdx.x, tid = threadIdx.x;
    float g[kPerThread];
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
        const int v = tid + kConsumerThreads * i;
        float gv = -INFINITY;
        if (v < p.V) {
            const float un = logits[(size_t)ch * p.V + v], co = logits[((size_t)p.C + ch) * p.V + v];
            gv = __fadd_rn(co, __fmul_rn(p.cfg_scale, __fsub_rn(co, un)));
            if ((ch > 0 && v == p.eos) || v == p.pad || v == p.bos) gv = -INFINITY;
        }
        g[i] = gv;
    }
    const int tok = sample_channel_cta(g, p.V, p.temperature, p.top_p, p.top_k, p.seed, draw, ch,
                                       probs ? probs + (size_t)ch * p.V : nullptr, &sm, tid);
    if (tid == 0) pred[ch] = tok;
}

int step_kernel_smem_bytes() { return kSmemBytes; }

cudaError_t launch_step_kernel(const StepParams& p, bool cooperative, cudaStream_t st) {
    static bool attr_set[64] = {};                     // function attributes are per device
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        e = cudaFuncSetAttribute(dia_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    (void)cooperative;     // CTAs wait on each other's data: co-residency is always required
    void* args[] = {const_cast<StepParams*>(&p)};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(dia_step_kernel), dim3(p.G), dim3(kThreads), args,
                                       kSmemBytes, st);
}

cudaError_t launch_head_sample(const StepParams& p, const float* logits, unsigned long long draw, int* pred,
                               float* probs, cudaStream_t st) {
    dia_head_sample_kernel<<<p.C, kConsumerThreads, 0, st>>>(p, logits, draw, pred, probs);
    return cudaGetLastError();
}

// carve the LL region: fills the ll_* pointers of `out` from `base` (may be null to only size it); returns bytes
size_t ll_layout(const StepParams& g, StepParams* out, unsigned long long* base) {
    size_t off = 0;
    auto take = [&](unsigned long long*& ptr, size_t words) {
        if (out) ptr = base ? base + off : nullptr;
        off += (words + 15) & ~(size_t)15;
    };
    StepParams scratch;
    StepParams& o = out ? *out : scratch;
    const size_t nq = (size_t)g.Hq * kHeadDim, nkv = (size_t)g.Hkv * kHeadDim, nc = (size_t)g.Hc * kHeadDim;
    take(o.ll_x, (size_t)g.D * 2);
    take(o.ll_attn, nq * 2);
    take(o.ll_cattn, nc * 2);
    take(o.ll_hidden, (size_t)g.F * 2);
    take(o.ll_qkv, (nq + 2 * nkv) * 2);
    take(o.ll_cq, nc * 2);
    take(o.ll_ssq, (size_t)g.G * 2);
    take(o.ll_sa_part, (size_t)2 * g.Hkv * g.sa_nsplit * 4 * 132);
    take(o.ll_ca_part, (size_t)g.Hc * g.ca_nsplit * 132);
    take(o.ll_glog, (size_t)g.C * g.V);
    take(o.ll_pred, DIA_B200_MAX_CHANNELS);
    take(o.ll_tok, DIA_B200_MAX_CHANNELS);
    return off * sizeof(unsigned long long);
}

}  // namespace dia
