// step_kernel.cu - the persistent Dia decode-step kernel for sm_100a.
//
// One CTA per SM (148 on B200), launched cooperatively.  A decode step is a chain of
// ~147 strictly dependent stages (embed, 8 per decoder layer, logits, sample) over a
// batch of 2 rows (CFG uncond/cond) - 2 FLOP per weight byte, i.e. bound by streaming
// 2.53 GB of bf16 weights from HBM3e, not by math.  The design follows from that:
//
//  * every CTA owns a fixed slice of the output columns of every GEMM, and the host
//    repacks the weights ONCE into one contiguous bf16 byte stream per CTA in exactly
//    the order that CTA consumes them;
//  * a producer warp (one elected lane) walks that stream with 1-D bulk async copies
//    (cp.async.bulk, the TMA engine) into a 19 x 8 KB shared-memory ring guarded by
//    full/empty mbarriers.  It never joins the grid barriers, so it keeps prefetching
//    the NEXT stages' weights while the math warps sit in a barrier or in a
//    latency-bound attention stage - HBM stays busy across all 147 stage boundaries;
//  * 8 math warps consume ring slots (one slot per warp at a time), convert bf16 pairs
//    with two ALU ops and accumulate in fp32 with packed FFMA2; activations, residual
//    stream, norms, softmax and the KV cache stay fp32 (required for
//    the 2e-2 / bit-exact-greedy bar), reductions run in a fixed order (deterministic);
//  * stages are separated by a counter-based grid barrier (one atomic + an acquire poll);
//  * self-attention streams its K/V tiles through the same ring (split-KV over CTAs,
//    4 query heads share each KV tile, RoPE fused, K/V append fused, last-arriver
//    combine), cross-attention reads only the valid keys of the conditional row;
//  * the sampling stage (CFG combine, masks, argmax / top-k / top-p / Philox draw) and the
//    EOS state machine run on the device, so a whole run of steps needs no host sync.
//
// Reference semantics: dia/layers.py:671-720 (decode_step), :530-584 (DecoderLayer),
// :238-346 (Attention), :92-105 (MlpBlock); dia/model.py:429-488, 32-82, 748-807.
#include <cuda_bf16.h>

#include "common.cuh"
#include "engine_internal.h"

namespace dia {

// per-GEMM constants of this CTA, computed once per launch (no integer divisions in the stage loop)
struct GemmCfg {
    int gc, g0, K;
    int row_bytes;     // gc * 16
    int rpc;           // rows of the slab per ring slot (multiple of 16 = one MMA k-block)
    int n_chunks;
    int n_mt;          // 16-column MMA tiles = ceil(gc / 2)
};

struct SharedMisc {
    uint64_t full[kNumSlots];
    uint64_t empty[kNumSlots];
    CtaTable tab;
    GemmCfg gcfg[G_COUNT];
    float ssq_part[kConsumerWarps][2];   // per-warp sums of x_new^2 of a residual stage (published at the barrier)
    float inv_rms[2];                    // 1/rms of the stage input, per batch row
    float stat[32];
    int stages_done;     // consumer -> producer progress (global stage index + 1)
    int flag;
};
static_assert(sizeof(SharedMisc) <= kMiscBytes, "misc region too small");

struct Ctx {
    const StepParams* p;
    unsigned char* ring;
    float2* xs;
    float* red;
    SharedMisc* misc;
    int tid, warp, lane;
    unsigned cbase;      // ring chunk index at the start of the current stage
    unsigned nbar;       // grid barriers passed so far in this launch
    long long* tstamp;   // debug: 8 clock64 slots of the current stage (CTA 0, thread 0 only)
};

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
// slots [0, slot); the key/value of THIS step is taken from the qkv scratch by split 0.
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
    w.has_new = (w.split == 0);
    return w;
}
// cross-attention: conditional row only, one query head per KV head, keys [0, text_len)
__device__ __forceinline__ AttnWork cross_attn_work(const StepParams& p, int cta) {
    AttnWork w;
    const int nsplit = p.ca_nsplit, n = p.text_len;
    w.pair = cta / nsplit;
    w.split = cta - w.pair * nsplit;
    int per = (n + nsplit - 1) / nsplit;
    per = (per + 15) & ~15;
    if (per < 32) per = 32;
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
// grid barrier (consumer warps only)
// ------------------------------------------------------------------------------------------
// thread 0, after a consumer_sync: this CTA's share of sum(x^2) of the new residual stream (fixed order)
__device__ __forceinline__ void publish_ssq_partials(Ctx& c) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int w = 0; w < kConsumerWarps; ++w) { s0 += c.misc->ssq_part[w][0]; s1 += c.misc->ssq_part[w][1]; }
    c.p->ssq[blockIdx.x] = s0;
    c.p->ssq[gridDim.x + blockIdx.x] = s1;
}

__device__ __forceinline__ void grid_barrier(Ctx& c, bool publish_ssq) {
    const StepParams& p = *c.p;
    consumer_sync();                       // every consumer thread's stage output is issued
    c.nbar++;
    if (c.tid == 0) {
        if (publish_ssq) publish_ssq_partials(c);
        // release: orders the whole CTA's prior writes (bar.sync above makes them visible to this
        // thread, the release is cumulative) before the arrival becomes visible at gpu scope
        red_release_add_u32(p.grid_bar, 1u);
        const unsigned target = c.nbar * gridDim.x;
        if (ld_relaxed_u32(p.grid_bar) < target) {
            const unsigned long long t0 = clock64();
            while (ld_relaxed_u32(p.grid_bar) < target) {
                if (clock64() - t0 > kWatchdogCycles) {
                    *reinterpret_cast<volatile int*>(p.err) = kErrGridBarrierTimeout;
                    __threadfence_system();
                    __trap();
                }
            }
        }
        // acquire side: every read of data another CTA produced goes to L2 (ld.global.cg / bulk copies),
        // never through L1, and is issued after this loop exits (GPUs do not speculate past the branch),
        // so no L1 invalidation / fence is needed here.  The named barrier below orders the other threads.
    }
    consumer_sync();
}

// ------------------------------------------------------------------------------------------
// producer: walks this CTA's byte stream
// ------------------------------------------------------------------------------------------
struct Producer {
    unsigned char* ring;
    SharedMisc* misc;
    int* err;
    unsigned pc;
    uint64_t pol_stream, pol_keep;
    __device__ __forceinline__ void issue(const void* src, uint32_t bytes, bool keep) {
        const unsigned slot = pc % kNumSlots;
        const unsigned ph = (pc / kNumSlots) & 1u;
        mbar_wait(&misc->empty[slot], ph ^ 1u, err, kErrEmptyBarrierTimeout);
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
    for (int n = 0; n < p.n_steps; ++n) {
        const int slot = p.slot0 + n;
        for (int s = p.stage_begin; s < p.stage_end; ++s) {
            int kind, layer;
            decode_stage(s, p.L, kind, layer);
            const int gt = gemm_of_kind(kind);
            if (gt >= 0) {
                const GemmCfg& g = misc->gcfg[gt];
                if (g.gc == 0) continue;
                const unsigned char* base = p.wstream + tab.stream_base +
                    (gt == G_LOGITS ? tab.logits_off
                                    : (unsigned long long)layer * tab.layer_bytes + tab.slab_off[gt]);
                for (int r0 = 0; r0 < g.K; r0 += g.rpc) {
                    const int rows = min(g.rpc, g.K - r0);
                    pr.issue(base + (size_t)r0 * g.row_bytes, rows * g.row_bytes, false);
                    if (r0 == 0 && p.timing != nullptr && cta == 0) p.timing[((size_t)n * S + s) * 8 + 7] = clock64();
                }
            } else if (kind == S_SATTN || kind == S_CATTN) {
                const bool self = kind == S_SATTN;
                const AttnWork w = self ? self_attn_work(p, cta, slot) : cross_attn_work(p, cta);
                if (!w.active || w.k_hi <= w.k_lo) continue;
                if (self && n > 0) {
                    // rows < slot were written by other CTAs during the same stage of step n-1:
                    // do not run ahead of that (never blocks at Dia-1.6B sizes)
                    const int need = (n - 1) * S + s + 1;
                    if (ld_acquire_cta_s32(&misc->stages_done) < need) {
                        const unsigned long long t0 = clock64();
                        while (ld_acquire_cta_s32(&misc->stages_done) < need) {
                            if (clock64() - t0 > kWatchdogCycles) {
                                *reinterpret_cast<volatile int*>(p.err) = kErrStepDoneTimeout;
                                __threadfence_system();
                                __trap();
                            }
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
                for (int pass = 0; pass < 2; ++pass) {
                    const float* b = pass == 0 ? kb : vb;
                    for (int k0 = w.k_lo; k0 < w.k_hi; k0 += 16) {
                        const int nk = min(16, w.k_hi - k0);
                        pr.issue(b + (size_t)k0 * kHeadDim, nk * kHeadDim * 4, !self);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// GEMM stage: y[2][N_cta] = x[2][K] . W_slab on the tensor cores (mma.sync m16n8k16, bf16 x bf16 ->
// fp32).  Swap-AB: the 16 rows of the MMA are 16 OUTPUT COLUMNS (two 8-column groups of the slab,
// fed from the ring with ldmatrix.trans - the slab is [k][n], n contiguous), the 8 columns of the MMA
// carry the two batch rows, each split into three bf16 terms  x = hi + lo + lo2  (exact to 24 bits),
// so the product keeps fp32-activation accuracy while the weights stay bf16.
//   MMA column n: 0,1,2 = row 0 (hi, lo, lo2)   4,5,6 = row 1 (hi, lo, lo2)   3,7 = zero
// Activation vectors live in global memory ALREADY split and laid out as B fragments ("parts"):
// k-block kb (16 rows) = 256 B, the 8-byte word of lane l = n*4 + kq holds the bf16 of rows
// {2kq, 2kq+1, 8+2kq, 9+2kq} of column n.  The epilogue that produces a vector writes it in this form
// once (pre-multiplied by the consumer's RMSNorm weight), so a consumer warp fetches exactly the
// fragments of the k-range it owns straight from L2 - no staging pass, no block sync before the MMAs.
// The 1/rms factor commutes with the GEMM and is applied in the epilogue from per-CTA partial sums.
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
// x = hi + lo + lo2 with each term bf16 (round-to-nearest at every level; the residuals are exact)
__device__ __forceinline__ void split3(float x, unsigned short (&t)[3]) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h);
    const __nv_bfloat16 l = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(l);
    const __nv_bfloat16 l2 = __float2bfloat16_rn(r2);
    t[0] = __bfloat16_as_ushort(h); t[1] = __bfloat16_as_ushort(l); t[2] = __bfloat16_as_ushort(l2);
}
// write element (k, batch row r) of a vector into its parts buffer
__device__ __forceinline__ void store_parts(unsigned short* parts, int k, int r, float v) {
    unsigned short t[3];
    split3(v, t);
    const int kk = k & 15;
    unsigned short* base = parts + (size_t)(k >> 4) * 128 + ((kk & 7) >> 1) * 4 + (kk & 1) + ((kk >> 3) << 1);
#pragma unroll
    for (int i = 0; i < 3; ++i) base[(r * 4 + i) * 16] = t[i];
}

constexpr int kMaxTiles = 8;     // <= 16 column groups per CTA and GEMM
constexpr int kMaxKb = 16;       // k-blocks per ring slot (slot rows are capped at 256)
constexpr int kBStageBytes = kMaxKb * 256;   // B fragments of one slot; two such buffers per warp

__device__ __forceinline__ void cp_async16_cg(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// L2 -> this warp's staging buffer: the B fragments of rows [r0, r0 + 16*nkb) (256 B per k-block;
// half-warps take alternate k-blocks, 16 B per lane, L2-only so no stale L1 line can be read)
__device__ __forceinline__ void stage_b_frags(uint32_t dst, const unsigned char* parts, int r0, int nkb, int lane) {
    const unsigned char* src = parts + (size_t)(r0 >> 4) * 256 + (lane & 15) * 16;
    const uint32_t d = dst + (lane & 15) * 16;
    for (int kb = lane >> 4; kb < nkb; kb += 2) cp_async16_cg(d + kb * 256, src + (size_t)kb * 256);
    cp_async_commit();
}
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}

// The MMAs of one ring slot.  Three compact, rolled variants (the kernel has ~150 short stages per
// step; straight-line code that overflows the instruction caches costs more than it saves):
//   1 tile : 4 k-blocks per iteration on 4 independent accumulator chains
//   2 tiles: 2 k-blocks per iteration, 2 chains per tile
//   n tiles: 1 k-block per iteration, the tiles are the independent chains
__device__ __forceinline__ void mma_slot_1(float (&acc)[kMaxTiles][4], uint32_t a_addr, uint32_t kb_bytes, int nkb,
                                           uint32_t b_addr) {
    int kb = 0;
    for (; kb + 4 <= nkb; kb += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint2 b = lds_u2(b_addr + (kb + u) * 256);
            uint32_t a[4];
            ldmatrix_x4_trans(a, a_addr + (kb + u) * kb_bytes);
            mma_bf16_16816(acc[u], a, b.x, b.y);
        }
    }
    for (; kb < nkb; ++kb) {
        const uint2 b = lds_u2(b_addr + kb * 256);
        uint32_t a[4];
        ldmatrix_x4_trans(a, a_addr + kb * kb_bytes);
        mma_bf16_16816(acc[0], a, b.x, b.y);
    }
}
__device__ __forceinline__ void mma_slot_2(float (&acc)[kMaxTiles][4], uint32_t a_addr, uint32_t kb_bytes, int nkb,
                                           uint32_t b_addr) {
    int kb = 0;
    for (; kb + 2 <= nkb; kb += 2) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint2 b = lds_u2(b_addr + (kb + u) * 256);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                uint32_t a[4];
                ldmatrix_x4_trans(a, a_addr + (kb + u) * kb_bytes + mt * 32);
                mma_bf16_16816(acc[u * 2 + mt], a, b.x, b.y);
            }
        }
    }
    if (kb < nkb) {
        const uint2 b = lds_u2(b_addr + kb * 256);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            uint32_t a[4];
            ldmatrix_x4_trans(a, a_addr + kb * kb_bytes + mt * 32);
            mma_bf16_16816(acc[mt], a, b.x, b.y);
        }
    }
}
__device__ __forceinline__ void mma_slot_n(float (&acc)[kMaxTiles][4], uint32_t a_addr, uint32_t kb_bytes, int nkb,
                                           uint32_t b_addr, int n_mt) {
    for (int kb = 0; kb < nkb; ++kb) {
        const uint2 b = lds_u2(b_addr + kb * 256);
#pragma unroll
        for (int mt = 0; mt < kMaxTiles; ++mt) {
            if (mt < n_mt) {
                uint32_t a[4];
                ldmatrix_x4_trans(a, a_addr + kb * kb_bytes + mt * 32);
                mma_bf16_16816(acc[mt], a, b.x, b.y);
            }
        }
    }
}

__device__ void gemm_stage(Ctx& c, int gt, int layer) {
    const StepParams& p = *c.p;
    const GemmCfg& g = c.misc->gcfg[gt];
    const int gc = g.gc, g0 = g.g0, K = g.K, n_mt = g.n_mt;
    const int lane = c.lane;
    const bool resid = (gt == G_SO || gt == G_CO || gt == G_WO);
    if (resid && lane == 0) { c.misc->ssq_part[c.warp][0] = 0.f; c.misc->ssq_part[c.warp][1] = 0.f; }
    if (gc == 0) return;

    const unsigned char* parts;
    bool normed = false;
    switch (gt) {
        case G_SO: parts = reinterpret_cast<const unsigned char*>(p.attn_parts); break;
        case G_CO: parts = reinterpret_cast<const unsigned char*>(p.cattn_parts); break;
        case G_WO: parts = reinterpret_cast<const unsigned char*>(p.hidden_parts); break;
        default: parts = reinterpret_cast<const unsigned char*>(p.xparts); normed = true; break;
    }
    const int n_chunks = g.n_chunks, rpc = g.rpc;
    // this warp's slots are ci = warp, warp + 8, ...; the B fragments of its first slot start moving now
    const uint32_t bstage = smem_u32(c.xs) + c.warp * (2 * kBStageBytes);
    if (c.warp < n_chunks) stage_b_frags(bstage, parts, c.warp * rpc, min(rpc, K - c.warp * rpc) >> 4, lane);

    // epilogue role of this thread: (tile, row, column-in-tile); its residual is fetched now, used last
    const int e_mt = c.tid >> 5, e_r = (c.tid >> 4) & 1, e_m = c.tid & 15;
    const int e_group = 2 * e_mt + (e_m >> 3);
    const bool e_valid = e_mt < n_mt && e_group < gc;
    const int e_n = (g0 + e_group) * 8 + (e_m & 7);
    float resid_v = 0.f;
    if (resid && e_valid) resid_v = ldcg_f(reinterpret_cast<const float*>(p.x) + (size_t)e_n * 2 + e_r);

    if (normed && c.warp == kConsumerWarps - 1) {
        // 1/rms of the input from the per-CTA partial sums its producers published (fixed order)
        float s0 = 0.f, s1 = 0.f;
        for (int i = lane; i < (int)gridDim.x; i += 32) { s0 += ldcg_f(p.ssq + i); s1 += ldcg_f(p.ssq + gridDim.x + i); }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        if (lane == 0) {   // torch.nn.RMSNorm: x * rsqrt(mean(x^2) + eps) * w
            c.misc->inv_rms[0] = 1.0f / sqrtf(s0 / (float)K + p.eps);
            c.misc->inv_rms[1] = 1.0f / sqrtf(s1 / (float)K + p.eps);
        }
    }

    float acc[kMaxTiles][4];
#pragma unroll
    for (int i = 0; i < kMaxTiles; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

    // ldmatrix row address of this lane inside a k-block: tile t = lane/8 -> (k half = t/2, group = t%2)
    const uint32_t lm_off = (uint32_t)((((lane >> 4) << 3) + (lane & 7)) * g.row_bytes + ((lane >> 3) & 1) * 16);
    const uint32_t ring_base = smem_u32(c.ring) + lm_off;
    const uint32_t kb_bytes = 16u * g.row_bytes;
    if (c.tstamp) c.tstamp[1] = clock64();

    int buf = 0;
    for (int ci = c.warp; ci < n_chunks; ci += kConsumerWarps, buf ^= 1) {
        const int cn = ci + kConsumerWarps;
        if (cn < n_chunks) stage_b_frags(bstage + (buf ^ 1) * kBStageBytes, parts, cn * rpc, min(rpc, K - cn * rpc) >> 4, lane);
        const unsigned idx = c.cbase + ci;
        const unsigned slot = idx % kNumSlots;
        mbar_wait(&c.misc->full[slot], (idx / kNumSlots) & 1u, p.err, kErrFullBarrierTimeout);
        if (cn < n_chunks) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();
        if (c.tstamp && ci == c.warp) c.tstamp[6] = clock64();
        const int nkb = min(rpc, K - ci * rpc) >> 4;
        const uint32_t a_addr = ring_base + slot * kSlotBytes;
        const uint32_t b_addr = bstage + buf * kBStageBytes + lane * 8;
        if (n_mt == 1) mma_slot_1(acc, a_addr, kb_bytes, nkb, b_addr);
        else if (n_mt == 2) mma_slot_2(acc, a_addr, kb_bytes, nkb, b_addr);
        else mma_slot_n(acc, a_addr, kb_bytes, nkb, b_addr, n_mt);
        __syncwarp();
        if (lane == 0) mbar_arrive(&c.misc->empty[slot]);
    }
    c.cbase += n_chunks;
    if (n_mt == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[0][j] = (acc[0][j] + acc[1][j]) + (acc[2][j] + acc[3][j]);
    } else if (n_mt == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[0][j] += acc[2][j]; acc[1][j] += acc[3][j]; }
    }
    if (c.tstamp) c.tstamp[2] = clock64();

    // ---- sum the three bf16 terms (MMA columns) of each batch row, then the 8 warps through smem ----
    // C fragment: c0,c1 = D[m][2q], D[m][2q+1]; c2,c3 = D[m+8][..] with m = lane/4, q = lane%4.
    // q = 0,1 hold row 0 (hi+lo | lo2+0), q = 2,3 hold row 1.
    const int q = lane & 3, m = lane >> 2;
#pragma unroll
    for (int mt = 0; mt < kMaxTiles; ++mt) {
        if (mt < n_mt) {
            float lo = acc[mt][0] + acc[mt][1], hi = acc[mt][2] + acc[mt][3];
            lo += __shfl_xor_sync(0xffffffffu, lo, 1);
            hi += __shfl_xor_sync(0xffffffffu, hi, 1);
            if ((q & 1) == 0) {
                float* r = c.red + ((size_t)c.warp * n_mt + mt) * 32 + (q >> 1) * 16;
                r[m] = lo;
                r[m + 8] = hi;
            }
        }
    }
    consumer_sync();
    if (c.tstamp) c.tstamp[3] = clock64();

    const float inv = normed ? c.misc->inv_rms[e_r] : 1.0f;
    auto total = [&](int mm) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; ++w) s += c.red[((size_t)w * n_mt + e_mt) * 32 + e_r * 16 + mm];
        return s;
    };
    if (gt == G_WI) {
        // a tile = (gate group, up group) of the same 8 hidden units: h = silu(gate) * up (dia/layers.py:95-101)
        if (!e_valid || e_m >= 8) return;
        const float gate = total(e_m) * inv, up = total(e_m + 8) * inv;
        const float h = (gate / (1.0f + expf(-gate))) * up;
        const int n = ((g0 >> 1) + e_mt) * 8 + e_m;
        reinterpret_cast<float*>(p.hidden)[(size_t)n * 2 + e_r] = h;
        store_parts(p.hidden_parts, n, e_r, h);
        return;
    }
    if (!resid) {
        if (!e_valid) return;
        const float y = total(e_m) * inv;
        if (gt == G_QKV) {
            reinterpret_cast<float*>(p.qkv)[(size_t)e_n * 2 + e_r] = y;
        } else if (gt == G_CQ) {
            reinterpret_cast<float*>(p.cq)[(size_t)e_n * 2 + e_r] = y;
        } else {
            const int ch = e_n / p.Vpad, vv = e_n - ch * p.Vpad;
            if (ch < p.C && vv < p.V) p.logits[((size_t)e_r * p.C + ch) * p.V + vv] = y;
        }
        return;
    }
    // residual add (dia/layers.py:555,574,582); the new stream is also written as the parts of
    // x * w_norm for the NEXT consumer, and its sum of squares is collected for that consumer's RMSNorm
    float xn = 0.f;
    if (e_valid) {
        xn = resid_v + total(e_m);
        reinterpret_cast<float*>(p.x)[(size_t)e_n * 2 + e_r] = xn;
        const float* wn = gt == G_SO ? p.norms + ((size_t)layer * 3 + 1) * p.D
                        : gt == G_CO ? p.norms + ((size_t)layer * 3 + 2) * p.D
                                     : p.norms + ((size_t)(layer + 1) * 3) * p.D;   // next layer's pre_sa_norm, or
        store_parts(p.xparts, e_n, e_r, xn * __ldg(wn + e_n));                      // the final norm after layer L-1
    }
    float sq = xn * xn;                                                // half-warps = batch rows
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((lane & 15) == 0) c.misc->ssq_part[c.warp][e_r] = sq;
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

template <int HPK, bool kSelf>
__device__ void attn_stage(Ctx& c, int layer, int pos, int slot) {
    const StepParams& p = *c.p;
    const int cta = blockIdx.x;
    const AttnWork w = kSelf ? self_attn_work(p, cta, slot) : cross_attn_work(p, cta);
    if (!w.active) return;
    const int r = kSelf ? w.pair / p.Hkv : 1;
    const int kvh = kSelf ? w.pair - r * p.Hkv : w.pair;
    const int head0 = kSelf ? kvh * HPK : w.pair;
    const int nsplit = kSelf ? p.sa_nsplit : p.ca_nsplit;

    float* qs = reinterpret_cast<float*>(c.xs);          // [HPK][128] rotated, pre-scaled queries
    float* kn = qs + HPK * kHeadDim;                     // [128] rotated key of this step
    float* vn = kn + kHeadDim;                           // [128] value of this step
    float* stat = vn + kHeadDim;                         // m[HPK] at 0.., l[HPK] at 8..
    float* sc = stat + 16;                               // [(n_keys + 1)][HPK] scores -> probabilities
    float* racc = reinterpret_cast<float*>(c.xs) + 8192; // [warps][HPK][128]
    const float2* qsrc = kSelf ? p.qkv : p.cq;
    const int pclamp = min(pos, p.n_pos - 1);
    const float* sinr = p.rope_sin + (size_t)pclamp * 64;
    const float* cosr = p.rope_cos + (size_t)pclamp * 64;
    const float scale = 0.08838834764831845f;           // 1/sqrt(128)

    for (int i = c.tid; i < HPK * 64; i += kConsumerThreads) {
        const int h = i >> 6, d = i & 63;
        const float* q2 = reinterpret_cast<const float*>(qsrc + (size_t)(head0 + h) * kHeadDim);
        const float a = ldcg_f(q2 + 2 * d + r), b = ldcg_f(q2 + 2 * (d + 64) + r);
        const float sn = __ldg(sinr + d), cs = __ldg(cosr + d);
        qs[h * kHeadDim + d] = (a * cs - b * sn) * scale;
        qs[h * kHeadDim + d + 64] = (a * sn + b * cs) * scale;
    }
    if (kSelf && w.has_new) {
        for (int d = c.tid; d < 64; d += kConsumerThreads) {
            const float* k2 = reinterpret_cast<const float*>(p.qkv + (size_t)(p.Hq + kvh) * kHeadDim);
            const float* v2 = reinterpret_cast<const float*>(p.qkv + (size_t)(p.Hq + p.Hkv + kvh) * kHeadDim);
            const float a = ldcg_f(k2 + 2 * d + r), b = ldcg_f(k2 + 2 * (d + 64) + r);
            const float sn = __ldg(sinr + d), cs = __ldg(cosr + d);
            kn[d] = a * cs - b * sn;
            kn[d + 64] = a * sn + b * cs;
            vn[d] = ldcg_f(v2 + 2 * d + r);
            vn[d + 64] = ldcg_f(v2 + 2 * (d + 64) + r);
        }
    }
    consumer_sync();
    if (kSelf && w.has_new) {
        // KVCache.update (dia/state.py:99-103): append this step's K/V at `slot`
        const size_t row = ((size_t)w.pair * p.Lmax + slot) * kHeadDim;
        for (int d = c.tid; d < kHeadDim; d += kConsumerThreads) {
            p.self_k[layer][row + d] = kn[d];
            p.self_v[layer][row + d] = vn[d];
        }
    }
    float4 q[HPK];
#pragma unroll
    for (int h = 0; h < HPK; ++h) q[h] = reinterpret_cast<const float4*>(qs + h * kHeadDim)[c.lane];

    const int nk = w.k_hi - w.k_lo;
    const int nkc = (nk + 15) >> 4;

    // ---- K pass: scores --------------------------------------------------------------------
    for (int ci = c.warp; ci < nkc; ci += kConsumerWarps) {
        const unsigned idx = c.cbase + ci;
        const unsigned sl = idx % kNumSlots;
        mbar_wait(&c.misc->full[sl], (idx / kNumSlots) & 1u, p.err, kErrFullBarrierTimeout);
        const int keys_in = min(16, nk - ci * 16);
        const float4* kt = reinterpret_cast<const float4*>(c.ring + sl * kSlotBytes) + c.lane;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float v[8 * HPK];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int key = half * 8 + i;
                float4 kv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (key < keys_in) kv = kt[key * 32];
#pragma unroll
                for (int h = 0; h < HPK; ++h)
                    v[i * HPK + h] = kv.x * q[h].x + kv.y * q[h].y + kv.z * q[h].z + kv.w * q[h].w;
            }
            transpose_reduce<8 * HPK>(v, c.lane);
            // HPK = 4: lane l holds (key l/4, head l%4); HPK = 1: lanes 4k..4k+3 hold key k
            const int key = half * 8 + (HPK == 4 ? (c.lane >> 2) : (c.lane >> 2));
            const int h = HPK == 4 ? (c.lane & 3) : 0;
            const bool writer = HPK == 4 ? true : ((c.lane & 3) == 0);
            if (writer && key < keys_in) sc[(ci * 16 + key) * HPK + h] = v[0];
        }
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&c.misc->empty[sl]);
    }
    if (kSelf && w.has_new && c.warp == 0) {
        const float4 kv = reinterpret_cast<const float4*>(kn)[c.lane];
#pragma unroll
        for (int h = 0; h < HPK; ++h) {
            float s = kv.x * q[h].x + kv.y * q[h].y + kv.z * q[h].z + kv.w * q[h].w;
            s = warp_sum(s);
            if (c.lane == 0) sc[nk * HPK + h] = s;
        }
    }
    consumer_sync();
    const int n_tot = nk + ((kSelf && w.has_new) ? 1 : 0);
    if (c.warp < HPK) {
        const int h = c.warp;
        float m = -INFINITY;
        for (int i = c.lane; i < n_tot; i += 32) m = fmaxf(m, sc[i * HPK + h]);
        m = warp_max(m);
        float l = 0.f;
        for (int i = c.lane; i < n_tot; i += 32) {
            const float e = expf(sc[i * HPK + h] - m);
            sc[i * HPK + h] = e;
            l += e;
        }
        l = warp_sum(l);
        if (c.lane == 0) { stat[h] = m; stat[8 + h] = l; }
    }
    consumer_sync();

    // ---- V pass ----------------------------------------------------------------------------
    float4 acc[HPK];
#pragma unroll
    for (int h = 0; h < HPK; ++h) acc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ci = c.warp; ci < nkc; ci += kConsumerWarps) {
        const unsigned idx = c.cbase + nkc + ci;
        const unsigned sl = idx % kNumSlots;
        mbar_wait(&c.misc->full[sl], (idx / kNumSlots) & 1u, p.err, kErrFullBarrierTimeout);
        const int keys_in = min(16, nk - ci * 16);
        const float4* vt = reinterpret_cast<const float4*>(c.ring + sl * kSlotBytes) + c.lane;
        for (int key = 0; key < keys_in; ++key) {
            const float4 vv = vt[key * 32];
#pragma unroll
            for (int h = 0; h < HPK; ++h) {
                const float pr = sc[(ci * 16 + key) * HPK + h];
                acc[h].x = fmaf(pr, vv.x, acc[h].x); acc[h].y = fmaf(pr, vv.y, acc[h].y);
                acc[h].z = fmaf(pr, vv.z, acc[h].z); acc[h].w = fmaf(pr, vv.w, acc[h].w);
            }
        }
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&c.misc->empty[sl]);
    }
    c.cbase += 2 * nkc;
    if (kSelf && w.has_new && c.warp == 0) {
        const float4 vv = reinterpret_cast<const float4*>(vn)[c.lane];
#pragma unroll
        for (int h = 0; h < HPK; ++h) {
            const float pr = sc[nk * HPK + h];
            acc[h].x = fmaf(pr, vv.x, acc[h].x); acc[h].y = fmaf(pr, vv.y, acc[h].y);
            acc[h].z = fmaf(pr, vv.z, acc[h].z); acc[h].w = fmaf(pr, vv.w, acc[h].w);
        }
    }
#pragma unroll
    for (int h = 0; h < HPK; ++h)
        reinterpret_cast<float4*>(racc + ((size_t)c.warp * HPK + h) * kHeadDim)[c.lane] = acc[h];
    consumer_sync();

    float* outf = reinterpret_cast<float*>(kSelf ? p.attn : p.cattn);
    unsigned short* oparts = kSelf ? p.attn_parts : p.cattn_parts;
    float* part = kSelf ? p.sa_part : p.ca_part;
    for (int i = c.tid; i < HPK * kHeadDim; i += kConsumerThreads) {
        const int h = i >> 7, d = i & 127;
        float o = 0.f;
#pragma unroll
        for (int ww = 0; ww < kConsumerWarps; ++ww) o += racc[((size_t)ww * HPK + h) * kHeadDim + d];
        const int head = head0 + h;
        if (w.n_active == 1) {
            const float l = stat[8 + h];
            const float val = l > 0.f ? o / l : 0.f;
            if (kSelf) outf[((size_t)head * kHeadDim + d) * 2 + r] = val;
            else reinterpret_cast<float2*>(outf)[(size_t)head * kHeadDim + d] = make_float2(0.f, val);
            store_parts(oparts, head * kHeadDim + d, r, val);      // row 0 of the cross output stays all-zero
        } else {
            float* pp = part + (((size_t)(kSelf ? r * p.Hq + head : head)) * nsplit + w.split) * 132;
            pp[4 + d] = o;
            if (d == 0) { pp[0] = stat[h]; pp[1] = stat[8 + h]; }
        }
    }
    if (w.n_active == 1) return;

    // ---- last-arriving split of this pair combines all splits (fixed order => deterministic) ----
    __threadfence();
    consumer_sync();
    unsigned* cnt = p.pair_cnt + (kSelf ? w.pair : 2 * p.Hkv + w.pair);
    if (c.tid == 0) {
        const unsigned old = atomicAdd(cnt, 1u);
        const int last = old == (unsigned)(w.n_active - 1);
        if (last) { *cnt = 0u; __threadfence(); }
        c.misc->flag = last;
    }
    consumer_sync();
    if (!c.misc->flag) return;
    for (int i = c.tid; i < HPK * kHeadDim; i += kConsumerThreads) {
        const int h = i >> 7, d = i & 127;
        const int head = head0 + h;
        const float* pb = part + ((size_t)(kSelf ? r * p.Hq + head : head)) * nsplit * 132;
        float M = -INFINITY;
        for (int s = 0; s < w.n_active; ++s) M = fmaxf(M, ldcg_f(pb + (size_t)s * 132));
        float Lsum = 0.f, O = 0.f;
        for (int s = 0; s < w.n_active; ++s) {
            const float f = expf(ldcg_f(pb + (size_t)s * 132) - M);
            Lsum = fmaf(ldcg_f(pb + (size_t)s * 132 + 1), f, Lsum);
            O = fmaf(ldcg_f(pb + (size_t)s * 132 + 4 + d), f, O);
        }
        const float val = Lsum > 0.f ? O / Lsum : 0.f;
        if (kSelf) outf[((size_t)head * kHeadDim + d) * 2 + r] = val;
        else reinterpret_cast<float2*>(outf)[(size_t)head * kHeadDim + d] = make_float2(0.f, val);
        store_parts(oparts, head * kHeadDim + d, r, val);
    }
}

// ------------------------------------------------------------------------------------------
// embedding gather-sum (dia/layers.py:691-696): x = ((e0 + e1) + e2) ... + e8, both CFG rows
// ------------------------------------------------------------------------------------------
__device__ void embed_stage(Ctx& c, int pos) {
    const StepParams& p = *c.p;
    const int d = blockIdx.x * kConsumerThreads + c.tid;
    float s0 = 0.f, s1 = 0.f;
    if (d < p.D) {
        const int* t0;
        const int* t1;
        if (p.tokens != nullptr) { t0 = p.tokens; t1 = p.tokens + p.C; }
        else { t0 = t1 = p.grid + (size_t)(pos - 1) * p.C; }
        for (int ch = 0; ch < p.C; ++ch) {
            int a = ldcg_i(t0 + ch), b = ldcg_i(t1 + ch);
            if (a < 0 || a >= p.V || b < 0 || b >= p.V) {
                // steps executed after the utterance finished read unwritten (-1) grid rows: harmless no-ops
                const bool dead = p.tokens == nullptr && p.gs != nullptr && ldcg_i(&p.gs->finished) != 0;
                if (!dead) *p.err = kErrBadState;
                a = 0; b = 0;
            }
            const float* tab = p.emb + (size_t)ch * p.V * p.D;
            const float e0 = __ldg(tab + (size_t)a * p.D + d), e1 = __ldg(tab + (size_t)b * p.D + d);
            s0 = ch == 0 ? e0 : s0 + e0;
            s1 = ch == 0 ? e1 : s1 + e1;
        }
        p.x[d] = make_float2(s0, s1);
        const float w = __ldg(p.norms + d);                 // layer 0 pre_sa_norm: the first consumer of x
        store_parts(p.xparts, d, 0, s0 * w);
        store_parts(p.xparts, d, 1, s1 * w);
    }
    const float q0 = warp_sum(s0 * s0), q1 = warp_sum(s1 * s1);
    if (c.lane == 0) { c.misc->ssq_part[c.warp][0] = q0; c.misc->ssq_part[c.warp][1] = q1; }
}

// ------------------------------------------------------------------------------------------
// sampling: CFG + masks + argmax / top-k / top-p / multinomial (dia/model.py:447-488, 32-82)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t (&ctr)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
        const uint32_t n0 = hi1 ^ ctr[1] ^ k0, n1 = lo1, n2 = hi0 ^ ctr[3] ^ k1, n3 = lo0;
        ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

constexpr int kMaxCand = 64;

// one warp, one channel.  g: guided + masked logits [V] in shared memory (scaled and consumed IN
// PLACE on the sampling path); cv/ci: candidate lists [kMaxCand].  Returns the token id (all lanes).
__device__ int sample_channel(float* g, int V, float temperature, float top_p, int top_k,
                              unsigned long long seed, unsigned long long draw, int ch, float* probs_out,
                              float* cv, int* ci, int lane) {
    float* work = g;
    if (temperature == 0.0f) {          // torch.argmax: first maximal index
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = lane; i < V; i += 32) {
            const float x = g[i];
            if (x > bv || bi == 0x7fffffff) { bv = x; bi = i; }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, m);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        return bi;
    }
    for (int i = lane; i < V; i += 32) work[i] = work[i] / temperature;
    __syncwarp();
    // ---- top-k: extract maxima in descending order; keep ties with the k-th value ------------
    int ncand = 0;
    float kth = 0.f;
    const int k = top_k;
    while (ncand < kMaxCand) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = lane; i < V; i += 32) {
            const float x = work[i];
            if (x > bv) { bv = x; bi = i; }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, m);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (bi == 0x7fffffff) break;                         // nothing finite left
        if (ncand >= k && !(bv == kth)) break;               // past k and not tied with the k-th
        if (lane == 0) { cv[ncand] = bv; ci[ncand] = bi; work[bi] = -INFINITY; }
        ncand++;
        if (ncand == k) kth = bv;
        __syncwarp();
    }
    __syncwarp();
    // ---- softmax over the survivors, top-p on the sorted cumulative sum -----------------------
    const float mx = cv[0];
    float e0 = lane < ncand ? expf(cv[lane] - mx) : 0.f;
    float e1 = lane + 32 < ncand ? expf(cv[lane + 32] - mx) : 0.f;
    float Z = warp_sum(e0 + e1);
    int nkeep = ncand;
    if (top_p < 1.0f) {
        if (lane == 0) {
            float cum = 0.f;
            nkeep = 0;
            for (int i = 0; i < ncand; ++i) {
                // entry i is removed iff the cumulative probability BEFORE it already exceeds top_p
                if (i > 0 && cum > top_p) break;
                cum += expf(cv[i] - mx) / Z;
                nkeep = i + 1;
            }
        }
        nkeep = __shfl_sync(0xffffffffu, nkeep, 0);
    }
    e0 = lane < nkeep ? e0 : 0.f;
    e1 = lane + 32 < nkeep ? e1 : 0.f;
    const float Z2 = warp_sum(e0 + e1);
    if (probs_out != nullptr) {
        for (int i = lane; i < V; i += 32) probs_out[i] = 0.f;
        __syncwarp();
        if (lane < nkeep) probs_out[ci[lane]] = e0 / Z2;
        if (lane + 32 < nkeep) probs_out[ci[lane + 32]] = e1 / Z2;
    }
    // ---- multinomial(1): inverse CDF over the survivors with a Philox uniform -------------------
    int tok = ci[0];
    if (lane == 0) {
        uint32_t ctr[4] = {(uint32_t)draw, (uint32_t)(draw >> 32), (uint32_t)ch, 0x44494131u};
        philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
        const float u = (float)(ctr[0] >> 8) * (1.0f / 16777216.0f);   // [0, 1)
        const float target = u * Z2;
        float cum = 0.f;
        tok = ci[nkeep - 1];
        for (int i = 0; i < nkeep; ++i) {
            cum += expf(cv[i] - mx);
            if (cum > target) { tok = ci[i]; break; }
        }
    }
    return __shfl_sync(0xffffffffu, tok, 0);
}

// guided = cond + s * (cond - uncond), then the -inf masks (dia/model.py:450-478)
__device__ __forceinline__ float guided_logit(float un, float co, float s, int ch, int v, int V, int eos, int pad,
                                              int bos) {
    float gv = __fadd_rn(co, __fmul_rn(s, __fsub_rn(co, un)));
    if ((ch > 0 && v == eos) || v == pad || v == bos) gv = -INFINITY;
    if (V <= eos + 1 && v >= V) gv = -INFINITY;
    return gv;
}

// shared by the in-kernel sampling stage and the standalone head_sample kernel.
// smem: gbuf [C*V] floats, candidate lists [warps][kMaxCand]; preds int[C]
__device__ void sample_all_channels(const StepParams& p, const float* logits, unsigned long long draw, float* gbuf,
                                    float* cv, int* ci, int* preds, float* probs_out, int tid, int warp,
                                    int lane, int nwarps) {
    const int CV = p.C * p.V;
    for (int i = tid; i < CV; i += nwarps * 32) {
        const int ch = i / p.V, v = i - ch * p.V;
        gbuf[i] = guided_logit(ldcg_f(logits + i), ldcg_f(logits + CV + i), p.cfg_scale, ch, v, p.V, p.eos, p.pad,
                               p.bos);
    }
    asm volatile("bar.sync 1, %0;" ::"r"(nwarps * 32) : "memory");
    for (int ch = warp; ch < p.C; ch += nwarps) {
        const int t = sample_channel(gbuf + (size_t)ch * p.V, p.V, p.temperature, p.top_p, p.top_k, p.seed, draw, ch,
                                     probs_out ? probs_out + (size_t)ch * p.V : nullptr,
                                     cv + warp * kMaxCand, ci + warp * kMaxCand, lane);
        if (lane == 0) preds[ch] = t;
    }
    asm volatile("bar.sync 1, %0;" ::"r"(nwarps * 32) : "memory");
}

__device__ void sample_stage(Ctx& c, int step_index, int pos) {
    const StepParams& p = *c.p;
    if (blockIdx.x != 0) return;
    float* gbuf = reinterpret_cast<float*>(c.xs);                       // C*V floats (37 KB of the 64 KB)
    float* cv = c.red;                                                  // [warps][64]
    int* ci = reinterpret_cast<int*>(c.red + kConsumerWarps * kMaxCand);
    int* preds = reinterpret_cast<int*>(c.red + 2 * kConsumerWarps * kMaxCand);
    const unsigned long long draw = (unsigned long long)(p.gs ? p.gs->steps_run : step_index);
    sample_all_channels(p, p.logits, draw, gbuf, cv, ci, preds, p.probs_out, c.tid, c.warp, c.lane,
                        kConsumerWarps);
    if (c.tid != 0) return;
    for (int ch = 0; ch < p.C; ++ch) p.pred_out[ch] = preds[ch];
    GenState* gs = p.gs;
    if (gs == nullptr || p.grid == nullptr) return;
    // ---- the body of the reference's while loop after _decoder_step (dia/model.py:771-807) ----
    if (!gs->finished) {
        if (gs->dec_step >= p.max_tokens - 1) {
            gs->finished = 1;
        } else {
            int dmax = 0;
            for (int ch = 0; ch < p.C; ++ch) dmax = max(dmax, p.delay[ch]);
            const int cur = gs->dec_step + 1;
            if (cur != pos) *p.err = kErrBadState;
            int pr[DIA_B200_MAX_CHANNELS];
            for (int ch = 0; ch < p.C; ++ch) pr[ch] = preds[ch];
            if (!gs->eos_detected && pr[0] == p.eos) { gs->eos_detected = 1; gs->eos_countdown = dmax; }
            if (gs->eos_countdown > 0) {
                const int s = dmax - gs->eos_countdown;
                for (int ch = 0; ch < p.C; ++ch) {
                    if (s == p.delay[ch]) pr[ch] = p.eos;
                    else if (s > p.delay[ch] && pr[ch] != p.eos) pr[ch] = p.pad;
                }
                gs->eos_countdown -= 1;
            }
            gs->bos_countdown = max(0, gs->bos_countdown - 1);
            int* row = p.grid + (size_t)cur * p.C;
            for (int ch = 0; ch < p.C; ++ch) {
                if (gs->bos_countdown > 0) { if (row[ch] == -1) row[ch] = pr[ch]; }   // update_one(apply_mask=True)
                else row[ch] = pr[ch];
            }
            if (gs->eos_countdown == 0) {
                gs->finished = 1;                                   // break: dec_step is NOT advanced
            } else {
                if (cur >= p.max_tokens - dmax - 1 && !gs->eos_detected) {
                    gs->eos_detected = 1;
                    gs->eos_countdown = dmax;
                }
                gs->dec_step += 1;
                if (gs->dec_step >= p.max_tokens - 1) gs->finished = 1;
            }
        }
    }
    gs->steps_run += 1;
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(kThreads, 1) dia_step_kernel(const __grid_constant__ StepParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* ring = smem;
    float2* xs = reinterpret_cast<float2*>(smem + kNumSlots * kSlotBytes);
    float* red = reinterpret_cast<float*>(smem + kNumSlots * kSlotBytes + kXsBytes);
    SharedMisc* misc = reinterpret_cast<SharedMisc*>(smem + kNumSlots * kSlotBytes + kXsBytes + kRedBytes);

    const int tid = threadIdx.x;
    // a launch queued behind the one that finished the utterance is a no-op.  Uniform: `finished`
    // is only written by the sampler after >100 grid barriers, i.e. after every CTA has read it.
    if (p.gs != nullptr && p.grid != nullptr && ldcg_i(&p.gs->finished) != 0) return;
    if (tid == 0) {
        for (int i = 0; i < kNumSlots; ++i) { mbar_init(&misc->full[i], 1); mbar_init(&misc->empty[i], 1); }
        misc->stages_done = 0;
        misc->flag = 0;
        fence_mbar_init();
    }
    {   // copy this CTA's table
        const int* src = reinterpret_cast<const int*>(p.cta_tab + blockIdx.x);
        int* dst = reinterpret_cast<int*>(&misc->tab);
        for (int i = tid; i < (int)(sizeof(CtaTable) / 4); i += kThreads) dst[i] = src[i];
    }
    __syncthreads();
    if (tid < G_COUNT) {
        GemmCfg& g = misc->gcfg[tid];
        g.gc = misc->tab.gc[tid]; g.g0 = misc->tab.g0[tid]; g.K = p.Kdim[tid];
        g.row_bytes = g.gc * 16;
        g.rpc = g.gc > 0 ? min(kMaxKb * 16, (kSlotBytes / (g.gc * 16)) & ~15) : 16;
        g.n_chunks = g.gc > 0 ? (g.K + g.rpc - 1) / g.rpc : 0;
        g.n_mt = (g.gc + 1) >> 1;
    }
    __syncthreads();

    if (tid >= kConsumerThreads) {
        if (tid == kConsumerThreads) producer_loop(p, ring, misc);
        return;
    }

    Ctx c;
    c.p = &p; c.ring = ring; c.xs = xs; c.red = red; c.misc = misc;
    c.tid = tid; c.warp = tid >> 5; c.lane = tid & 31;
    c.cbase = 0; c.nbar = 0;
    const int S = 8 * p.L + 3;

    for (int n = 0; n < p.n_steps; ++n) {
        const int pos = p.pos0 + n, slot = p.slot0 + n;
        for (int s = p.stage_begin; s < p.stage_end; ++s) {
            int kind, layer;
            decode_stage(s, p.L, kind, layer);
            c.tstamp = (p.timing != nullptr && blockIdx.x == 0 && tid == 0) ? p.timing + ((size_t)n * S + s) * 8 : nullptr;
            if (c.tstamp) c.tstamp[0] = clock64();
            switch (kind) {
                case S_EMBED: embed_stage(c, pos); break;
                case S_SATTN: attn_stage<4, true>(c, layer, pos, slot); break;
                case S_CATTN: attn_stage<1, false>(c, layer, pos, slot); break;
                case S_SAMPLE: sample_stage(c, n, pos); break;
                default: gemm_stage(c, gemm_of_kind(kind), layer); break;
            }
            const bool last = (n == p.n_steps - 1) && (s == p.stage_end - 1);
            const bool tm = p.timing != nullptr && blockIdx.x == 0 && tid == 0;
            if (tm) p.timing[((size_t)n * S + s) * 8 + 4] = clock64();
            const bool publish = kind == S_EMBED || kind == S_SO || kind == S_CO || kind == S_WO;
            if (!last) grid_barrier(c, publish);
            else if (publish) { consumer_sync(); if (tid == 0) publish_ssq_partials(c); }
            if (tm) p.timing[((size_t)n * S + s) * 8 + 5] = clock64();
            if (tid == 0) st_release_cta_s32(&misc->stages_done, n * S + s + 1);
        }
    }
}

// standalone head: CFG + masks + sampling on caller-provided logits (dia/model.py:447-488)
extern "C" __global__ void __launch_bounds__(kConsumerThreads, 1)
dia_head_sample_kernel(const __grid_constant__ StepParams p, const float* logits, unsigned long long draw, int* pred,
                       float* probs) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* gbuf = reinterpret_cast<float*>(smem);
    float* cv = gbuf + ((p.C * p.V + 3) & ~3);
    int* ci = reinterpret_cast<int*>(cv + kConsumerWarps * kMaxCand);
    int* preds = ci + kConsumerWarps * kMaxCand;
    const int tid = threadIdx.x;
    sample_all_channels(p, logits, draw, gbuf, cv, ci, preds, probs, tid, tid >> 5, tid & 31, kConsumerWarps);
    if (tid < p.C) pred[tid] = preds[tid];
}

int step_kernel_smem_bytes() { return kSmemBytes; }

cudaError_t launch_step_kernel(const StepParams& p, bool cooperative, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(dia_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (cooperative) {
        void* args[] = {const_cast<StepParams*>(&p)};
        return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(dia_step_kernel), dim3(p.G), dim3(kThreads),
                                           args, kSmemBytes, st);
    }
    dia_step_kernel<<<p.G, kThreads, kSmemBytes, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_head_sample(const StepParams& p, const float* logits, unsigned long long draw, int* pred,
                               float* probs, cudaStream_t st) {
    const int smem = ((p.C * p.V + 3) & ~3) * 4 + kConsumerWarps * kMaxCand * 8 + 64 * 4;
    static int attr_bytes = 0;
    if (smem > attr_bytes) {
        cudaError_t e = cudaFuncSetAttribute(dia_head_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_bytes = smem;
    }
    dia_head_sample_kernel<<<1, kConsumerThreads, smem, st>>>(p, logits, draw, pred, probs);
    return cudaGetLastError();
}

}  // namespace dia
