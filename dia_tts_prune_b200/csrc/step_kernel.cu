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
#include "common.cuh"
#include "engine_internal.h"

namespace dia {

struct SharedMisc {
    uint64_t full[kNumSlots];
    uint64_t empty[kNumSlots];
    CtaTable tab;
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

// rows of a slab that fit one ring slot, rounded down to a multiple of the rows a warp
// covers per iteration (R = 32 / gc lanes-groups)
__device__ __forceinline__ int rows_per_chunk(int gc, int R) {
    int rpc = kSlotBytes / (gc * 16);
    return rpc - rpc % R;
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
__device__ __forceinline__ void grid_barrier(Ctx& c) {
    const StepParams& p = *c.p;
    consumer_sync();
    c.nbar++;
    if (c.tid == 0) {
        __threadfence();
        red_release_add_u32(p.grid_bar, 1u);
        const unsigned target = c.nbar * gridDim.x;
        if (ld_acquire_u32(p.grid_bar) < target) {
            const unsigned long long t0 = clock64();
            while (ld_acquire_u32(p.grid_bar) < target) {
                if (clock64() - t0 > kWatchdogCycles) {
                    *reinterpret_cast<volatile int*>(p.err) = kErrGridBarrierTimeout;
                    __threadfence_system();
                    __trap();
                }
            }
        }
        __threadfence();
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
                const int gc = tab.gc[gt];
                if (gc == 0) continue;
                const int K = p.Kdim[gt];
                const int R = gc > 16 ? 1 : 32 / gc;
                const int rpc = rows_per_chunk(gc, R);
                const uint32_t row_bytes = gc * 16;
                const unsigned char* base = p.wstream + tab.stream_base +
                    (gt == G_LOGITS ? tab.logits_off
                                    : (unsigned long long)layer * tab.layer_bytes + tab.slab_off[gt]);
                for (int r0 = 0; r0 < K; r0 += rpc) {
                    const int rows = min(rpc, K - r0);
                    pr.issue(base + (size_t)r0 * row_bytes, rows * row_bytes, false);
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
// GEMM stage: y[2][N_cta] = xs[2][K] . W_slab, fp32 accumulate, fused prologue / epilogue
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void fma_row(f32x2 (&accA)[4], f32x2 (&accB)[4], const uint4& w, const float2& x) {
    const f32x2 xa = pack2(x.x, x.x), xb = pack2(x.y, x.y);
    const f32x2 w0 = bf16x2_to_f32x2(w.x), w1 = bf16x2_to_f32x2(w.y), w2 = bf16x2_to_f32x2(w.z),
                w3 = bf16x2_to_f32x2(w.w);
    accA[0] = ffma2(xa, w0, accA[0]); accB[0] = ffma2(xb, w0, accB[0]);
    accA[1] = ffma2(xa, w1, accA[1]); accB[1] = ffma2(xb, w1, accB[1]);
    accA[2] = ffma2(xa, w2, accA[2]); accB[2] = ffma2(xb, w2, accB[2]);
    accA[3] = ffma2(xa, w3, accA[3]); accB[3] = ffma2(xb, w3, accB[3]);
}

// load the stage's input vector [K][2] into xs, optionally RMS-normalised (fp32, eps, weight)
__device__ void load_vector(Ctx& c, const float2* src, int K, const float* normw) {
    const StepParams& p = *c.p;
    float4* xs4 = reinterpret_cast<float4*>(c.xs);
    const float4* s4 = reinterpret_cast<const float4*>(src);
    const int n4 = K >> 1;
    float ss0 = 0.f, ss1 = 0.f;
    for (int i = c.tid; i < n4; i += kConsumerThreads) {
        const float4 v = ldcg_f4(s4 + i);
        xs4[i] = v;
        ss0 = fmaf(v.x, v.x, ss0); ss0 = fmaf(v.z, v.z, ss0);
        ss1 = fmaf(v.y, v.y, ss1); ss1 = fmaf(v.w, v.w, ss1);
    }
    if (normw != nullptr) {
        ss0 = warp_sum(ss0);
        ss1 = warp_sum(ss1);
        if (c.lane == 0) { c.misc->stat[c.warp] = ss0; c.misc->stat[8 + c.warp] = ss1; }
        consumer_sync();
        float t0 = 0.f, t1 = 0.f;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; ++w) { t0 += c.misc->stat[w]; t1 += c.misc->stat[8 + w]; }
        const float inv0 = 1.0f / sqrtf(t0 / (float)K + p.eps);
        const float inv1 = 1.0f / sqrtf(t1 / (float)K + p.eps);
        const float2* w2 = reinterpret_cast<const float2*>(normw);
        for (int i = c.tid; i < n4; i += kConsumerThreads) {
            float4 v = xs4[i];
            const float2 w = __ldg(w2 + i);
            v.x = (v.x * inv0) * w.x; v.y = (v.y * inv1) * w.x;
            v.z = (v.z * inv0) * w.y; v.w = (v.w * inv1) * w.y;
            xs4[i] = v;
        }
    }
    consumer_sync();
}

__device__ void gemm_stage(Ctx& c, int gt, int layer) {
    const StepParams& p = *c.p;
    const CtaTable& tab = c.misc->tab;
    const int gc = tab.gc[gt];
    if (gc == 0) return;
    const int g0 = tab.g0[gt];
    const int K = p.Kdim[gt];

    const float2* src;
    const float* normw = nullptr;
    switch (gt) {
        case G_QKV: src = p.x; normw = p.norms + ((size_t)layer * 3 + 0) * p.D; break;
        case G_SO: src = p.attn; break;
        case G_CQ: src = p.x; normw = p.norms + ((size_t)layer * 3 + 1) * p.D; break;
        case G_CO: src = p.cattn; break;
        case G_WI: src = p.x; normw = p.norms + ((size_t)layer * 3 + 2) * p.D; break;
        case G_WO: src = p.hidden; break;
        default: src = p.x; normw = p.norms + (size_t)p.L * 3 * p.D; break;   // logits: final norm
    }
    load_vector(c, src, K, normw);

    const int R = gc > 16 ? 1 : 32 / gc;
    const int rpc = rows_per_chunk(gc, R);
    const int row_bytes = gc * 16;
    const int n_chunks = (K + rpc - 1) / rpc;
    const int j = c.lane / gc, g = c.lane - j * gc;
    const bool lane_active = j < R;
    const int lane_off = j * row_bytes + g * 16;
    const int it_bytes = R * row_bytes;

    f32x2 accA[4], accB[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { accA[i] = 0ull; accB[i] = 0ull; }

    for (int ci = c.warp; ci < n_chunks; ci += kConsumerWarps) {
        const unsigned idx = c.cbase + ci;
        const unsigned slot = idx % kNumSlots;
        mbar_wait(&c.misc->full[slot], (idx / kNumSlots) & 1u, p.err, kErrFullBarrierTimeout);
        const int rows = min(rpc, K - ci * rpc);
        if (lane_active) {
            const unsigned char* sp = c.ring + slot * kSlotBytes + lane_off;
            const float2* xp = c.xs + ci * rpc + j;
            const int full = rows / R;
            int it = 0;
            for (; it + 4 <= full; it += 4) {
                const uint4 w0 = *reinterpret_cast<const uint4*>(sp + (it + 0) * it_bytes);
                const uint4 w1 = *reinterpret_cast<const uint4*>(sp + (it + 1) * it_bytes);
                const uint4 w2 = *reinterpret_cast<const uint4*>(sp + (it + 2) * it_bytes);
                const uint4 w3 = *reinterpret_cast<const uint4*>(sp + (it + 3) * it_bytes);
                const float2 x0 = xp[(it + 0) * R], x1 = xp[(it + 1) * R], x2 = xp[(it + 2) * R],
                             x3 = xp[(it + 3) * R];
                fma_row(accA, accB, w0, x0);
                fma_row(accA, accB, w1, x1);
                fma_row(accA, accB, w2, x2);
                fma_row(accA, accB, w3, x3);
            }
            for (; it < full; ++it) {
                const uint4 w0 = *reinterpret_cast<const uint4*>(sp + it * it_bytes);
                fma_row(accA, accB, w0, xp[it * R]);
            }
            if (j < rows - full * R) {
                const uint4 w0 = *reinterpret_cast<const uint4*>(sp + full * it_bytes);
                fma_row(accA, accB, w0, xp[full * R]);
            }
        }
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&c.misc->empty[slot]);
    }
    c.cbase += n_chunks;

    // ---- reduce: over the R row-lanes of a warp by shuffles, then over warps through smem ----
    float v[16];   // e = row*8 + col
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        unpack2(accA[i], v[2 * i], v[2 * i + 1]);
        unpack2(accB[i], v[8 + 2 * i], v[8 + 2 * i + 1]);
    }
    if (R > 1) {
        int off = 1;
        while (off * 2 < R) off *= 2;
        for (; off >= 1; off >>= 1) {
            const int delta = off * gc;
            const bool ok = c.lane + delta < 32;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float t = __shfl_down_sync(0xffffffffu, v[e], delta);
                if (ok) v[e] += t;
            }
        }
    }
    if (c.lane < gc) {
        float4* r4 = reinterpret_cast<float4*>(c.red + ((size_t)c.warp * gc + c.lane) * 16);
        r4[0] = make_float4(v[0], v[1], v[2], v[3]);
        r4[1] = make_float4(v[4], v[5], v[6], v[7]);
        r4[2] = make_float4(v[8], v[9], v[10], v[11]);
        r4[3] = make_float4(v[12], v[13], v[14], v[15]);
    }
    consumer_sync();

    // ---- epilogue ----------------------------------------------------------------------------
    if (gt == G_WI) {
        // groups alternate (gate, up) of the same 8 hidden units: h = silu(gate) * up (dia/layers.py:95-101)
        const int n_out = (gc >> 1) * 16;
        for (int t = c.tid; t < n_out; t += kConsumerThreads) {
            const int hg = t >> 4, e = t & 15;
            float gate = 0.f, up = 0.f;
#pragma unroll
            for (int w = 0; w < kConsumerWarps; ++w) {
                gate += c.red[((size_t)w * gc + 2 * hg) * 16 + e];
                up += c.red[((size_t)w * gc + 2 * hg + 1) * 16 + e];
            }
            const float h = (gate / (1.0f + expf(-gate))) * up;
            const int n = ((g0 >> 1) + hg) * 8 + (e & 7);
            reinterpret_cast<float*>(p.hidden)[(size_t)n * 2 + (e >> 3)] = h;
        }
        return;
    }
    const int n_out = gc * 16;
    for (int t = c.tid; t < n_out; t += kConsumerThreads) {
        const int gg = t >> 4, e = t & 15;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; ++w) s += c.red[((size_t)w * gc + gg) * 16 + e];
        const int n = (g0 + gg) * 8 + (e & 7);
        const int r = e >> 3;
        if (gt == G_QKV) {
            reinterpret_cast<float*>(p.qkv)[(size_t)n * 2 + r] = s;
        } else if (gt == G_CQ) {
            reinterpret_cast<float*>(p.cq)[(size_t)n * 2 + r] = s;
        } else if (gt == G_LOGITS) {
            const int ch = n / p.Vpad, vv = n - ch * p.Vpad;
            if (ch < p.C && vv < p.V) p.logits[((size_t)r * p.C + ch) * p.V + vv] = s;
        } else {   // residual add (dia/layers.py:555,574,582)
            float* xp = reinterpret_cast<float*>(p.x) + (size_t)n * 2 + r;
            *xp = ldcg_f(xp) + s;
        }
    }
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
    }
}

// ------------------------------------------------------------------------------------------
// embedding gather-sum (dia/layers.py:691-696): x = ((e0 + e1) + e2) ... + e8, both CFG rows
// ------------------------------------------------------------------------------------------
__device__ void embed_stage(Ctx& c, int pos) {
    const StepParams& p = *c.p;
    const int d = blockIdx.x * kConsumerThreads + c.tid;
    if (d >= p.D) return;
    const int* t0;
    const int* t1;
    if (p.tokens != nullptr) { t0 = p.tokens; t1 = p.tokens + p.C; }
    else { t0 = t1 = p.grid + (size_t)(pos - 1) * p.C; }
    float s0 = 0.f, s1 = 0.f;
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
            switch (kind) {
                case S_EMBED: embed_stage(c, pos); break;
                case S_SATTN: attn_stage<4, true>(c, layer, pos, slot); break;
                case S_CATTN: attn_stage<1, false>(c, layer, pos, slot); break;
                case S_SAMPLE: sample_stage(c, n, pos); break;
                default: gemm_stage(c, gemm_of_kind(kind), layer); break;
            }
            const bool last = (n == p.n_steps - 1) && (s == p.stage_end - 1);
            if (!last) grid_barrier(c);
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
