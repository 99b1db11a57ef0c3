// sampler.cuh - one-channel sampling by one CTA (dia/model.py:32-82), shared by the single-utterance step kernel,
// the batched step kernel and the standalone head kernel.
#pragma once

#include "common.cuh"

namespace dia {

// ------------------------------------------------------------------------------------------
// sampling: argmax / top-k / top-p / multinomial (dia/model.py:32-82) for ONE channel by one CTA
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t (&ctr)[4], uint32_t k0, uint32_t k1) {
#pragma unroll 1
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
        const uint32_t n0 = hi1 ^ ctr[1] ^ k0, n1 = lo1, n2 = hi0 ^ ctr[3] ^ k1, n3 = lo0;
        ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

constexpr int kMaxCand = 64;
constexpr int kPerThread = 5;        // vocab <= 1280 entries per channel (Dia: 1028)

// float -> unsigned key with the same order (larger float = larger key; -inf is the smallest real key)
__device__ __forceinline__ uint32_t order_key(float v) {
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct SampleSmem {
    uint32_t keys[kPerThread * kConsumerThreads];
    uint32_t hist[4 * 256];          // one histogram per radix pass
    float cv[kMaxCand];
    int ci[kMaxCand];
    float sv[kMaxCand];
    int si[kMaxCand];
    float wv[kConsumerWarps];
    int wi[kConsumerWarps];
    int sel[4];                      // bin, count above, token
    // full-vocabulary path (top-k disabled or > kMaxCand): every entry in sorted order
    float gv[kPerThread * kConsumerThreads];     // exp(logit - max), descending
    short gi[kPerThread * kConsumerThreads];     // token id at each sorted position
    float gstat[4];                              // kept count, normaliser of the kept entries
};

// the k-th largest of the keys (each thread holds kPerThread of them; key 0 = no entry) by an 8-bit radix select.
// One block barrier per pass: every warp scans the histogram itself (redundantly) and keeps the result in
// registers, so there is no hand-over through shared memory and no single-warp section.  sm->hist must be zero.
// n_gt: entries strictly above the returned key.  Fewer than kk entries: returns 0 (everything is kept).
__device__ __forceinline__ uint32_t radix_select_kth(const uint32_t (&key)[kPerThread], int kk, SampleSmem* sm, int lane,
                                                     int& n_gt) {
    uint32_t prefix = 0u, pmask = 0u;
    n_gt = 0;
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        uint32_t* hist = sm->hist + 256 * pass;
#pragma unroll
        for (int i = 0; i < kPerThread; ++i)
            if (key[i] != 0u && (key[i] & pmask) == prefix) atomicAdd(&hist[(key[i] >> shift) & 255u], 1u);
        consumer_sync();
        // lane holds bins lane*8 .. lane*8+7; suffix sums locate the bin where the count from the top reaches kk
        int tot = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) tot += (int)hist[lane * 8 + b];
        int suf = tot;
#pragma unroll
        for (int m = 1; m <= 16; m <<= 1) {
            const int t = __shfl_down_sync(0xffffffffu, suf, m);
            if (lane + m < 32) suf += t;
        }
        int above = suf - tot, bin = 0;
        const bool mine = above < kk && kk <= suf;
        if (mine) {
#pragma unroll 1
            for (int b = 7; b >= 0; --b) {
                const int cb = (int)hist[lane * 8 + b];
                if (above + cb >= kk) { bin = lane * 8 + b; break; }
                above += cb;
            }
        }
        const unsigned who = __ballot_sync(0xffffffffu, mine);
        const int srcl = who ? __ffs(who) - 1 : 0;
        bin = __shfl_sync(0xffffffffu, bin, srcl);
        above = __shfl_sync(0xffffffffu, above, srcl);
        if (!who) { bin = 0; above = 0; }                   // fewer than kk candidates left: keep everything
        prefix |= (uint32_t)bin << shift;
        pmask |= 0xffu << shift;
        n_gt += above;
        kk -= above;
    }
    return prefix;
}

// The reference's optional top-k (dia/model.py:43-50: `cfg_filter_top_k` None or 0 skips the filter) and top-k wider
// than the fused sampler's candidate list: top-p over ALL entries.  Not the default path, so simple rather than fast:
// a rank sort of the whole channel (value descending, lower index first on ties), softmax, the sequential cumulative
// sum of dia/model.py:56-70 and the inverse-CDF draw by one thread.  sm->keys holds the keys, sm->hist is zero.
static __device__ __noinline__ int sample_full_vocab_cta(const uint32_t (&key)[kPerThread], int V, float top_p, int top_k,
                                                  unsigned long long seed, unsigned long long draw, int ch,
                                                  float* probs_out, SampleSmem* sm, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    uint32_t thr = 0u;
    if (top_k > 0 && top_k < V) { int n_gt; thr = radix_select_kth(key, top_k, sm, lane, n_gt); }
    // ---- rank of every entry among all of them ----------------------------------------------------------------
    int rank[kPerThread];
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) rank[i] = 0;
#pragma unroll 1
    for (int j = 0; j < V; ++j) {
        const uint32_t kj = sm->keys[j];
#pragma unroll
        for (int i = 0; i < kPerThread; ++i)
            rank[i] += (kj > key[i] || (kj == key[i] && j < tid + kConsumerThreads * i)) ? 1 : 0;
    }
    // the maximum (rank 0) first: everything else is relative to it
#pragma unroll
    for (int i = 0; i < kPerThread; ++i)
        if (tid + kConsumerThreads * i < V && rank[i] == 0) sm->gstat[2] = key_value(key[i]);
    consumer_sync();
    const float mx = sm->gstat[2];
    float zsum = 0.f;
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
        const int idx = tid + kConsumerThreads * i;
        if (idx < V) {
            // entries below the k-th value are masked to -inf (dia/model.py:49-50): probability exactly 0
            const float e = key[i] >= thr ? expf(key_value(key[i]) - mx) : 0.f;
            sm->gv[rank[i]] = e;
            sm->gi[rank[i]] = (short)idx;
            zsum += e;
        }
    }
    zsum = warp_sum(zsum);
    if (lane == 0) sm->wv[warp] = zsum;
    consumer_sync();
    if (tid == 0) {
        float Z = 0.f;
#pragma unroll 1
        for (int ww = 0; ww < kConsumerWarps; ++ww) Z += sm->wv[ww];
        // top-p: entry i of the sorted list is removed iff the cumulative probability BEFORE it already exceeds top_p
        int nkeep = V;
        if (top_p < 1.0f) {
            double cum = 0.0;                              // torch.cumsum on the CPU accumulates float rows in double
            nkeep = 0;
#pragma unroll 1
            for (int i = 0; i < V; ++i) {
                if (i > 0 && (float)cum > top_p) break;
                cum += (double)(sm->gv[i] / Z);
                nkeep = i + 1;
            }
        }
        float Z2 = 0.f;
#pragma unroll 1
        for (int i = 0; i < nkeep; ++i) Z2 += sm->gv[i];
        uint32_t ctr[4] = {(uint32_t)draw, (uint32_t)(draw >> 32), (uint32_t)ch, 0x44494131u};
        philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
        const float target = (float)(ctr[0] >> 8) * (1.0f / 16777216.0f) * Z2;
        float cum = 0.f;
        int hit = nkeep - 1;
#pragma unroll 1
        for (int i = 0; i < nkeep; ++i) {
            cum += sm->gv[i];
            if (cum > target) { hit = i; break; }
        }
        while (hit > 0 && sm->gv[hit] == 0.f) --hit;       // never a masked entry
        sm->sel[2] = sm->gi[hit];
        sm->gstat[0] = (float)nkeep;
        sm->gstat[1] = Z2;
    }
    consumer_sync();
    if (probs_out != nullptr) {
        const int nkeep = (int)sm->gstat[0];
        const float Z2 = sm->gstat[1];
#pragma unroll 1
        for (int i = tid; i < V; i += kConsumerThreads) probs_out[sm->gi[i]] = i < nkeep ? sm->gv[i] / Z2 : 0.f;
    }
    const int tok = sm->sel[2];
    consumer_sync();
    return tok;
}

// g[i] = guided + masked logit of entry tid + 256 i (-inf past V).  All 256 threads call; returns the token id
// (valid in every thread).  probs_out (optional): the filtered distribution the draw is made from.
static __device__ int sample_channel_cta(const float (&g)[kPerThread], int V, float temperature, float top_p, int top_k,
                                  unsigned long long seed, unsigned long long draw, int ch, float* probs_out,
                                  SampleSmem* sm, int tid, long long* ts = nullptr) {
    const int warp = tid >> 5, lane = tid & 31;
    if (temperature == 0.0f) {          // torch.argmax: first maximal index
        float bv = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            const int idx = tid + kConsumerThreads * i;
            if (idx < V && (g[i] > bv || bi == 0x7fffffff)) { bv = g[i]; bi = idx; }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, m);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) { sm->wv[warp] = bv; sm->wi[warp] = bi; }
        consumer_sync();
        bv = sm->wv[0]; bi = sm->wi[0];
#pragma unroll 1
        for (int ww = 1; ww < kConsumerWarps; ++ww) {
            const float ov = sm->wv[ww];
            const int oi = sm->wi[ww];
            if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        consumer_sync();
        return bi;
    }
    // ---- logits / temperature as order-preserving keys (dia/model.py:43) -----------------------------------
    uint32_t key[kPerThread];
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
        const int idx = tid + kConsumerThreads * i;
        key[i] = idx < V ? order_key(g[i] / temperature) : 0u;
        sm->keys[idx] = key[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) sm->hist[tid + kConsumerThreads * i] = 0u;   // one histogram per radix pass
    if (tid == 0) sm->sel[3] = 0;                                            // survivor counter
    consumer_sync();
    // ---- top-k threshold (dia/model.py:46-52): the k-th largest key, by an 8-bit radix select -------------
    if (top_k <= 0 || top_k > kMaxCand)
        return sample_full_vocab_cta(key, V, top_p, top_k, seed, draw, ch, probs_out, sm, tid);
    int n_gt = 0;
    const uint32_t prefix = radix_select_kth(key, top_k, sm, lane, n_gt);
    const uint32_t thr = prefix;
    if (ts && tid == 0) ts[2] = clock64();
    // ---- survivors: everything above the k-th value plus its ties (<= 64 kept) ---------------------------------
    // Every thread pushes its own candidates (unordered); the rank sort below orders them by (value desc, index asc).
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
        if (key[i] >= thr && key[i] != 0u) {
            const int posn = atomicAdd(&sm->sel[3], 1);
            if (posn < kMaxCand) { sm->cv[posn] = key_value(key[i]); sm->ci[posn] = tid + kConsumerThreads * i; }
        }
    }
    consumer_sync();
    const int n_all = sm->sel[3];
    if (n_all > kMaxCand) {
        // more than 64 candidates can only be ties at the threshold: keep the lowest indices among them (ordered scan)
        if (warp == 0) {
            int n_out = 0, ties = 0;
            const unsigned lt = (1u << lane) - 1u;
#pragma unroll 1
            for (int base = 0; base < V; base += 32) {
                const int idx = base + lane;
                const uint32_t kx = idx < V ? sm->keys[idx] : 0u;
                const bool gtk = kx > thr, eqk = kx == thr && kx != 0u;
                const unsigned m_eq = __ballot_sync(0xffffffffu, eqk);
                const bool keep = gtk || (eqk && (n_gt + ties + __popc(m_eq & lt)) < kMaxCand);
                const unsigned m_keep = __ballot_sync(0xffffffffu, keep);
                const int posn = n_out + __popc(m_keep & lt);
                if (keep && posn < kMaxCand) { sm->cv[posn] = key_value(kx); sm->ci[posn] = idx; }
                n_out += __popc(m_keep);
                ties += __popc(m_eq);
            }
        }
        consumer_sync();
    }
    const int ncand = min(n_all, kMaxCand);
    if (ts && tid == 0) ts[6] = clock64();
    // ---- sort descending by value (ties: lower index first): rank sort, 4 threads per candidate -----------------
    {
        const int e = tid >> 2, q4 = tid & 3;
        const float v = e < ncand ? sm->cv[e] : 0.f;
        const int id = e < ncand ? sm->ci[e] : 0;
        int rank = 0;
#pragma unroll 1
        for (int j = q4; j < ncand; j += 4) {
            const float wv = sm->cv[j];
            rank += (wv > v || (wv == v && sm->ci[j] < id)) ? 1 : 0;
        }
        rank += __shfl_xor_sync(0xffffffffu, rank, 1);
        rank += __shfl_xor_sync(0xffffffffu, rank, 2);
        if (e < ncand && q4 == 0) { sm->sv[rank] = v; sm->si[rank] = id; }
    }
    consumer_sync();
    if (warp == 0) {
        if (ts && tid == 0) ts[7] = clock64();
        // ---- softmax over the survivors, top-p on the sorted cumulative sum (dia/model.py:56-70) ------------
        const float mx = sm->sv[0];
        float e0 = lane < ncand ? expf(sm->sv[lane] - mx) : 0.f;
        float e1 = lane + 32 < ncand ? expf(sm->sv[lane + 32] - mx) : 0.f;
        const float Z = warp_sum(e0 + e1);
        if (lane < ncand) { sm->cv[lane] = e0; sm->sv[lane] = e0 / Z; }                  // cv: exp(l - max), sv: probability
        if (lane + 32 < ncand) { sm->cv[lane + 32] = e1; sm->sv[lane + 32] = e1 / Z; }   // (both in sorted order)
        __syncwarp();
        int nkeep = ncand;
        if (top_p < 1.0f) {
            // sequential cumulative sum (torch.cumsum order); every lane runs it on register-resident chunks
            float cum = 0.f;
            bool open = true;
            nkeep = 0;
#pragma unroll 1
            for (int i0 = 0; i0 < ncand && open; i0 += 16) {
                float pv[16];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<float4*>(pv + 4 * j) = *reinterpret_cast<const float4*>(sm->sv + i0 + 4 * j);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    // entry i is removed iff the cumulative probability BEFORE it already exceeds top_p
                    open = open && (i0 + j < ncand) && !(i0 + j > 0 && cum > top_p);
                    if (open) { cum += pv[j]; nkeep = i0 + j + 1; }
                }
            }
        }
        e0 = lane < nkeep ? e0 : 0.f;
        e1 = lane + 32 < nkeep ? e1 : 0.f;
        const float Z2 = warp_sum(e0 + e1);
        if (probs_out != nullptr) {
#pragma unroll 1
            for (int i = lane; i < V; i += 32) probs_out[i] = 0.f;
            __syncwarp();
            if (lane < nkeep) probs_out[sm->si[lane]] = e0 / Z2;
            if (lane + 32 < nkeep) probs_out[sm->si[lane + 32]] = e1 / Z2;
        }
        if (ts && tid == 0) ts[8] = clock64();
        // ---- multinomial(1): inverse CDF over the survivors with a Philox uniform ---------------------------
        {
            uint32_t ctr[4] = {(uint32_t)draw, (uint32_t)(draw >> 32), (uint32_t)ch, 0x44494131u};
            philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
            const float u = (float)(ctr[0] >> 8) * (1.0f / 16777216.0f);   // [0, 1)
            const float target = u * Z2;
            float cum = 0.f;
            int hit = nkeep - 1;
            bool open = true;
#pragma unroll 1
            for (int i0 = 0; i0 < nkeep && open; i0 += 16) {
                float pv[16];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<float4*>(pv + 4 * j) = *reinterpret_cast<const float4*>(sm->cv + i0 + 4 * j);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (open && i0 + j < nkeep) {
                        cum += pv[j];
                        if (cum > target) { hit = i0 + j; open = false; }
                    }
                }
            }
            if (lane == 0) sm->sel[2] = sm->si[hit];
        }
    }
    consumer_sync();
    const int tok = sm->sel[2];
    consumer_sync();
    return tok;
}

}  // namespace dia
