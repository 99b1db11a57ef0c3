// engine.cu - host side of libdia_b200.so: the C ABI declared in include/dia_b200.h.
//
// Owns the repacked weight stream, the per-CTA tables, the scratch vectors and the
// device-side generate state; builds StepParams and launches the persistent step kernel.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "engine_internal.h"

using namespace dia;

namespace {

constexpr int kTimingSteps = 16;
constexpr int kErrWords = 16 + 2 * 12 * 160;        // watchdog record + one (site, info) pair per warp of every CTA
thread_local std::string g_last_cuda_error;
std::atomic<long long> g_launches{0};

#define CK(expr)                                                                           \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
 if (_e != cudaSuccess) {                                                           \
            g_last_cuda_error = std::string(#expr) + ": " + cudaGetErrorString(_e);        \
            (void)cudaGetLastError(); /* non-sticky errors must not poison the next launch */ \
            return DIA_B200_ECUDA;                                                         \
        }                                                                                  \
    } while (0)

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// every ABI call runs on the engine's device and leaves the caller's current device as it found it
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) err = cudaSetDevice(device);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(dev) DeviceGuard _guard(dev); CK(_guard.err)

}  // namespace

struct dia_b200_engine {
    dia_b200_shape shape{};
    int device = 0;
    int G = 0;
    int Vpad = 0;
    int n_groups[G_COUNT]{};
    int Kdim[G_COUNT]{};
    int tclass[G_COUNT]{};
    std::vector<CtaTable> tab;
    std::vector<int> owner[G_COUNT], local[G_COUNT];
    size_t stream_bytes = 0;          // allocation size (with alignment padding)
    long long weight_bytes = 0;       // payload bytes = bytes streamed per step

    // device
    unsigned char* d_wstream = nullptr;
    CtaTable* d_tab = nullptr;
    int* d_owner[G_COUNT]{};
    int* d_local[G_COUNT]{};
    float* d_emb = nullptr;
    float* d_norms = nullptr;
    float* d_rope_sin = nullptr;
    float* d_rope_cos = nullptr;
    int n_pos = 0;
    float** d_ptrs = nullptr;         // [4][L] self_k, self_v, cross_k, cross_v
    float2* d_x = nullptr;
    float* d_logits = nullptr;
    unsigned long long* d_ll = nullptr;   // the flag-in-data region (zeroed before every launch)
    size_t ll_bytes = 0;
    int* d_err = nullptr;             // device alias of h_err (mapped pinned memory: survives a trapped kernel)
    int* d_pred = nullptr;
    int* d_tokens = nullptr;          // staging [2][C]
    GenState* d_gs = nullptr;
    long long* d_timing = nullptr;    // [kTimingSteps][S][8], debug only
    unsigned long long* d_cta_timing = nullptr;   // [S][G], debug only
    bool timing_on = false;
    int timing_cta = 0;
    // pinned host staging
    float** h_ptrs = nullptr;
    GenState* h_gs = nullptr;
    int* h_err = nullptr;

    bool weights_loaded = false, caches_bound = false, rope_set = false, gen_active = false;
    int text_len = 0;
    int sa_nsplit = 1, ca_nsplit = 1;
    int n_res = 0;                    // CTAs [0, n_res) own residual-stream columns
    // generate loop
    dia_b200_gen_params gp{};
    int* gen_grid = nullptr;
    int gen_pos = 0, gen_slot = 0;
    long long gen_steps = 0;          // decode steps launched since generate_begin (RNG draw index)

    int Kfull[G_COUNT]{};             // contraction lengths before K-row compaction
    int* d_rowmap[G_COUNT]{};         // device maps [L][Kfull] (logits [Kfull]) of the compacted GEMM families
    bool rowmap_set[G_COUNT]{};

    // batched engine (max_utts > 0): N utterances per launch on the tcgen05 step kernel (batch_kernel.cu)
    int max_utts = 0;
    int batch_mc = 0;                  // launch as CTA pairs that share their activation stages (every CTA has columns in every GEMM)
    unsigned char* d_bll = nullptr;   // its exchange region
    size_t bll_bytes = 0;
    bool utt_bound[kMaxUtt] = {};
    int utt_text_len[kMaxUtt] = {};
    int n_active = 0;                 // utterances of the running generate loop
    dia_b200_gen_params bgp[kMaxUtt] = {};
    int* bgrid[kMaxUtt] = {};
    int bpos[kMaxUtt] = {}, bslot[kMaxUtt] = {};
};

namespace {

int validate_shape(const dia_b200_shape& s) {
    if (s.n_layer <= 0 || s.d_model <= 0 || s.n_hidden <= 0 || s.q_heads <= 0 || s.kv_heads <= 0 ||
        s.cross_heads <= 0 || s.channels <= 0 || s.channels > DIA_B200_MAX_CHANNELS || s.vocab <= 0 ||
        s.max_audio_len <= 0 || s.max_text_len <= 0)
        return DIA_B200_EINVAL;
    if (s.q_heads != 4 * s.kv_heads) return DIA_B200_EUNSUPPORTED;       // kernels are built for GQA 4:1
    // every contraction length splits into 8 warp slices of whole 256-row fetch units (or one shorter unit)
    const int ks[4] = {s.d_model, s.n_hidden, s.q_heads * kHeadDim, s.cross_heads * kHeadDim};
    for (int k : ks) {
        const int sl = k / 8;
        if (k % 512 || (sl > 256 && sl % 256)) return DIA_B200_EUNSUPPORTED;
    }
    if (s.vocab > 5 * kConsumerThreads) return DIA_B200_EUNSUPPORTED;     // sampler: <= 5 entries of a channel per thread
    {
        const int full[7] = {s.d_model, s.q_heads * kHeadDim, s.d_model, s.cross_heads * kHeadDim, s.d_model, s.n_hidden, s.d_model};
        for (int t = 0; t < 7; ++t) {
            const int k = s.k_rows[t];
            if (k == 0) continue;
            if (k < 0 || k > full[t] || t == 5) return DIA_B200_EINVAL;                 // mlp-out is compacted through n_hidden
            if (k % 512 || (k / 8 > 256 && (k / 8) % 256)) return DIA_B200_EUNSUPPORTED;
            if (s.sparse24) return DIA_B200_EUNSUPPORTED;
        }
    }
    return DIA_B200_OK;
}

void fill_params(const dia_b200_engine* e, StepParams& p) {
    std::memset(&p, 0, sizeof(p));
    const dia_b200_shape& s = e->shape;
    p.L = s.n_layer; p.D = s.d_model; p.F = s.n_hidden; p.Hq = s.q_heads; p.Hkv = s.kv_heads; p.Hc = s.cross_heads;
    p.C = s.channels; p.V = s.vocab; p.Vpad = e->Vpad; p.Lmax = s.max_audio_len; p.Smax = s.max_text_len;
    for (int i = 0; i < G_COUNT; ++i) { p.Kdim[i] = e->Kdim[i]; p.tclass[i] = e->tclass[i]; }
    p.eps = s.norm_eps; p.G = e->G; p.n_res = e->n_res; p.sa_nsplit = e->sa_nsplit; p.ca_nsplit = e->ca_nsplit;
    p.sparse24 = s.sparse24 ? 1 : 0;
    for (int i = 0; i < G_COUNT; ++i) { p.rowmap[i] = e->d_rowmap[i]; p.Kfull[i] = e->Kfull[i]; }
    p.wstream = e->d_wstream; p.cta_tab = e->d_tab; p.emb = e->d_emb; p.norms = e->d_norms;
    p.rope_sin = e->d_rope_sin; p.rope_cos = e->d_rope_cos; p.n_pos = e->n_pos;
    p.self_k = e->d_ptrs; p.self_v = e->d_ptrs + s.n_layer;
    p.cross_k = const_cast<const float* const*>(e->d_ptrs + 2 * s.n_layer);
    p.cross_v = const_cast<const float* const*>(e->d_ptrs + 3 * s.n_layer);
    p.text_len = e->text_len;
    p.x = e->d_x;
    p.logits = e->d_logits;
    ll_layout(p, &p, e->d_ll);
    p.err = e->d_err;
    p.n_steps = 1;
    p.cfg_scale = 3.0f; p.temperature = 0.0f; p.top_p = 0.95f; p.top_k = 35; p.max_tokens = s.max_audio_len;
    p.eos = s.eos_value; p.pad = s.pad_value; p.bos = s.bos_value;
    for (int i = 0; i < DIA_B200_MAX_CHANNELS; ++i) p.delay[i] = s.delay_pattern[i];
    p.pred_out = e->d_pred;
    p.timing = e->timing_on ? e->d_timing : nullptr;
    p.timing_cta = e->timing_cta;
    p.cta_timing = e->timing_on ? e->d_cta_timing : nullptr;
}

// stage ranges a launch may cover: it starts at the embedding, a layer or the logits head (the residual stream is
// handed over in d_x) and ends after the embedding, a layer, the logits head or the sampler
bool stage_range_ok(const dia_b200_engine* e, int b, int en) {
    const int L = e->shape.n_layer;
    auto boundary = [&](int s) { return s == 0 || s == 8 * L + 2 || s == 8 * L + 3 || (s >= 1 && s <= 8 * L + 1 && (s - 1) % 8 == 0); };
    return b >= 0 && b < en && en <= 8 * L + 3 && boundary(b) && b != 8 * L + 3 && b != 8 * L + 2 && boundary(en);
}

int run_stages(dia_b200_engine* e, StepParams& p, bool cooperative, cudaStream_t st) {
    if (!stage_range_ok(e, p.stage_begin, p.stage_end)) return DIA_B200_EINVAL;
    if (p.n_steps > kTimingSteps) { p.timing = nullptr; p.cta_timing = nullptr; }
    // sequence flags restart at 1 in every launch: 16-bit flags bound the stages of one launch
    if ((long long)p.n_steps * (8 * p.L + 3) + 1 > 65535) return DIA_B200_EINVAL;
    CK(cudaMemsetAsync(e->d_ll, 0, e->ll_bytes, st));
    CK(launch_step_kernel(p, cooperative, st));
    g_launches++;
    return DIA_B200_OK;
}

int check_sampling_supported(float temperature, int top_k) {
    if (temperature == 0.0f) return DIA_B200_OK;
    if (temperature < 0.0f) return DIA_B200_EINVAL;
    (void)top_k;      // <= 0: no top-k filter (dia/model.py:43-50); 1..64: fused candidate list; wider: full-vocabulary path
    return DIA_B200_OK;
}

}  // namespace

extern "C" {

int dia_b200_abi_version(void) { return DIA_B200_ABI_VERSION; }

const char* dia_b200_error_string(int code) {
    switch (code) {
        case DIA_B200_OK: return "ok";
        case DIA_B200_EINVAL: return "invalid argument or unsupported shape";
        case DIA_B200_ECUDA: return "CUDA runtime error";
        case DIA_B200_ENOMEM: return "out of memory";
        case DIA_B200_ESTATE: return "call sequence error (weights / rope table / caches not set)";
        case DIA_B200_EUNSUPPORTED: return "not supported by the sm_100a decode path";
        default: return "unknown error";
    }
}

const char* dia_b200_last_cuda_error(void) { return g_last_cuda_error.c_str(); }
int64_t dia_b200_launch_count(void) { return g_launches.load(); }

static int create_engine(const dia_b200_shape* shape, int device, int n_ctas, int max_utts, dia_b200_engine** out);

int dia_b200_engine_create(const dia_b200_shape* shape, int device, int n_ctas, dia_b200_engine** out) {
    return create_engine(shape, device, n_ctas, 0, out);
}

int dia_b200_engine_create_batched(const dia_b200_shape* shape, int device, int n_ctas, int max_utterances,
                                   dia_b200_engine** out) {
    if (max_utterances < 1 || max_utterances > kMaxUtt) return DIA_B200_EINVAL;
    if (shape && shape->sparse24) return DIA_B200_EUNSUPPORTED;          // 2:4 slabs exist for the single-utterance kernel only
    if (shape) for (int t = 0; t < 7; ++t) if (shape->k_rows[t]) return DIA_B200_EUNSUPPORTED;   // so does K-row compaction
    return create_engine(shape, device, n_ctas, max_utterances, out);
}

int dia_b200_engine_max_utterances(const dia_b200_engine* e) { return e ? e->max_utts : DIA_B200_EINVAL; }

static int create_engine(const dia_b200_shape* shape, int device, int n_ctas, int max_utts, dia_b200_engine** out) {
    if (!shape || !out) return DIA_B200_EINVAL;
    int rc = validate_shape(*shape);
    if (rc) return rc;
    ON_DEVICE(device);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { g_last_cuda_error = "device is not sm_100-class"; return DIA_B200_EUNSUPPORTED; }
    int coop = 0;
    CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    if (!coop) { g_last_cuda_error = "cooperative launch unsupported"; return DIA_B200_EUNSUPPORTED; }

    dia_b200_engine* e = new (std::nothrow) dia_b200_engine();
    if (!e) return DIA_B200_ENOMEM;
    e->shape = *shape;
    e->device = device;
    e->max_utts = max_utts;
    const bool batch = max_utts > 0;
    e->G = n_ctas > 0 ? n_ctas : prop.multiProcessorCount;
    if (e->G > prop.multiProcessorCount) { delete e; return DIA_B200_EINVAL; }   // 1 CTA / SM must be co-resident
    const dia_b200_shape& s = e->shape;
    const int G = e->G;
    const int sp = s.sparse24 ? 1 : 0;
    if (G < 2 * s.kv_heads || G < s.cross_heads) { delete e; return DIA_B200_EINVAL; }
    e->Vpad = (s.vocab + 7) & ~7;
    e->sa_nsplit = G / (2 * s.kv_heads);
    e->ca_nsplit = G / s.cross_heads;
    if (batch) {
        // (row, kv head) pairs x key splits / (utterance, cross head) pairs x key splits over the CTAs
        if (G < 2 * max_utts * s.kv_heads || G < max_utts * s.cross_heads) { delete e; return DIA_B200_EINVAL; }
        e->sa_nsplit = G / (2 * max_utts * s.kv_heads);
        e->ca_nsplit = G / (max_utts * s.cross_heads);
    } else {
        int per = (s.max_audio_len + e->sa_nsplit - 1) / e->sa_nsplit;
        int perc = (s.max_text_len + e->ca_nsplit - 1) / e->ca_nsplit;
        if (((per + 15) & ~15) > 1024 || ((perc + 15) & ~15) > 1024) { delete e; return DIA_B200_EUNSUPPORTED; }
        if ((s.max_text_len + 127) / 128 > e->ca_nsplit) { delete e; return DIA_B200_EUNSUPPORTED; }
    }

    // ---- column-group partition of every GEMM over the CTAs --------------------------------
    // The three residual GEMMs (self-o, cross-o, mlp-out) share ONE partition, so each element of the residual
    // stream has a fixed owner thread and never leaves its register; the other GEMMs are balanced against the
    // bytes a CTA already streams per layer.
    const int nq = s.q_heads * kHeadDim, nkv = s.kv_heads * kHeadDim, nc = s.cross_heads * kHeadDim;
    // mlp-in: a group holds the gate AND the up columns of 4 hidden units, so any slab width keeps every pair in one CTA
    const int units[G_COUNT] = {(nq + 2 * nkv) / 8, s.d_model / 8, nc / 8, s.d_model / 8, s.n_hidden / 4,
                                s.d_model / 8, s.channels * e->Vpad / 8};
    const int mult[G_COUNT] = {1, 1, 1, 1, 1, 1, 1};
    const int kd[G_COUNT] = {s.d_model, nq, s.d_model, nc, s.d_model, s.n_hidden, s.d_model};
    if (s.d_model / 8 > 2 * G) { delete e; return DIA_B200_EUNSUPPORTED; }   // residual columns: one MMA tile per CTA
    e->tab.assign(G, CtaTable{});
    std::vector<long long> load(G, 0);                   // bytes per layer assigned so far (for balancing)
    const int order_t[G_COUNT] = {G_SO, G_CO, G_WO, G_WI, G_QKV, G_CQ, G_LOGITS};
    for (int oi = 0; oi < G_COUNT; ++oi) {
        const int t = order_t[oi];
        e->Kfull[t] = kd[t];
        e->Kdim[t] = s.k_rows[t] > 0 ? s.k_rows[t] : kd[t];
        e->n_groups[t] = units[t] * mult[t];
        std::vector<int> cnt(G, units[t] / G);
        if (t == G_CO || t == G_WO) {
            for (int c = 0; c < G; ++c) cnt[c] = e->tab[c].gc[G_SO];
        } else {
            // the CTAs with the least bytes so far take the `extra` units
            std::vector<int> order(G);
            for (int c = 0; c < G; ++c) order[c] = c;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return load[a] < load[b]; });
            for (int i = 0; i < units[t] % G; ++i) cnt[order[i]]++;
            // A slab row is gc * 16 bytes; ldmatrix reads 8 consecutive rows, which collide 4-way (8-way) in the
            // shared-memory banks when gc is a multiple of 4 (8) - an 8-group logits slab runs at half the HBM rate.
            // Trade such CTAs pairwise to one unit more / one unit less; a leftover one trades with a neighbour
            // that stays conflict-free.
            {
                auto is_bad = [&](int n) { const int gc = n * mult[t]; return gc >= 4 && gc % 4 == 0; };
                std::vector<int> bad;
                for (int c = 0; c < G; ++c) if (is_bad(cnt[c])) bad.push_back(c);
                for (size_t i = 0; i + 1 < bad.size(); i += 2) { cnt[bad[i]]++; cnt[bad[i + 1]]--; }
                if (bad.size() % 2) {
                    for (int c = 0; c < G; ++c) {
                        if (c != bad.back() && !is_bad(cnt[c]) && cnt[c] >= 2 && !is_bad(cnt[c] - 1) &&
                            cnt[c] <= cnt[bad.back()]) {
                            cnt[bad.back()]++; cnt[c]--;
                            break;
                        }
                    }
                }
            }
        }
        int g0 = 0;
        e->owner[t].resize(e->n_groups[t]);
        e->local[t].resize(e->n_groups[t]);
        for (int c = 0; c < G; ++c) {
            const int gc = cnt[c] * mult[t];
            if (gc > 16) { delete e; return DIA_B200_EUNSUPPORTED; }          // <= 8 MMA tiles per CTA and GEMM
            e->tab[c].g0[t] = g0;
            e->tab[c].gc[t] = gc;
            for (int i = 0; i < gc; ++i) { e->owner[t][g0 + i] = c; e->local[t][g0 + i] = i; }
            g0 += gc;
            if (t != G_LOGITS) load[c] += (long long)(batch ? bslab_bytes(gc, e->Kdim[t]) : gemm_slab_bytes(gc, e->Kdim[t], sp));
        }
    }
    for (int t = 0; t < G_COUNT; ++t) {
        int mx = 1;
        for (int c = 0; c < G; ++c) mx = std::max(mx, (e->tab[c].gc[t] + 1) / 2);
        e->tclass[t] = mx <= 1 ? 1 : mx <= 2 ? 2 : mx <= 4 ? 4 : 8;
    }
    for (int t = 0; t < G_COUNT && !batch; ++t) {
        const int need = e->tclass[t] == 1 ? 64 : 32;       // k-blocks in flight per MMA group; slots start on even k-blocks
        for (int c = 0; c < G; ++c)
            if (e->tab[c].gc[t] > 0 && gemm_slot_rows(e->tab[c].gc[t], e->Kdim[t], sp) % need) { delete e; return DIA_B200_EUNSUPPORTED; }
    }
    if (batch && G % 2 == 0) {
        // CTA pairs walk the activation ring in lockstep only if both have columns in every GEMM stage
        e->batch_mc = 1;
        for (int c = 0; c < G; ++c)
            for (int t = 0; t < G_COUNT; ++t)
                if (e->tab[c].gc[t] == 0) e->batch_mc = 0;
        // measured (tools/batch_bench.py, A/B on one box): sharing the stages couples the pair's pipelines and is 7 % slower
        // than two independent CTAs since the four-issuer loop - off unless asked for
        const char* env = std::getenv("DIA_BATCH_MULTICAST");
        if (!(env && env[0] == '1')) e->batch_mc = 0;
    }
    for (int c = 0; c < G; ++c) {
        if (e->tab[c].gc[G_SO] > 0) {
            if (c != e->n_res) { delete e; return DIA_B200_EUNSUPPORTED; }   // owners must be CTAs 0 .. n_res-1
            e->n_res = c + 1;
        }
    }
    if (e->n_res > 160) { delete e; return DIA_B200_EUNSUPPORTED; }          // sum(x^2) partials: 5 per lane of one warp
    unsigned long long off = 0;
    for (int c = 0; c < G; ++c) {
        CtaTable& t = e->tab[c];
        unsigned o = 0;
        auto slab_bytes = [&](int g) { return batch ? bslab_bytes(t.gc[g], e->Kdim[g]) : gemm_slab_bytes(t.gc[g], e->Kdim[g], sp); };
        for (int g = 0; g < G_LOGITS; ++g) { t.slab_off[g] = o; o += (unsigned)slab_bytes(g); }
        t.layer_bytes = o;
        t.slab_off[G_LOGITS] = 0;
        t.logits_off = (unsigned long long)o * s.n_layer;
        const unsigned long long total = t.logits_off + slab_bytes(G_LOGITS);
        t.stream_base = off;
        e->weight_bytes += (long long)total;
        off += (total + 255ull) & ~255ull;
    }
    e->stream_bytes = off + 65536;       // slack: bulk copies never read past a slab, this is belt and braces

    // ---- device allocations ---------------------------------------------------------------------
    const size_t D = s.d_model;
#define ALLOC(ptr, bytes)                                                                         \
    do {                                                                                          \
        cudaError_t _e = cudaMalloc(reinterpret_cast<void**>(&(ptr)), (bytes));                   \
        if (_e != cudaSuccess) {                                                                  \
            g_last_cuda_error = std::string("cudaMalloc " #ptr ": ") + cudaGetErrorString(_e);    \
            dia_b200_engine_destroy(e);                                                           \
            return _e == cudaErrorMemoryAllocation ? DIA_B200_ENOMEM : DIA_B200_ECUDA;            \
        }                                                                                         \
        cudaMemset((ptr), 0, (bytes));                                                            \
    } while (0)
    ALLOC(e->d_wstream, e->stream_bytes);
    ALLOC(e->d_tab, sizeof(CtaTable) * G);
    for (int t = 0; t < G_COUNT; ++t) {
        ALLOC(e->d_owner[t], sizeof(int) * e->n_groups[t]);
        ALLOC(e->d_local[t], sizeof(int) * e->n_groups[t]);
    }
    for (int t = 0; t < G_COUNT; ++t)
        if (s.k_rows[t] > 0) ALLOC(e->d_rowmap[t], sizeof(int) * (size_t)(t == G_LOGITS ? 1 : s.n_layer) * e->Kfull[t]);
    ALLOC(e->d_emb, sizeof(float) * (size_t)s.channels * s.vocab * D);
    ALLOC(e->d_norms, sizeof(float) * ((size_t)s.n_layer * 3 + 1) * D);
    const int n_utt = batch ? kMaxUtt : 1;
    ALLOC(e->d_ptrs, sizeof(float*) * 4 * s.n_layer * n_utt);
    ALLOC(e->d_x, sizeof(float2) * D);
    ALLOC(e->d_logits, sizeof(float) * 2 * n_utt * s.channels * s.vocab);
    if (batch) {
        BatchParams geom;
        std::memset(&geom, 0, sizeof(geom));
        geom.D = s.d_model; geom.F = s.n_hidden; geom.Hq = s.q_heads; geom.Hkv = s.kv_heads; geom.Hc = s.cross_heads;
        geom.C = s.channels; geom.V = s.vocab; geom.G = G; geom.U = max_utts; geom.R = 2 * max_utts;
        e->bll_bytes = batch_ll_layout(geom, nullptr, nullptr);
        ALLOC(e->d_bll, e->bll_bytes);
    }
    {
        StepParams geom;
        std::memset(&geom, 0, sizeof(geom));
        geom.D = s.d_model; geom.F = s.n_hidden; geom.Hq = s.q_heads; geom.Hkv = s.kv_heads; geom.Hc = s.cross_heads;
        geom.C = s.channels; geom.V = s.vocab; geom.G = G; geom.sa_nsplit = e->sa_nsplit; geom.ca_nsplit = e->ca_nsplit;
        e->ll_bytes = ll_layout(geom, nullptr, nullptr);
    }
    ALLOC(e->d_ll, e->ll_bytes);
    ALLOC(e->d_pred, sizeof(int) * DIA_B200_MAX_CHANNELS * n_utt);
    ALLOC(e->d_tokens, sizeof(int) * 2 * DIA_B200_MAX_CHANNELS * n_utt);
    ALLOC(e->d_gs, sizeof(GenState) * n_utt);
    ALLOC(e->d_timing, sizeof(long long) * 16 * kTimingSteps * (8 * s.n_layer + 3));
    ALLOC(e->d_cta_timing, sizeof(unsigned long long) * G * (8 * s.n_layer + 3));
#undef ALLOC
    if (cudaMallocHost(reinterpret_cast<void**>(&e->h_ptrs), sizeof(float*) * 4 * s.n_layer * n_utt) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void**>(&e->h_gs), sizeof(GenState) * n_utt) != cudaSuccess ||
        cudaHostAlloc(reinterpret_cast<void**>(&e->h_err), sizeof(int) * kErrWords, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&e->d_err), e->h_err, 0) != cudaSuccess) {
        dia_b200_engine_destroy(e);
        return DIA_B200_ENOMEM;
    }
    std::memset(e->h_err, 0, sizeof(int) * kErrWords);
    CK(cudaMemcpy(e->d_tab, e->tab.data(), sizeof(CtaTable) * G, cudaMemcpyHostToDevice));
    for (int t = 0; t < G_COUNT; ++t) {
        CK(cudaMemcpy(e->d_owner[t], e->owner[t].data(), sizeof(int) * e->n_groups[t], cudaMemcpyHostToDevice));
        CK(cudaMemcpy(e->d_local[t], e->local[t].data(), sizeof(int) * e->n_groups[t], cudaMemcpyHostToDevice));
    }
    *out = e;
    return DIA_B200_OK;
}

int dia_b200_engine_destroy(dia_b200_engine* e) {
    if (!e) return DIA_B200_OK;
    DeviceGuard _guard(e->device);
    cudaDeviceSynchronize();
    void* dev[] = {e->d_wstream, e->d_tab, e->d_emb, e->d_norms, e->d_rope_sin, e->d_rope_cos, e->d_ptrs, e->d_x,
                   e->d_logits, e->d_ll, e->d_pred, e->d_tokens, e->d_gs, e->d_timing, e->d_cta_timing, e->d_bll};
    for (void* p : dev) if (p) cudaFree(p);
    for (int t = 0; t < G_COUNT; ++t) { if (e->d_owner[t]) cudaFree(e->d_owner[t]); if (e->d_local[t]) cudaFree(e->d_local[t]); if (e->d_rowmap[t]) cudaFree(e->d_rowmap[t]); }
    if (e->h_ptrs) cudaFreeHost(e->h_ptrs);
    if (e->h_gs) cudaFreeHost(e->h_gs);
    if (e->h_err) cudaFreeHost(e->h_err);
    delete e;
    return DIA_B200_OK;
}

int dia_b200_engine_num_ctas(const dia_b200_engine* e) { return e ? e->G : DIA_B200_EINVAL; }
int64_t dia_b200_engine_weight_stream_bytes(const dia_b200_engine* e) { return e ? e->weight_bytes : DIA_B200_EINVAL; }

int dia_b200_load_decoder_weights(dia_b200_engine* e, const void* const* tensors, int n_tensors, int dense_dtype,
                                  void* stream) {
    if (!e || !tensors || (dense_dtype != 0 && dense_dtype != 1)) return DIA_B200_EINVAL;
    const dia_b200_shape& s = e->shape;
    const int expect = s.channels + 11 * s.n_layer + 2;
    if (n_tensors != expect) return DIA_B200_EINVAL;
    for (int i = 0; i < n_tensors; ++i) if (!tensors[i]) return DIA_B200_EINVAL;
    ON_DEVICE(e->device);
    cudaStream_t st = S(stream);
    const size_t D = s.d_model;
    int ti = 0;
    for (int c = 0; c < s.channels; ++c, ++ti)
        CK(cudaMemcpyAsync(e->d_emb + (size_t)c * s.vocab * D, tensors[ti], sizeof(float) * s.vocab * D,
                           cudaMemcpyDeviceToDevice, st));
    RepackArgs a{};
    a.sparse = s.sparse24 ? 1 : 0;
    a.violations = e->d_pred;                              // scratch int (the prediction buffer is idle during a load)
    if (a.sparse) {
        CK(cudaMemsetAsync(e->d_wstream, 0, e->stream_bytes, st));     // metadata nibbles are OR-ed in
        CK(cudaMemsetAsync(e->d_pred, 0, sizeof(int), st));
    }
    a.batch_format = e->max_utts > 0 ? 1 : 0;
    a.src_bf16 = dense_dtype; a.Hq = s.q_heads; a.Hkv = s.kv_heads; a.F = s.n_hidden; a.V = s.vocab; a.Vpad = e->Vpad;
    a.C = s.channels; a.tab = e->d_tab; a.wstream = e->d_wstream;
    auto repack = [&](int gemm, int layer, const void* s0, const void* s1, const void* s2, int N) -> int {
        a.src[0] = s0; a.src[1] = s1; a.src[2] = s2; a.gemm = gemm; a.layer = layer; a.K = e->Kdim[gemm];
        a.n_groups = e->n_groups[gemm]; a.N = N; a.owner = e->d_owner[gemm]; a.local = e->d_local[gemm];
        CK(launch_repack(a, st));
        g_launches++;
        return DIA_B200_OK;
    };
    for (int l = 0; l < s.n_layer; ++l) {
        for (int n = 0; n < 3; ++n, ++ti)
            CK(cudaMemcpyAsync(e->d_norms + ((size_t)l * 3 + n) * D, tensors[ti], sizeof(float) * D,
                               cudaMemcpyDeviceToDevice, st));
        const void *q = tensors[ti], *k = tensors[ti + 1], *v = tensors[ti + 2], *o = tensors[ti + 3];
        const void *cq = tensors[ti + 4], *co = tensors[ti + 5], *wi = tensors[ti + 6], *wo = tensors[ti + 7];
        ti += 8;
        int rc;
        if ((rc = repack(G_QKV, l, q, k, v, 0))) return rc;
        if ((rc = repack(G_SO, l, o, nullptr, nullptr, s.d_model))) return rc;
        if ((rc = repack(G_CQ, l, cq, nullptr, nullptr, s.cross_heads * kHeadDim))) return rc;
        if ((rc = repack(G_CO, l, co, nullptr, nullptr, s.d_model))) return rc;
        if ((rc = repack(G_WI, l, wi, nullptr, nullptr, 0))) return rc;
        if ((rc = repack(G_WO, l, wo, nullptr, nullptr, s.d_model))) return rc;
    }
    CK(cudaMemcpyAsync(e->d_norms + (size_t)s.n_layer * 3 * D, tensors[ti], sizeof(float) * D, cudaMemcpyDeviceToDevice,
                       st));
    ++ti;
    int rc = repack(G_LOGITS, 0, tensors[ti], nullptr, nullptr, 0);
    if (rc) return rc;
    if (a.sparse) {
        // the compressed slabs are only valid for a model that really is 2:4 along K (the check is part of the repack)
        int bad = 0;
        CK(cudaMemcpyAsync(&bad, e->d_pred, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (bad != 0) { e->weights_loaded = false; return DIA_B200_EINVAL; }
    }
    e->weights_loaded = true;
    return DIA_B200_OK;
}

int dia_b200_set_row_map(dia_b200_engine* e, int gemm, const int32_t* map_host, void* stream) {
    if (!e || !map_host || gemm < 0 || gemm >= G_COUNT) return DIA_B200_EINVAL;
    if (!e->d_rowmap[gemm]) return DIA_B200_ESTATE;                      // this GEMM family is not compacted
    const int L = gemm == G_LOGITS ? 1 : e->shape.n_layer, Kf = e->Kfull[gemm], Kc = e->Kdim[gemm];
    std::vector<char> hit((size_t)Kc);
    for (int l = 0; l < L; ++l) {                                        // every compacted row exactly once per layer
        std::fill(hit.begin(), hit.end(), 0);
        for (int k = 0; k < Kf; ++k) {
            const int v = map_host[(size_t)l * Kf + k];
            if (v < -1 || v >= Kc) return DIA_B200_EINVAL;
            if (v >= 0) { if (hit[v]) return DIA_B200_EINVAL; hit[v] = 1; }
        }
        for (int v = 0; v < Kc; ++v) if (!hit[v]) return DIA_B200_EINVAL;
    }
    ON_DEVICE(e->device);
    CK(cudaMemcpyAsync(e->d_rowmap[gemm], map_host, sizeof(int) * (size_t)L * Kf, cudaMemcpyHostToDevice, S(stream)));
    CK(cudaStreamSynchronize(S(stream)));
    e->rowmap_set[gemm] = true;
    return DIA_B200_OK;
}

int dia_b200_set_rope_table(dia_b200_engine* e, const float* sin_host, const float* cos_host, int n_pos) {
    if (!e || !sin_host || !cos_host || n_pos <= 0) return DIA_B200_EINVAL;
    ON_DEVICE(e->device);
    if (e->d_rope_sin) { cudaFree(e->d_rope_sin); e->d_rope_sin = nullptr; }
    if (e->d_rope_cos) { cudaFree(e->d_rope_cos); e->d_rope_cos = nullptr; }
    const size_t bytes = sizeof(float) * (size_t)n_pos * (kHeadDim / 2);
    CK(cudaMalloc(reinterpret_cast<void**>(&e->d_rope_sin), bytes));
    CK(cudaMalloc(reinterpret_cast<void**>(&e->d_rope_cos), bytes));
    CK(cudaMemcpy(e->d_rope_sin, sin_host, bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_rope_cos, cos_host, bytes, cudaMemcpyHostToDevice));
    e->n_pos = n_pos;
    e->rope_set = true;
    return DIA_B200_OK;
}

int dia_b200_bind_caches(dia_b200_engine* e, void* const* self_k, void* const* self_v, const void* const* cross_k,
                         const void* const* cross_v, int n_layer, int text_len, void* stream) {
    if (!e || !self_k || !self_v || !cross_k || !cross_v) return DIA_B200_EINVAL;
    if (e->max_utts > 0) return DIA_B200_ESTATE;
    if (n_layer != e->shape.n_layer || text_len < 0 || text_len > e->shape.max_text_len) return DIA_B200_EINVAL;
    ON_DEVICE(e->device);
    const int L = n_layer;
    // the pinned staging buffer is reused: make sure the previous async copy has drained
    CK(cudaStreamSynchronize(S(stream)));
    for (int l = 0; l < L; ++l) {
        if (!self_k[l] || !self_v[l] || !cross_k[l] || !cross_v[l]) return DIA_B200_EINVAL;
        if ((reinterpret_cast<uintptr_t>(self_k[l]) | reinterpret_cast<uintptr_t>(self_v[l]) |
             reinterpret_cast<uintptr_t>(cross_k[l]) | reinterpret_cast<uintptr_t>(cross_v[l])) & 15)
            return DIA_B200_EINVAL;                      // bulk copies need 16-byte aligned rows
        e->h_ptrs[l] = static_cast<float*>(self_k[l]);
        e->h_ptrs[L + l] = static_cast<float*>(self_v[l]);
        e->h_ptrs[2 * L + l] = const_cast<float*>(static_cast<const float*>(cross_k[l]));
        e->h_ptrs[3 * L + l] = const_cast<float*>(static_cast<const float*>(cross_v[l]));
    }
    CK(cudaMemcpyAsync(e->d_ptrs, e->h_ptrs, sizeof(float*) * 4 * L, cudaMemcpyHostToDevice, S(stream)));
    e->text_len = text_len;
    e->caches_bound = true;
    e->gen_active = false;
    return DIA_B200_OK;
}

static int ready(const dia_b200_engine* e, bool need_caches) {
    if (!e) return DIA_B200_EINVAL;
    if (e->max_utts > 0) return DIA_B200_ESTATE;              // a batched engine: use the dia_b200_batch_* entry points
    if (!e->weights_loaded || !e->rope_set) return DIA_B200_ESTATE;
    for (int t = 0; t < G_COUNT; ++t) if (e->d_rowmap[t] && !e->rowmap_set[t]) return DIA_B200_ESTATE;
    if (need_caches && !e->caches_bound) return DIA_B200_ESTATE;
    return DIA_B200_OK;
}

int dia_b200_decode_step(dia_b200_engine* e, const int32_t* tokens, int pos, int slot, float* logits, void* stream) {
    int rc = ready(e, true);
    if (rc) return rc;
    if (!tokens || !logits || pos < 0 || slot < 0 || slot >= e->shape.max_audio_len) return DIA_B200_EINVAL;
    ON_DEVICE(e->device);
    StepParams p;
    fill_params(e, p);
    p.tokens = tokens;
    p.stage_begin = 0;
    p.stage_end = 8 * p.L + 2;          // embed .. logits (no sampling)
    p.pos0 = pos; p.slot0 = slot;
    p.want_x = 1;
    p.logits = logits;
    return run_stages(e, p, true, S(stream));
}

int dia_b200_decoder_layer_step(dia_b200_engine* e, int layer, const float* x_in, float* x_out, int pos, int slot,
                                void* stream) {
    int rc = ready(e, true);
    if (rc) return rc;
    if (!x_in || !x_out || layer < 0 || layer >= e->shape.n_layer || pos < 0 || slot < 0 ||
        slot >= e->shape.max_audio_len)
        return DIA_B200_EINVAL;
    ON_DEVICE(e->device);
    cudaStream_t st = S(stream);
    CK(launch_interleave(x_in, e->d_x, e->shape.d_model, st));
    g_launches++;
    StepParams p;
    fill_params(e, p);
    p.stage_begin = 1 + 8 * layer;
    p.stage_end = p.stage_begin + 8;
    p.pos0 = pos; p.slot0 = slot;
    p.want_x = 1;
    rc = run_stages(e, p, true, st);
    if (rc) return rc;
    CK(launch_deinterleave(e->d_x, x_out, e->shape.d_model, st));
    g_launches++;
    return DIA_B200_OK;
}

int dia_b200_embed_sum(dia_b200_engine* e, const int32_t* tokens, int n_rows, float* x, void* stream) {
    if (!e || n_rows < 0) return DIA_B200_EINVAL;
    if (n_rows == 0) return DIA_B200_OK;
    if (!tokens || !x) return DIA_B200_EINVAL;
    if (!e->weights_loaded) return DIA_B200_ESTATE;
    ON_DEVICE(e->device);
    CK(launch_embed_sum(e->d_emb, tokens, n_rows, e->shape.channels, e->shape.vocab, e->shape.d_model, x, S(stream)));
    g_launches++;
    return DIA_B200_OK;
}

int dia_b200_head_sample(dia_b200_engine* e, const float* logits, float cfg_scale, float temperature, float top_p,
                         int top_k, uint64_t seed, uint64_t draw, int32_t* pred, float* probs, void* stream) {
    if (!e || !logits || !pred) return DIA_B200_EINVAL;
    int rc = check_sampling_supported(temperature, top_k);
    if (rc) return rc;
    ON_DEVICE(e->device);
    StepParams p;
    fill_params(e, p);
    p.cfg_scale = cfg_scale; p.temperature = temperature; p.top_p = top_p; p.top_k = top_k; p.seed = seed;
    CK(launch_head_sample(p, logits, draw, pred, probs, S(stream)));
    g_launches++;
    return DIA_B200_OK;
}

int dia_b200_generate_begin(dia_b200_engine* e, int32_t* grid, const dia_b200_gen_params* gp, void* stream) {
    int rc = ready(e, true);
    if (rc) return rc;
    if (!grid || !gp) return DIA_B200_EINVAL;
    if (gp->prefill_step < 1 || gp->max_tokens < 1 || gp->max_tokens > e->shape.max_audio_len ||
        gp->first_slot < 0 || gp->first_slot >= e->shape.max_audio_len)
        return DIA_B200_EINVAL;
    rc = check_sampling_supported(gp->temperature, gp->top_k);
    if (rc) return rc;
    ON_DEVICE(e->device);
    CK(cudaStreamSynchronize(S(stream)));     // h_gs staging reuse
    int dmax = 0;
    for (int c = 0; c < e->shape.channels; ++c) dmax = std::max(dmax, (int)e->shape.delay_pattern[c]);
    std::memset(e->h_gs, 0, sizeof(GenState));
    e->h_gs->dec_step = gp->prefill_step - 1;          // dia/model.py:736
    e->h_gs->bos_countdown = dmax;                     // :739
    e->h_gs->eos_countdown = -1;                       // :741
    e->h_gs->finished = (e->h_gs->dec_step >= gp->max_tokens - 1) ? 1 : 0;
    CK(cudaMemcpyAsync(e->d_gs, e->h_gs, sizeof(GenState), cudaMemcpyHostToDevice, S(stream)));
    std::memset(e->h_err, 0, sizeof(int) * kErrWords);         // the stream is idle (synchronised above)
    e->gp = *gp;
    e->gen_grid = grid;
    e->gen_pos = gp->prefill_step;                     // first iteration: cur = dec_step + 1
    e->gen_slot = gp->first_slot;                      // KVCache.current_idx after the (optional) prefill
    e->gen_steps = 0;
    e->gen_active = true;
    return DIA_B200_OK;
}

int dia_b200_generate_steps(dia_b200_engine* e, int n_steps, void* stream) {
    int rc = ready(e, true);
    if (rc) return rc;
    if (!e->gen_active) return DIA_B200_ESTATE;
    if (n_steps < 0) return DIA_B200_EINVAL;
    // never run past the cache / grid: the device loop is finished by then anyway
    n_steps = std::min(n_steps, e->shape.max_audio_len - e->gen_slot);
    n_steps = std::min(n_steps, e->shape.max_audio_len - e->gen_pos);
    if (n_steps <= 0) return DIA_B200_OK;
    ON_DEVICE(e->device);
    const int max_per_launch = 65534 / (8 * e->shape.n_layer + 3);      // 16-bit sequence flags
    while (n_steps > 0) {
        const int n = std::min(n_steps, max_per_launch);
        StepParams p;
        fill_params(e, p);
        p.stage_begin = 0;
        p.stage_end = 8 * p.L + 3;
        p.n_steps = n;
        p.pos0 = e->gen_pos; p.slot0 = e->gen_slot;
        p.grid = e->gen_grid; p.gs = e->d_gs;
        p.cfg_scale = e->gp.cfg_scale; p.temperature = e->gp.temperature; p.top_p = e->gp.top_p; p.top_k = e->gp.top_k;
        p.max_tokens = e->gp.max_tokens; p.seed = e->gp.seed; p.draw0 = (unsigned long long)e->gen_steps;
        rc = run_stages(e, p, true, S(stream));
        if (rc) return rc;
        e->gen_pos += n;
        e->gen_slot += n;
        e->gen_steps += n;
        n_steps -= n;
    }
    return DIA_B200_OK;
}

int dia_b200_generate_status(dia_b200_engine* e, dia_b200_gen_status* out, void* stream) {
    if (!e || !out) return DIA_B200_EINVAL;
    ON_DEVICE(e->device);
    CK(cudaMemcpyAsync(e->h_gs, e->d_gs, sizeof(GenState), cudaMemcpyDeviceToHost, S(stream)));
    CK(cudaStreamSynchronize(S(stream)));
    out->dec_step = e->h_gs->dec_step;
    out->finished = e->h_gs->finished;
    out->eos_detected = e->h_gs->eos_detected;
    out->eos_countdown = e->h_gs->eos_countdown;
    out->bos_countdown = e->h_gs->bos_countdown;
    out->steps_run = e->h_gs->steps_run;
    out->device_error = e->h_err[0];
    out->reserved = 0;
    return DIA_B200_OK;
}

// ---- batched engine: N utterances per launch (batch_kernel.cu) ---------------------------------------------------------------
namespace {

int batch_ready(const dia_b200_engine* e, int n_utts) {
    if (!e) return DIA_B200_EINVAL;
    if (e->max_utts <= 0 || !e->weights_loaded || !e->rope_set) return DIA_B200_ESTATE;
    if (n_utts < 1 || n_utts > e->max_utts) return DIA_B200_EINVAL;
    for (int u = 0; u < n_utts; ++u) if (!e->utt_bound[u]) return DIA_B200_ESTATE;
    return DIA_B200_OK;
}

void fill_batch_params(const dia_b200_engine* e, BatchParams& p, int n_utts) {
    std::memset(&p, 0, sizeof(p));
    const dia_b200_shape& s = e->shape;
    p.L = s.n_layer; p.D = s.d_model; p.F = s.n_hidden; p.Hq = s.q_heads; p.Hkv = s.kv_heads; p.Hc = s.cross_heads;
    p.C = s.channels; p.V = s.vocab; p.Vpad = e->Vpad; p.Lmax = s.max_audio_len; p.Smax = s.max_text_len;
    for (int i = 0; i < G_COUNT; ++i) p.Kdim[i] = e->Kdim[i];
    p.eps = s.norm_eps; p.G = e->G; p.U = n_utts; p.R = 2 * n_utts;
    p.mc = e->batch_mc;
    // the splits are sized for the utterances of THIS launch (fewer utterances: more CTAs per pair)
    p.sa_nsplit = std::max(1, e->G / (2 * n_utts * s.kv_heads));
    p.ca_nsplit = std::max(1, e->G / (n_utts * s.cross_heads));
    p.wstream = e->d_wstream; p.cta_tab = e->d_tab; p.emb = e->d_emb; p.norms = e->d_norms;
    p.rope_sin = e->d_rope_sin; p.rope_cos = e->d_rope_cos; p.n_pos = e->n_pos;
    const int UL = kMaxUtt * s.n_layer;
    p.self_k = e->d_ptrs; p.self_v = e->d_ptrs + UL;
    p.cross_k = const_cast<const float* const*>(e->d_ptrs + 2 * UL);
    p.cross_v = const_cast<const float* const*>(e->d_ptrs + 3 * UL);
    for (int u = 0; u < n_utts; ++u) p.utt[u].text_len = e->utt_text_len[u];
    BatchParams geom = p;
    geom.U = e->max_utts; geom.R = 2 * e->max_utts;            // one carve-up for every launch width
    batch_ll_layout(geom, &p, e->d_bll);
    p.err = e->d_err;
    p.n_steps = 1;
    p.cfg_scale = 3.0f; p.temperature = 0.0f; p.top_p = 0.95f; p.top_k = 35; p.max_tokens = s.max_audio_len;
    p.eos = s.eos_value; p.pad = s.pad_value; p.bos = s.bos_value;
    for (int i = 0; i < DIA_B200_MAX_CHANNELS; ++i) p.delay[i] = s.delay_pattern[i];
    p.pred_out = e->d_pred;
    p.prof = e->timing_on ? reinterpret_cast<unsigned long long*>(e->d_timing) : nullptr;
}

int run_batch(dia_b200_engine* e, BatchParams& p, cudaStream_t st) {
    // activation flags are one generation bit and the 32-bit sequence flags restart at 1: the region is zeroed per launch
    CK(cudaMemsetAsync(e->d_bll, 0, e->bll_bytes, st));
    CK(launch_batch_kernel(p, st));
    g_launches++;
    return DIA_B200_OK;
}

}  // namespace

int dia_b200_batch_bind_caches(dia_b200_engine* e, int utterance, void* const* self_k, void* const* self_v,
                               const void* const* cross_k, const void* const* cross_v, int n_layer, int text_len,
                               void* stream) {
    if (!e || !self_k || !self_v || !cross_k || !cross_v) return DIA_B200_EINVAL;
    if (e->max_utts <= 0) return DIA_B200_ESTATE;
    if (utterance < 0 || utterance >= e->max_utts || n_layer != e->shape.n_layer || text_len < 0 ||
        text_len > e->shape.max_text_len)
        return DIA_B200_EINVAL;
    ON_DEVICE(e->device);
    const int L = n_layer, UL = kMaxUtt * L;
    CK(cudaStreamSynchronize(S(stream)));                     // the pinned staging buffer is reused
    for (int l = 0; l < L; ++l) {
        if (!self_k[l] || !self_v[l] || !cross_k[l] || !cross_v[l]) return DIA_B200_EINVAL;
        if ((reinterpret_cast<uintptr_t>(self_k[l]) | reinterpret_cast<uintptr_t>(self_v[l]) |
             reinterpret_cast<uintptr_t>(cross_k[l]) | reinterpret_cast<uintptr_t>(cross_v[l])) & 15)
            return DIA_B200_EINVAL;
        e->h_ptrs[utterance * L + l] = static_cast<float*>(self_k[l]);
        e->h_ptrs[UL + utterance * L + l] = static_cast<float*>(self_v[l]);
        e->h_ptrs[2 * UL + utterance * L + l] = const_cast<float*>(static_cast<const float*>(cross_k[l]));
        e->h_ptrs[3 * UL + utterance * L + l] = const_cast<float*>(static_cast<const float*>(cross_v[l]));
    }
    for (int a = 0; a < 4; ++a)
        CK(cudaMemcpyAsync(e->d_ptrs + a * UL + utterance * L, e->h_ptrs + a * UL + utterance * L, sizeof(float*) * L,
                           cudaMemcpyHostToDevice, S(stream)));
    e->utt_text_len[utterance] = text_len;
    e->utt_bound[utterance] = true;
    e->gen_active = false;
    return DIA_B200_OK;
}

int dia_b200_batch_decode_step(dia_b200_engine* e, int n_utterances, const int32_t* tokens, const int32_t* pos_host,
                               const int32_t* slot_host, float* logits, void* stream) {
    int rc = batch_ready(e, n_utterances);
    if (rc) return rc;
    if (!tokens || !pos_host || !slot_host || !logits) return DIA_B200_EINVAL;
    for (int u = 0; u < n_utterances; ++u)
        if (pos_host[u] < 0 || slot_host[u] < 0 || slot_host[u] >= e->shape.max_audio_len) return DIA_B200_EINVAL;
    ON_DEVICE(e->device);
    BatchParams p;
    fill_batch_params(e, p, n_utterances);
    for (int u = 0; u < n_utterances; ++u) { p.utt[u].pos0 = pos_host[u]; p.utt[u].slot0 = slot_host[u]; }
    p.tokens = tokens;
    p.with_sample = 0;
    p.logits = logits;
    return run_batch(e, p, S(stream));
}

int dia_b200_batch_generate_begin(dia_b200_engine* e, int n_utterances, int32_t* const* grids, const dia_b200_gen_params* gp,
                                  void* stream) {
    int rc = batch_ready(e, n_utterances);
    if (rc) return rc;
    if (!grids || !gp) return DIA_B200_EINVAL;
    for (int u = 0; u < n_utterances; ++u) {
        if (!grids[u] || gp[u].prefill_step < 1 || gp[u].max_tokens < 1 || gp[u].max_tokens > e->shape.max_audio_len ||
            gp[u].first_slot < 0 || gp[u].first_slot >= e->shape.max_audio_len)
            return DIA_B200_EINVAL;
        // one sampling configuration per launch (the rows share the sampler code path); seeds and prompts are per utterance
        if (gp[u].cfg_scale != gp[0].cfg_scale || gp[u].temperature != gp[0].temperature || gp[u].top_p != gp[0].top_p ||
            gp[u].top_k != gp[0].top_k || gp[u].max_tokens != gp[0].max_tokens)
            return DIA_B200_EINVAL;
    }
    rc = check_sampling_supported(gp[0].temperature, gp[0].top_k);
    if (rc) return rc;
    ON_DEVICE(e->device);
    CK(cudaStreamSynchronize(S(stream)));
    int dmax = 0;
    for (int c = 0; c < e->shape.channels; ++c) dmax = std::max(dmax, (int)e->shape.delay_pattern[c]);
    std::memset(e->h_gs, 0, sizeof(GenState) * kMaxUtt);
    for (int u = 0; u < n_utterances; ++u) {
        GenState& g = e->h_gs[u];
        g.dec_step = gp[u].prefill_step - 1;
        g.bos_countdown = dmax;
        g.eos_countdown = -1;
        g.finished = (g.dec_step >= gp[u].max_tokens - 1) ? 1 : 0;
        e->bgp[u] = gp[u];
        e->bgrid[u] = grids[u];
        e->bpos[u] = gp[u].prefill_step;
        e->bslot[u] = gp[u].first_slot;
    }
    CK(cudaMemcpyAsync(e->d_gs, e->h_gs, sizeof(GenState) * kMaxUtt, cudaMemcpyHostToDevice, S(stream)));
    std::memset(e->h_err, 0, sizeof(int) * kErrWords);
    e->n_active = n_utterances;
    e->gen_steps = 0;
    e->gen_active = true;
    return DIA_B200_OK;
}

int dia_b200_batch_generate_steps(dia_b200_engine* e, int n_steps, void* stream) {
    if (!e) return DIA_B200_EINVAL;
    if (e->max_utts <= 0 || !e->gen_active) return DIA_B200_ESTATE;
    int rc = batch_ready(e, e->n_active);
    if (rc) return rc;
    if (n_steps < 0) return DIA_B200_EINVAL;
    // Never run past a cache / grid.  Utterances may start at different depths (prompts): one that has used up its budget is
    // finished on the device (dec_step >= max_tokens - 1 <= Lmax - 1) and idles at the last slot; the launch runs as long as the
    // utterance with the most room left.
    {
        int longest = 0;
        for (int u = 0; u < e->n_active; ++u)
            longest = std::max(longest, std::min(e->shape.max_audio_len - e->bslot[u], e->shape.max_audio_len - e->bpos[u]));
        n_steps = std::min(n_steps, longest);
    }
    if (n_steps <= 0) return DIA_B200_OK;
    ON_DEVICE(e->device);
    const int max_per_launch = 65534 / (8 * e->shape.n_layer + 3);
    while (n_steps > 0) {
        const int n = std::min(n_steps, max_per_launch);
        BatchParams p;
        fill_batch_params(e, p, e->n_active);
        p.n_steps = n;
        p.with_sample = 1;
        for (int u = 0; u < e->n_active; ++u) {
            p.utt[u].pos0 = e->bpos[u]; p.utt[u].slot0 = e->bslot[u];
            p.utt[u].grid = e->bgrid[u]; p.utt[u].gs = e->d_gs + u; p.utt[u].seed = e->bgp[u].seed;
        }
        p.cfg_scale = e->bgp[0].cfg_scale; p.temperature = e->bgp[0].temperature; p.top_p = e->bgp[0].top_p;
        p.top_k = e->bgp[0].top_k; p.max_tokens = e->bgp[0].max_tokens; p.draw0 = (unsigned long long)e->gen_steps;
        rc = run_batch(e, p, S(stream));
        if (rc) return rc;
        for (int u = 0; u < e->n_active; ++u) {
            e->bpos[u] = std::min(e->bpos[u] + n, e->shape.max_audio_len);
            e->bslot[u] = std::min(e->bslot[u] + n, e->shape.max_audio_len - 1);
        }
        e->gen_steps += n;
        n_steps -= n;
    }
    return DIA_B200_OK;
}

int dia_b200_batch_generate_status(dia_b200_engine* e, dia_b200_gen_status* out, void* stream) {
    if (!e || !out) return DIA_B200_EINVAL;
    if (e->max_utts <= 0 || e->n_active <= 0) return DIA_B200_ESTATE;
    ON_DEVICE(e->device);
    CK(cudaMemcpyAsync(e->h_gs, e->d_gs, sizeof(GenState) * kMaxUtt, cudaMemcpyDeviceToHost, S(stream)));
    CK(cudaStreamSynchronize(S(stream)));
    for (int u = 0; u < e->n_active; ++u) {
        const GenState& g = e->h_gs[u];
        out[u].dec_step = g.dec_step; out[u].finished = g.finished; out[u].eos_detected = g.eos_detected;
        out[u].eos_countdown = g.eos_countdown; out[u].bos_countdown = g.bos_countdown; out[u].steps_run = g.steps_run;
        out[u].device_error = e->h_err[0];
        out[u].reserved = 0;
    }
    return DIA_B200_OK;
}

int dia_b200_delay_apply_i32(const int32_t* in, int32_t* out, int B, int T, int C, const int32_t* delay_host,
                             int32_t pad_value, int32_t bos_value, void* stream) {
    if (B < 0 || T < 0 || C < 0 || C > DIA_B200_MAX_CHANNELS || !delay_host) return DIA_B200_EINVAL;
    if ((long long)B * T * C == 0) return DIA_B200_OK;
    if (!in || !out || in == out) return DIA_B200_EINVAL;
    CK(launch_delay_apply(in, out, B, T, C, delay_host, pad_value, bos_value, S(stream)));
    g_launches++;
    return DIA_B200_OK;
}

int dia_b200_delay_revert_i32(const int32_t* in, int32_t* out, int B, int T, int C, const int32_t* delay_host,
                              int32_t pad_value, int T_orig, void* stream) {
    if (B < 0 || T < 0 || C < 0 || C > DIA_B200_MAX_CHANNELS || !delay_host) return DIA_B200_EINVAL;
    if ((long long)B * T * C == 0) return DIA_B200_OK;
    if (!in || !out || in == out) return DIA_B200_EINVAL;
    CK(launch_delay_revert(in, out, B, T, C, delay_host, pad_value, T_orig, S(stream)));
    g_launches++;
    return DIA_B200_OK;
}

int dia_b200_finalize_codes_i32(const int32_t* in, int32_t* out, int T, int C, const int32_t* delay_host,
                                int32_t pad_value, int codebook_size, void* stream) {
    if (T < 0 || C <= 0 || C > DIA_B200_MAX_CHANNELS || !delay_host || codebook_size <= 0) return DIA_B200_EINVAL;
    int dmax = 0;
    for (int i = 0; i < C; ++i) dmax = std::max(dmax, (int)delay_host[i]);
    if (T - dmax <= 0) return DIA_B200_OK;
    if (!in || !out || in == out) return DIA_B200_EINVAL;
    CK(launch_finalize_codes(in, out, T, C, delay_host, pad_value, codebook_size, S(stream)));
    g_launches++;
    return DIA_B200_OK;
}

int dia_b200_build_delay_indices(int32_t* t_idx, int64_t* indices, int B, int T, int C, const int32_t* delay_host,
                                 void* stream) {
    if (B < 0 || T < 0 || C < 0 || C > DIA_B200_MAX_CHANNELS || !delay_host) return DIA_B200_EINVAL;
    if ((long long)B * T * C == 0) return DIA_B200_OK;
    if (!t_idx || !indices) return DIA_B200_EINVAL;
    CK(launch_build_delay_indices(t_idx, reinterpret_cast<long long*>(indices), B, T, C, delay_host, S(stream)));
    g_launches++;
    return DIA_B200_OK;
}

int dia_b200_build_revert_indices(int64_t* t_idx, int64_t* indices, int B, int T, int C, const int32_t* delay_host,
                                  void* stream) {
    if (B < 0 || T < 0 || C < 0 || C > DIA_B200_MAX_CHANNELS || !delay_host) return DIA_B200_EINVAL;
    if ((long long)B * T * C == 0) return DIA_B200_OK;
    if (!t_idx || !indices) return DIA_B200_EINVAL;
    CK(launch_build_revert_indices(reinterpret_cast<long long*>(t_idx), reinterpret_cast<long long*>(indices), B, T,
                                   C, delay_host, S(stream)));
    g_launches++;
    return DIA_B200_OK;
}

int dia_b200_dense_prepare_weight(const void* w, int src_dtype, void* wt_bf16, int K, int N, void* stream) {
    if (!w || !wt_bf16 || K <= 0 || N <= 0 || (src_dtype != 0 && src_dtype != 1)) return DIA_B200_EINVAL;
    CK(launch_transpose_to_bf16(w, src_dtype, wt_bf16, K, N, S(stream)));
    g_launches++;
    return DIA_B200_OK;
}

size_t dia_b200_dense_workspace_bytes(int M, int K) { return (M > 0 && K > 0) ? gemm_workspace_bytes(M, K) : 0; }

int dia_b200_dense_forward_fused(const float* x, const float* norm_weight, float eps, const void* wt_bf16,
                                 const float* residual, float* y, void* workspace, int M, int N, int K, void* stream) {
    if (M < 0 || N <= 0 || K <= 0) return DIA_B200_EINVAL;
    if (M == 0) return DIA_B200_OK;
    if (!x || !wt_bf16 || !y || !workspace) return DIA_B200_EINVAL;
    if ((reinterpret_cast<uintptr_t>(wt_bf16) | reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(y) |
         reinterpret_cast<uintptr_t>(residual)) & 15)
        return DIA_B200_EINVAL;
    cudaError_t e = launch_gemm_tcgen05(x, norm_weight, eps, wt_bf16, residual, y, workspace, M, N, K, S(stream));
    if (e == cudaErrorNotSupported) { (void)cudaGetLastError(); return DIA_B200_EUNSUPPORTED; }
    CK(e);
    g_launches += 2;
    return DIA_B200_OK;
}

int dia_b200_dense_forward(const float* x, const void* wt_bf16, float* y, void* workspace, int M, int N, int K, void* stream) {
    return dia_b200_dense_forward_fused(x, nullptr, 0.f, wt_bf16, nullptr, y, workspace, M, N, K, stream);
}

int dia_b200_attention_rows(const float* q, const float* k, const float* v, float* out, int B, int Tq, int Tk, int Hq, int Hkv,
                            int Tk_stride, int mode, const int32_t* n_valid_host, void* stream) {
    if (B < 0 || Tq < 0 || Tk < 0 || Hq <= 0 || Hkv <= 0 || Hq % Hkv || Tk > Tk_stride || mode < 0 || mode > 2 || B > 16)
        return DIA_B200_EINVAL;
    if (B == 0 || Tq == 0) return DIA_B200_OK;
    if (!q || !k || !v || !out) return DIA_B200_EINVAL;
    if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
         reinterpret_cast<uintptr_t>(out)) & 15)
        return DIA_B200_EINVAL;
    if (n_valid_host)
        for (int b = 0; b < B; ++b) if (n_valid_host[b] < 0 || n_valid_host[b] > Tk) return DIA_B200_EINVAL;
    CK(launch_attention_rows(q, k, v, out, B, Tq, Tk, Hq, Hkv, Tk_stride, mode, n_valid_host, S(stream)));
    g_launches++;
    return DIA_B200_OK;
}

int dia_b200_rope_rows(const float* src, float* dst, const float* sin_tab, const float* cos_tab, const int32_t* pos, int B,
                       int T, int H, int rotate, int to_cache, int dst_T, int dst_t0, int n_pos, void* stream) {
    if (B < 0 || T < 0 || H <= 0 || dst_t0 < 0 || (to_cache && dst_t0 + T > dst_T)) return DIA_B200_EINVAL;
    if ((long long)B * T == 0) return DIA_B200_OK;
    if (!src || !dst || (rotate && (!sin_tab || !cos_tab || !pos || n_pos <= 0))) return DIA_B200_EINVAL;
    CK(launch_rope_rows(src, dst, sin_tab, cos_tab, pos, B, T, H, rotate, to_cache, dst_T, dst_t0, n_pos, S(stream)));
    g_launches++;
    return DIA_B200_OK;
}

int dia_b200_rmsnorm_rows(const float* x, const float* weight, float eps, float* y, int M, int D, void* stream) {
    if (M < 0 || D <= 0) return DIA_B200_EINVAL;
    if (M == 0) return DIA_B200_OK;
    if (!x || !weight || !y) return DIA_B200_EINVAL;
    CK(launch_rmsnorm_rows(x, weight, eps, y, M, D, S(stream)));
    g_launches++;
    return DIA_B200_OK;
}

int dia_b200_silu_mul(const float* gu, float* h, int M, int F, void* stream) {
    if (M < 0 || F <= 0) return DIA_B200_EINVAL;
    if (M == 0) return DIA_B200_OK;
    if (!gu || !h) return DIA_B200_EINVAL;
    CK(launch_silu_mul(gu, h, M, F, S(stream)));
    g_launches++;
    return DIA_B200_OK;
}

int dia_b200_embed_rows(const float* table, const int32_t* ids, float* out, int n_rows, int vocab, int D, void* stream) {
    if (n_rows < 0 || vocab <= 0 || D <= 0 || D % 4) return DIA_B200_EINVAL;
    if (n_rows == 0) return DIA_B200_OK;
    if (!table || !ids || !out) return DIA_B200_EINVAL;
    if ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) & 15) return DIA_B200_EINVAL;
    CK(launch_embed_rows(table, ids, out, n_rows, vocab, D, S(stream)));
    g_launches++;
    return DIA_B200_OK;
}

int dia_b200_debug_run_stages(dia_b200_engine* e, const int32_t* tokens, int stage_begin, int stage_end, int pos,
                              int slot, int cooperative, void* stream) {
    int rc = ready(e, true);
    if (rc) return rc;
    if (!stage_range_ok(e, stage_begin, stage_end) || stage_end > 8 * e->shape.n_layer + 2) return DIA_B200_EINVAL;
    if (stage_begin == 0 && !tokens) return DIA_B200_EINVAL;
    ON_DEVICE(e->device);
    StepParams p;
    fill_params(e, p);
    p.tokens = tokens;
    p.stage_begin = stage_begin; p.stage_end = stage_end;
    p.pos0 = pos; p.slot0 = slot;
    p.want_x = 1;
    return run_stages(e, p, cooperative != 0, S(stream));
}

static int buffer_of(dia_b200_engine* e, int which, void** ptr, size_t* bytes) {
    const dia_b200_shape& s = e->shape;
    switch (which) {
        case DIA_B200_BUF_X: *ptr = e->d_x; *bytes = 8 * (size_t)s.d_model; break;
        case DIA_B200_BUF_CTA_TIMING: *ptr = e->d_cta_timing; *bytes = 8 * (size_t)e->G * (8 * s.n_layer + 3); break;
        case DIA_B200_BUF_LOGITS: *ptr = e->d_logits; *bytes = 4 * 2 * (size_t)s.channels * s.vocab; break;
        case DIA_B200_BUF_PRED: *ptr = e->d_pred; *bytes = 4 * (size_t)s.channels; break;
        case DIA_B200_BUF_TIMING: *ptr = e->d_timing; *bytes = 128 * (size_t)kTimingSteps * (8 * s.n_layer + 3); break;
        default: return DIA_B200_EINVAL;
    }
    return DIA_B200_OK;
}

int dia_b200_debug_last_device_error(dia_b200_engine* e, int32_t* out, int n_words) {
    if (!e || !out || n_words < 0) return DIA_B200_EINVAL;
    for (int i = 0; i < n_words && i < kErrWords; ++i) out[i] = reinterpret_cast<volatile int*>(e->h_err)[i];
    return DIA_B200_OK;
}

int dia_b200_debug_enable_timing(dia_b200_engine* e, int enable) {
    if (!e) return DIA_B200_EINVAL;
    e->timing_on = enable != 0;
    e->timing_cta = enable > 0 ? std::min(enable - 1, e->G - 1) : 0;       // enable = 1 + CTA to observe
    return DIA_B200_OK;
}

int dia_b200_debug_read(dia_b200_engine* e, int which, void* host_dst, size_t nbytes, void* stream) {
    if (!e || !host_dst) return DIA_B200_EINVAL;
    void* p; size_t b;
    int rc = buffer_of(e, which, &p, &b);
    if (rc) return rc;
    if (nbytes > b) return DIA_B200_EINVAL;
    ON_DEVICE(e->device);
    CK(cudaMemcpyAsync(host_dst, p, nbytes, cudaMemcpyDeviceToHost, S(stream)));
    CK(cudaStreamSynchronize(S(stream)));
    return DIA_B200_OK;
}

}  // extern "C"
