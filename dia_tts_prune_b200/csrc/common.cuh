// common.cuh - sm_100a PTX helpers shared by the Dia decode kernels.
//
// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) for the weight /
// KV stream, packed fp32x2 FMA (SASS FFMA2), L2-only loads for data other CTAs
// wrote, watchdog-bounded waits (a hang on a shared GPU box is never acceptable:
// every spin has a cycle budget, after which the kernel records an error code
// and traps).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace dia {

constexpr int kHeadDim = 128;
constexpr int kConsumerWarps = 8;                       // math warps per CTA
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kProducerWarp = 9;                        // on another scheduler than warp 0
constexpr int kNormWarp = 8;                            // gathers the RMSNorm partial sums while the math warps run MMAs
constexpr int kThreads = (kProducerWarp + 1) * 32;      // 8 math warps + the norm warp + the producer warp (one elected lane issues copies)
constexpr int kSlotBytes = 8192;                        // one ring slot = one bulk copy
constexpr int kNumSlots = 20;                           // 160 KB of weights / KV in flight per SM
constexpr int kBStageBytes = 4096;                      // per-warp B-fragment staging of one ring slot (16 k-blocks)
constexpr int kXsBytes = 49152;                         // B staging (GEMM), attention / sampler scratch
constexpr int kRedBytes = 8192;                         // cross-warp reduction scratch (two of these, ping-pong)
constexpr int kMiscBytes = 1024;                        // mbarriers + small shared scalars
constexpr int kSmemBytes = kNumSlots * kSlotBytes + kXsBytes + 2 * kRedBytes + kMiscBytes;
constexpr unsigned long long kWatchdogCycles = 4000000000ull;   // ~2 s at 1.9 GHz

enum DeviceError : int {
    kErrNone = 0,
    kErrGridBarrierTimeout = 1,
    kErrFullBarrierTimeout = 2,
    kErrEmptyBarrierTimeout = 3,
    kErrStepDoneTimeout = 4,
    kErrBadState = 5,
    kErrFlagTimeout = 6,
};
constexpr unsigned kMaxSpins = 1u << 23;                        // polls of one flag word before the watchdog fires

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
static __device__ __noinline__ void ll_timeout(int* err, int code, unsigned info = 0);
__device__ __forceinline__ void ll_check_abort(int* err, unsigned spins, int site, unsigned info);
// the same with a suspend-time hint: the warp sleeps in hardware until the phase completes or ~`ns` elapse,
// instead of burning the issue slots of the scheduler it shares with two math warps
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
// bounded wait; on timeout record `code` and trap (kills the launch, never hangs the box).  The slow path is
// out of line and sleeps in hardware between polls.
static __device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity, int* err, int code, unsigned info) {
    unsigned polls = 0;
#ifdef DIA_NO_WAIT_HINT
    while (!mbar_try_wait(bar, parity)) {
#else
    while (!mbar_try_wait_hint(bar, parity, 20000u)) {
#endif
        if (++polls > 4000000u) ll_timeout(err, code, info);   // seconds, even if a poll returns in a microsecond
        ll_check_abort(err, (polls << 8) | 0xffu, 100 + code, info);   // a poll may sleep 20 us: look every 64 polls
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code, unsigned info = 0) {
    if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity, err, code, info);
}

// ---- bulk async copy global -> shared (TMA engine, no tensor map needed for 1-D) -------------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// One lane of a converged warp (elect.sync).  Code that issues tcgen05 / TMA instructions for the whole CTA runs with the
// WHOLE warp in the loop and only the issue under this predicate: the operands (descriptors, TMEM addresses) are then
// provably warp-uniform and stay in uniform registers.  Behind a `lane == 0` branch the compiler instead wraps every
// UTCHMMA in an ELECT / 5 x R2UR.BROADCAST / branch "waterfall" (measured: ~170 cycles per MMA instead of ~30).
__device__ __forceinline__ uint32_t elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred;
}

// ---- named barrier over the consumer warps only (the producer warp never joins) --------------
__device__ __forceinline__ void consumer_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
}

// ---- loads that must observe other CTAs' writes: L2 only ---------------------------------------
__device__ __forceinline__ float ldcg_f(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float2 ldcg_f2(const float2* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ldcg_f4(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ int ldcg_i(const int* p) { return __ldcg(p); }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_s32(int* p, int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// ---- flag-in-data words ("LL" protocol): one aligned 8-byte word carries payload + sequence flag, is
// written with one strong store and read with one strong load, so a reader that sees the flag has the
// payload - no fence, no barrier, no separate counter between a producing and a consuming CTA.
__device__ __forceinline__ void ll_st(unsigned long long* p, uint32_t lo, uint32_t hi) {
    unsigned long long v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(lo), "r"(hi));
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint2 ll_ld(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    uint2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(r.x), "=r"(r.y) : "l"(v));
    return r;
}
// two adjacent words with one 16-byte load (each 8-byte half carries its own flag)
__device__ __forceinline__ uint4 ll_ld2(const void* p) {
    unsigned long long a, b;
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    uint4 r;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(r.x), "=r"(r.y) : "l"(a));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(r.z), "=r"(r.w) : "l"(b));
    return r;
}
// Watchdog plumbing.  `err` is mapped pinned HOST memory (readable after the context died):
//   [0..7]  first failure: code, block, thread, info, 4 site-specific words
//   [16 + 2 * (block * 12 + warp)]  where every other waiting warp was stuck when it noticed the failure
static __device__ __noinline__ void ll_report(int* err, int site, unsigned info) {
    volatile int* e = reinterpret_cast<volatile int*>(err);
    const int slot = 16 + (blockIdx.x * 12 + (threadIdx.x >> 5)) * 2;
    e[slot] = site;
    e[slot + 1] = (int)info;
    __threadfence_system();
}
static __device__ __noinline__ void ll_timeout(int* err, int code, unsigned info) {
    volatile int* e = reinterpret_cast<volatile int*>(err);
    if (e[0] == 0) { e[0] = code; e[1] = blockIdx.x; e[2] = threadIdx.x; e[3] = (int)info; }
    ll_report(err, code, info);
    for (int i = 0; i < 100; ++i) __nanosleep(1000000);         // let the other waiters report before the context dies
    __trap();
}
// called from spin loops every few thousand polls: has somebody else's watchdog fired?
__device__ __forceinline__ void ll_check_abort(int* err, unsigned spins, int site, unsigned info) {
    if ((spins & 0x3fffu) == 0x3fffu && reinterpret_cast<volatile int*>(err)[0] != 0) {
        ll_report(err, site, info);
        for (;;) __nanosleep(1000000);
    }
}
// one thread spins on one word until its 32-bit flag matches; returns the payload
static __device__ __noinline__ uint32_t ll_wait32_slow(const unsigned long long* p, uint32_t flag, int* err) {
    unsigned spins = 0;
    uint2 w = ll_ld(p);
    while (w.y != flag) {
        if (++spins > kMaxSpins) ll_timeout(err, kErrFlagTimeout + 3, flag);
        ll_check_abort(err, spins, 100 + kErrFlagTimeout + 3, flag);
        w = ll_ld(p);
    }
    return w.x;
}
__device__ __forceinline__ uint32_t ll_wait32(const unsigned long long* p, uint32_t flag, int* err) {
    const uint2 w = ll_ld(p);
    return w.y == flag ? w.x : ll_wait32_slow(p, flag, err);
}

// ---- 2:4 sparse MMA (mma.sp m16n8k32, bf16): metadata layout -------------------------------------------
// Measured on sm_100a with tools/microbench/mma_sp_probe.cu (profiles/r1_mma_sp_metadata_probe.txt), sparsity selector
// 0: lane (g = lane / 4, q = lane % 4) with q < 2 supplies one 32-bit word; nibble j < 4 is the index pair of row g,
// 4-column group 4 q + j of the 16 x 32 tile, nibble 4 + j the one of row g + 8.  A 64-byte metadata block of a
// (16-column tile, 32-row block) stores the word of lane (g, q) at index 2 g + q.
__host__ __device__ inline int sp_meta_word(int n, int gq) { return (n & 7) * 2 + (gq >> 2); }
__host__ __device__ inline int sp_meta_shift(int n, int gq) { return 4 * ((n >> 3) * 4 + (gq & 3)); }

// ---- packed fp32x2 math (Blackwell FFMA2) -------------------------------------------------------
typedef unsigned long long f32x2;   // two packed floats in a 64-bit register
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// two bf16 packed in a u32 -> (lo element, hi element) as exact fp32
__device__ __forceinline__ f32x2 bf16x2_to_f32x2(uint32_t w) {
    return pack2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

}  // namespace dia
