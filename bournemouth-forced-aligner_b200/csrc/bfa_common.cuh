// bfa_common.cuh -- shared device helpers for libbfa_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bfa_b200.h"

namespace bfa {

constexpr unsigned FULL = 0xffffffffu;
constexpr int MAX_WORDS = BFA_MAX_C / 32;  // class-bitmask words per utterance

// ---- work item: one DP problem (a whole utterance or one silence-anchored speech segment) ----
struct __align__(16) Item {
    long long lp_off;    // float offset of the item's first row inside logp
    long long stat_off;  // global frame number of the item's first row (index into rowstat[])
    long long out_off;   // index into frame_ph/frame_idx where local frame `trim` is written
    long long out_lim;   // exclusive end of the utterance's output region (truncation, :465-467)
    long long seq_off;   // structured: offset into tgt of the first phoneme; explicit: offset into path[]
    int T, L, band, stride;  // stride 0 => explicit path/true_idx arrays
    int n, idx0, trim, n_out;
    int utt, anchor_off, flags;
    int lead;            // floats between the previous 16-byte boundary and the item's first row (0..3): the banded kernel's bulk copies start there
};
enum : int { ITEM_FINAL = 1, ITEM_STATS = 2, ITEM_FLOOR = 4, ITEM_ANCHOR = 8 };

// ---- small PTX wrappers (mbarrier + 1-D bulk async copy = TMA's non-tensor path, SASS UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared bulk copy, completion signalled on an mbarrier (16-B aligned, size % 16 == 0)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// same with an L2 evict-first policy: posteriors are read exactly once
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(pol)
        : "memory");
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, d));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v = max(v, __shfl_xor_sync(FULL, v, d));
    return v;
}

// ---- target-class bitmask: unique_targets = set(seq) - {blank, -100}, p < C (:44-49) ----
// Warp-cooperative; returns this lane's view: w[i] = mask word i (all lanes), all_ok = every id is a valid
// non-blank class, has_sil = the target holds silence_id.
struct TgtInfo {
    uint32_t w[MAX_WORDS];
    bool all_ok, has_sil;
};
// `sw` = MAX_WORDS words of shared memory owned by the calling warp: the lanes OR their targets' bits into it (one
// shared atomic per target instead of a MAX_WORDS-way select on register words), then every lane reads the mask back.
__device__ __forceinline__ TgtInfo target_info(const int32_t* tgt, long long b, long long e, int C, int blank_id, int silence_id,
                                               int lane, uint32_t* sw) {
    TgtInfo r;
    if (lane < MAX_WORDS) sw[lane] = 0;
    __syncwarp();
    bool all_ok = true, has_sil = false;
    for (long long j = b + lane; j < e; j += 32) {
        const int c = tgt[j];
        has_sil |= (c == silence_id);
        if (c == blank_id || c == -100 || c < 0 || c >= C) { all_ok = false; continue; }
        atomicOr(&sw[c >> 5], 1u << (c & 31));
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < MAX_WORDS; ++i) r.w[i] = sw[i];
    r.all_ok = __all_sync(FULL, all_ok);
    r.has_sil = __any_sync(FULL, has_sil);
    return r;
}

// Row statistics of the boosted row (forced_alignment.py:51-54): (max, log sum exp(x + b - max)).
// One warp per row, lane owns classes lane + 32*i; `get(c)` returns the raw value of class c.  Shared by
// rowstat_kernel (planner's silence scan) and the generic Viterbi kernel so both see identical bits.
template <typename Get>
__device__ __forceinline__ float2 row_stats_warp(Get get, int C, int lane, uint32_t tbits, float boost) {
    float v[MAX_WORDS];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < MAX_WORDS; ++i) {
        const int c = lane + 32 * i;
        v[i] = -INFINITY;
        if (c < C) {
            v[i] = get(c) + (((tbits >> i) & 1u) ? boost : 0.0f);
            mx = fmaxf(mx, v[i]);
        }
    }
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_WORDS; ++i)
        if (lane + 32 * i < C) s += expf(v[i] - mx);
    s = warp_sum(s);
    return make_float2(mx, logf(s));
}

// Modified log-prob of one class of one row = what the reference materialises at
// forced_alignment.py:121-129:  boost (+5 on target classes) -> log_softmax -> floor at log(1e-8).
// (m, ls) are the row's max and log-sum-exp of the boosted row (rowstat kernel).
__device__ __forceinline__ float mod_value(float x, bool is_target, bool use_stats, bool do_floor, float boost, float m,
                                           float ls, float min_lp) {
    if (use_stats) x = ((x + (is_target ? boost : 0.0f)) - m) - ls;
    if (do_floor && is_target) x = fmaxf(x, min_lp);
    return x;
}

// Programmatic dependent launch: a kernel launched with the stream-serialization attribute may become resident while the
// kernel before it in the stream is still draining; pdl_wait() blocks until that kernel has completed and its writes are
// visible (a no-op for an ordinary launch), pdl_release() lets the next kernel's CTAs take SM slots as this grid frees them.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace bfa
