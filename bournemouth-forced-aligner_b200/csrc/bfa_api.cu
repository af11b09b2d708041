// bfa_api.cu -- C-ABI entry points of libbfa_b200.so (see include/bfa_b200.h).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <type_traits>
#include <vector>

#include "assort.cuh"
#include "bfa_common.cuh"
#include "frontend.cuh"
#include "plan.cuh"
#include "viterbi_band3.cuh"
#include "viterbi_generic.cuh"
#include "viterbi_wide.cuh"

using namespace bfa;

namespace {

std::atomic<long long> g_launches{0};
int* g_last_counters = nullptr;   // development aid, see bfa_debug_item_counts
int g_last_direct_B = 0;          // utterances offered to the direct kernel by that call (0: not used)
thread_local char g_cuda_err[256] = "";

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            snprintf(g_cuda_err, sizeof(g_cuda_err), "%s at %s:%d", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return BFA_E_CUDA;                                                                      \
        }                                                                                           \
    } while (0)

#define LAUNCH_CHECK()                                   \
    do {                                                 \
        g_launches.fetch_add(1, std::memory_order_relaxed); \
        CUDA_TRY(cudaGetLastError());                    \
    } while (0)

// Optional timing of the dominant (Viterbi) kernel with CUDA events recorded on the launch stream.
struct Prof {
    std::mutex mu;
    bool on = false;
    int mode = 0;     // 2: also time the planner and the stamp kernel (development)
    int every = 1;    // bracket the banded kernel of every `every`-th bfa_align_batch call only (the events cost device time)
    unsigned calls = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> tagged;   // mode 2: (0 planner, 1 stamps) kernels too
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
} g_prof;

struct DeviceInfo {
    int sms = 0;
    int vg_ctas_per_sm = 0;      // short-path class (L <= 256)
    int vg_ctas_per_sm_big = 0;  // long-path class
    bool band_ok = false;        // banded kernel usable (dynamic smem attribute set)
    int nsmid = 0;               // %nsmid: exclusive upper bound of %smid (the direct kernel's per-SM scratch slots)
    bool ok = false;
};
__global__ void nsmid_kernel(int* out) {
    unsigned n;
    asm("mov.u32 %0, %%nsmid;" : "=r"(n));
    *out = (int)n;
}

// The banded-kernel variants launched per call: 8 lanes per utterance, G groups per lane -> window 24 / 40 / 64 groups;
// each exists specialised for C = 66 (the benchmark width, class loop fully unrolled) and for a run-time C.
constexpr int VW_CTAS = 16;      // DP problems of more than 1024 states in flight (each owns a back-pointer slab)
constexpr int BAND_NV = 3;
constexpr int BAND_G[BAND_NV] = {3, 5, 8};
constexpr int BAND_WARPS = B3_PAIRS;   // (DP, helper) warp pairs per CTA
inline int band_rec_words(int G) { return 6 * G + 1; }   // per lane: its 3G cells as (w0, w1) pairs + the slide word
inline int band_smem_bytes_per_warp(int C, int G) { return (int)band3_smem_per_warp(C, G); }

cudaError_t band_set_attr(int bytes) {
    cudaError_t e = cudaFuncSetAttribute(viterbi_band3_kernel<66, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(viterbi_band3_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(viterbi_band3_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(viterbi_band3_kernel<67, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(viterbi_band3_kernel<17, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(viterbi_band3_direct_kernel<67, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(viterbi_band3_direct_kernel<17, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(viterbi_band3_direct_kernel<66, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(viterbi_band3_direct_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(viterbi_band3_direct_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(viterbi_band3_direct_kernel<66, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(viterbi_band3_direct_kernel<67, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(viterbi_band3_direct_kernel<17, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(viterbi_band3_direct_kernel<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e;
}
constexpr int BAND_SMEM_MAX = 227 * 1024;   // dynamic shared memory one CTA may opt in to on sm_100
inline int band_warps(int smem_per_warp) { return std::max(1, std::min(BAND_WARPS, BAND_SMEM_MAX / smem_per_warp)); }
// Launch with programmatic stream serialization: the grid may become resident while the kernel before it in the stream
// drains; the kernel itself waits (pdl_wait) before it touches anything that kernel wrote.  After anything but a kernel
// (a copy, an event wait) this is an ordinary launch.
template <typename A>
cudaError_t launch_pdl(void (*k)(A), int grid, int block, size_t smem, cudaStream_t st, const A& arg) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, arg);
}
template <typename A>
cudaError_t launch_maybe_pdl(void (*k)(A), int grid, int block, size_t smem, cudaStream_t st, const A& arg, bool pdl) {
    if (pdl) return launch_pdl(k, grid, block, smem, st, arg);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    return cudaLaunchKernelEx(&cfg, k, arg);
}
// The direct kernel (one launch per batch on the common path).  At least half of an SM's shared memory is requested so that two
// CTAs of it can never share an SM: its scratch is indexed by the physical SM.
void direct_launch(const Band3Args& ba, bool exact, int grid, cudaStream_t st, bool pdl, bool logits = false) {
    const size_t smem = std::max((size_t)ba.cls[0].npairs * ba.cls[0].smem_per_warp, (size_t)116 * 1024);
    if (logits) {       // un-normalised logits in (boosting on, hence never `exact`): the reduction also carries the rows' plain log-sum-exp
        if (ba.C == 66) launch_maybe_pdl(viterbi_band3_direct_kernel<66, false, true>, grid, BAND_WARPS * 64, smem, st, ba, pdl);
        else if (ba.C == 67) launch_maybe_pdl(viterbi_band3_direct_kernel<67, false, true>, grid, BAND_WARPS * 64, smem, st, ba, pdl);
        else if (ba.C == 17) launch_maybe_pdl(viterbi_band3_direct_kernel<17, false, true>, grid, BAND_WARPS * 64, smem, st, ba, pdl);
        else launch_maybe_pdl(viterbi_band3_direct_kernel<0, false, true>, grid, BAND_WARPS * 64, smem, st, ba, pdl);
        return;
    }
    if (exact) launch_maybe_pdl(viterbi_band3_direct_kernel<0, true>, grid, BAND_WARPS * 64, smem, st, ba, pdl);
    else if (ba.C == 66) launch_maybe_pdl(viterbi_band3_direct_kernel<66, false>, grid, BAND_WARPS * 64, smem, st, ba, pdl);
    else if (ba.C == 67) launch_maybe_pdl(viterbi_band3_direct_kernel<67, false>, grid, BAND_WARPS * 64, smem, st, ba, pdl);
    else if (ba.C == 17) launch_maybe_pdl(viterbi_band3_direct_kernel<17, false>, grid, BAND_WARPS * 64, smem, st, ba, pdl);
    else launch_maybe_pdl(viterbi_band3_direct_kernel<0, false>, grid, BAND_WARPS * 64, smem, st, ba, pdl);
}
// One launch for the three window classes.  exact: the items carry the caller's log-probs unchanged (no fused log-softmax),
// every decision is taken on the sums.
void band_launch(const Band3Args& ba, bool exact, int grid, cudaStream_t st) {
    size_t smem = 0;
    for (int v = 0; v < BAND_NV; ++v) smem = std::max(smem, (size_t)ba.cls[v].npairs * ba.cls[v].smem_per_warp);
    if (exact) launch_pdl(viterbi_band3_kernel<0, true>, grid, BAND_WARPS * 64, smem, st, ba);
    else if (ba.C == 66) launch_pdl(viterbi_band3_kernel<66, false>, grid, BAND_WARPS * 64, smem, st, ba);
    else if (ba.C == 67) launch_pdl(viterbi_band3_kernel<67, false>, grid, BAND_WARPS * 64, smem, st, ba);
    else if (ba.C == 17) launch_pdl(viterbi_band3_kernel<17, false>, grid, BAND_WARPS * 64, smem, st, ba);
    else launch_pdl(viterbi_band3_kernel<0, false>, grid, BAND_WARPS * 64, smem, st, ba);
}

// One internal side stream per device: when the caller expects items for the exact kernel (BfaShape.reserved), its first pass
// runs there, next to the banded kernel, on SMs the banded kernel leaves free.  Created once, never destroyed; events only,
// so the fork and the join can be captured into a CUDA graph.
struct Fork {
    cudaStream_t s = nullptr, s2 = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr, join2 = nullptr;
    bool ok = false, tried = false;
    std::mutex use;   // one call at a time enqueues its fork / side-stream launches / join on these shared events and streams
};
Fork* fork_stream() {
    static std::mutex mu;
    static Fork cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    Fork& f = cache[dev & 63];
    if (!f.tried) {
        f.tried = true;
        f.ok = cudaStreamCreateWithFlags(&f.s, cudaStreamNonBlocking) == cudaSuccess &&
               cudaStreamCreateWithFlags(&f.s2, cudaStreamNonBlocking) == cudaSuccess &&
               cudaEventCreateWithFlags(&f.join2, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&f.join, cudaEventDisableTiming) == cudaSuccess;
        if (!f.ok) (void)cudaGetLastError();
    }
    return f.ok ? &f : nullptr;
}

int device_info(DeviceInfo& out) {
    static std::mutex mu;
    static DeviceInfo cache[64];
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    DeviceInfo& d = cache[dev & 63];
    if (!d.ok) {
        CUDA_TRY(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
        const size_t smem = sizeof(WarpSmem) * VG_WARPS;
        CUDA_TRY(cudaFuncSetAttribute(viterbi_generic_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(viterbi_generic_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(viterbi_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(WarpSmem) * VW_WARPS)));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.vg_ctas_per_sm, viterbi_generic_kernel<0>, VG_WARPS * 32, smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.vg_ctas_per_sm_big, viterbi_generic_kernel<1>, VG_WARPS * 32, smem));
        if (d.vg_ctas_per_sm < 1) d.vg_ctas_per_sm = 1;
        if (d.vg_ctas_per_sm_big < 1) d.vg_ctas_per_sm_big = 1;
        CUDA_TRY(cudaFuncSetAttribute(assort_confidence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)assort_smem(ASSORT_TS_MAX, ASSORT_SS_MAX)));
        CUDA_TRY(cudaFuncSetAttribute(soft_boundaries_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(silprob_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BAND_SMEM_MAX));
        CUDA_TRY(cudaFuncSetAttribute(silprob_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BAND_SMEM_MAX));
        const int band_smem_max = BAND_SMEM_MAX;
        d.band_ok = band_set_attr(band_smem_max) == cudaSuccess;
        if (!d.band_ok) (void)cudaGetLastError();   // do not leave a sticky error behind: the exact kernel still runs
        {   // one-time query of %nsmid (first call on this device only; never inside a stream capture: bfa_workspace_bytes comes first)
            int* dn = nullptr;
            int hn = 0;
            if (cudaMalloc(&dn, sizeof(int)) == cudaSuccess) {
                nsmid_kernel<<<1, 1>>>(dn);
                if (cudaMemcpy(&hn, dn, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) hn = 0;
                cudaFree(dn);
            }
            (void)cudaGetLastError();
            d.nsmid = hn >= d.sms ? hn : 0;     // 0: direct kernel unavailable
        }
        d.ok = true;
    }
    out = d;
    return BFA_OK;
}

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// Workspace carve-up shared by bfa_workspace_bytes and bfa_align_batch.
struct Layout {
    bool segmenting, want_sil;
    int cb_pitch, sil_nst;
    int wide_L, wide_ctas;       // > 0: some utterance may need more than 1024 path states (viterbi_wide_kernel)
    long long wide_slab_words;
    size_t off_bp_wide;
    int item_cap, gmax, amax, anchor_words, list_ints, max_L, bp_words_per_lane;
    int resident_warps;
    long long slab_words;
    int band_grid, band_smem_per_warp[BAND_NV];
    long long band_slab_words[BAND_NV];
    size_t off_tmask, off_tgtok, off_need, off_sild, off_silunits, off_items_local, off_items, off_fast[BAND_NV], off_lists, off_cbase, off_anchors, off_counters,
        off_pathlp, off_gcls, off_bp, off_bp_band[BAND_NV], total;
    // direct kernel
    bool direct;                 // usable for this shape / parameter set
    int d_tpitch, d_ncap, d_region, d_smem_per_pair, d_npairs;
    long long d_slab_words;
    size_t off_deferred, off_uflag, off_dslab, off_pscr_lp, off_pscr_gs;
};


int make_layout(const BfaParams& p, const BfaShape& s, const DeviceInfo& d, Layout& L) {
    if (s.B < 0 || s.C <= 0 || s.max_T < 0 || s.max_N < 0) return BFA_E_INVALID;
    if (s.C > BFA_MAX_C) return BFA_E_UNSUPPORTED;
    L.segmenting = (p.mode == BFA_MODE_FULL && p.silence_anchors > 0 && p.silence_id >= 0);
    L.gmax = L.segmenting ? s.max_N / 2 + 1 : 0;
    L.amax = L.segmenting ? s.max_T / 2 + 2 : 0;
    L.item_cap = L.segmenting ? L.gmax + 1 : 1;
    L.anchor_words = L.segmenting ? (s.max_T + 6 * L.item_cap) / 8 + L.item_cap + 1 : 0;
    L.list_ints = L.segmenting ? 14 * L.gmax + 4 * L.amax + 16 : 0;
    // longest CTC path: stride*N+1 <= 4*max_N+1; the non-simple paths also guarantee L <= 1.2*(T+6)
    long long maxL = 4LL * s.max_N + 1;
    if (p.mode == BFA_MODE_FULL) {
        long long byT = (long long)(1.2 * (double)(s.max_T + 2 * p.boundary_pad)) + 1;
        if (byT < maxL) maxL = byT;
    }
    if (maxL > BFA_MAX_L) maxL = BFA_MAX_L;   // longer paths are refused per utterance by the planner (BFA_ST_UNSUPPORTED), not per batch
    // paths of more than 1024 states (targets of more than 255 phonemes in one piece) go to the wide kernel: one CTA per problem
    L.wide_L = maxL > VW_SPAN ? (int)maxL : 0;
    L.wide_ctas = L.wide_L ? std::max(1, std::min(s.B, VW_CTAS)) : 0;
    L.wide_slab_words = L.wide_L ? vw_slab_words(s.max_T, L.wide_L) : 0;
    if (maxL > VW_SPAN) maxL = VW_SPAN;
    L.max_L = (int)maxL;
    L.bp_words_per_lane = (maxL > 512) ? 2 : 1;
    L.resident_warps = d.sms * d.vg_ctas_per_sm * VG_WARPS;
    L.slab_words = (long long)(s.max_T + 2) * 32 * L.bp_words_per_lane;
    L.band_grid = d.sms;
    for (int v = 0; v < BAND_NV; ++v) {
        const int G = BAND_G[v];
        L.band_smem_per_warp[v] = band_smem_bytes_per_warp(s.C, G);
        L.band_slab_words[v] = (long long)((s.max_T + 31) / 32 + 1) * band_rec_words(G) * 32;
    }
    size_t o = 0;
    L.off_tmask = o; o = align_up(o + (size_t)s.B * MAX_WORDS * 4);
    L.off_tgtok = o; o = align_up(o + (size_t)s.B * 4);
    L.off_need = o; o = align_up(o + (size_t)s.B * 4);
    // running sums of the silence probabilities (silprob_kernel) for the planner's silence scan: one double per frame
    L.want_sil = L.segmenting && !(p.reserved & BFA_HINT_NO_SIL);
    L.sil_nst = ss_stages(s.C, (size_t)BAND_SMEM_MAX);
    L.cb_pitch = s.max_T / SS_CHUNK + 2;
    L.off_sild = o; o = align_up(o + (L.segmenting ? (size_t)s.total_frames * 8 + 16 : 0));    // also when the caller hinted "no SIL": a wrong hint costs time, never results
    L.off_silunits = o; o = align_up(o + (L.want_sil ? ((size_t)s.total_frames / SS_CHUNK + s.B) * 8 : 0));
    L.off_items_local = o; o = align_up(o + (size_t)s.B * L.item_cap * sizeof(Item));
    L.off_items = o; o = align_up(o + (size_t)s.B * L.item_cap * sizeof(Item));
    for (int v = 0; v < BAND_NV; ++v) { L.off_fast[v] = o; o = align_up(o + (size_t)s.B * L.item_cap * sizeof(Item)); }
    L.off_lists = o; o = align_up(o + (size_t)s.B * L.list_ints * 4);
    L.off_cbase = o; o = align_up(o + (L.segmenting ? (size_t)s.B * L.cb_pitch * 8 : 0));
    L.off_anchors = o; o = align_up(o + (size_t)s.B * L.anchor_words * 4);
    L.off_counters = o; o = align_up(o + 64);
    L.off_pathlp = o; o = align_up(o + (size_t)s.total_frames * 4);
    L.off_gcls = o; o = align_up(o + (size_t)s.total_frames);
    // the banded variants and the exact kernel run concurrently (internal streams): every kernel owns its slabs
    L.off_bp = o; o = align_up(o + (size_t)L.resident_warps * (size_t)L.slab_words * 4);
    L.off_bp_wide = o; o = align_up(o + (size_t)L.wide_ctas * (size_t)L.wide_slab_words * 4);
    for (int v = 0; v < BAND_NV; ++v) {
        L.off_bp_band[v] = o;
        o = align_up(o + (size_t)L.band_grid * BAND_WARPS * (size_t)L.band_slab_words[v] * 4);
    }
    // ---- direct kernel: in-kernel planning + stamps; needs ignore_noise (stamps = phoneme runs) and a shape whose back-trace
    //      staging fits at least 4 pairs per SM.  Scratch is per physical SM (d.nsmid slots).
    L.direct = false;
    L.d_tpitch = (std::max(s.max_T, 1) + 15) & ~15;
    L.d_ncap = std::max(1, std::min(s.max_N, B3_NMAX));
    L.d_region = (int)band3_direct_region(s.C, L.d_tpitch, L.d_ncap);
    L.d_smem_per_pair = (int)band3_smem_per_pair((size_t)L.d_region);
    L.d_npairs = std::min(BAND_WARPS, BAND_SMEM_MAX / L.d_smem_per_pair);
    L.d_slab_words = (long long)((s.max_T + 31) / 32 + 1) * band_rec_words(3) * 32;
    if (d.band_ok && d.nsmid > 0 && s.C <= B3_KK && p.ignore_noise && !(p.reserved & (BFA_FLAG_EXACT_ONLY | BFA_FLAG_NO_DIRECT)) &&
        (p.mode == BFA_MODE_FULL || p.mode == BFA_MODE_SIMPLE) && L.d_npairs >= 4 && s.max_T < (1 << 22) && s.B > 0)
        L.direct = true;
    L.off_deferred = o; o = align_up(o + (L.direct ? (size_t)s.B * 4 : 0));
    L.off_uflag = o; o = align_up(o + (L.direct ? (size_t)s.B * 4 : 0));
    const size_t dslots = L.direct ? (size_t)d.nsmid * L.d_npairs : 0;
    L.off_dslab = o; o = align_up(o + dslots * (size_t)L.d_slab_words * 4);
    L.off_pscr_lp = o; o = align_up(o + dslots * B3_UPW * (size_t)L.d_tpitch * 4);
    L.off_pscr_gs = o; o = align_up(o + dslots * B3_UPW * (size_t)L.d_tpitch);
    L.total = o;
    return BFA_OK;
}

// max_L: upper bound of the path length of any item; the long-path kernel is only launched when
// an item can need it.  Both classes use the same per-warp slabs (kernels are stream-ordered).
int launch_viterbi(VitArgs& va, int max_items, int max_L, const DeviceInfo& d, cudaStream_t st, bool profile = true, int max_ctas = 1 << 30,
                   cudaStream_t st_long = nullptr) {   // st_long: the long-path class runs there, side by side with the short-path class
    int want = (max_items + VG_WARPS - 1) / VG_WARPS;
    if (want < 1) want = 1;
    if (want > max_ctas) want = max_ctas;
    int ctas = want < d.sms * d.vg_ctas_per_sm ? want : d.sms * d.vg_ctas_per_sm;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_prof.mu);
        if (g_prof.on && profile) { e0 = g_prof.get(); e1 = g_prof.get(); }
    }
    if (e0) cudaEventRecord(e0, st);
    launch_pdl(viterbi_generic_kernel<0>, ctas, VG_WARPS * 32, sizeof(WarpSmem) * VG_WARPS, st, va);
    LAUNCH_CHECK();
    if (e0) {
        cudaEventRecord(e1, st);
        std::lock_guard<std::mutex> lk(g_prof.mu);
        g_prof.pending.emplace_back(e0, e1);
    }
    if (max_L > 256) {
        const int ctas0 = ctas;
        ctas = want < d.sms * d.vg_ctas_per_sm_big ? want : d.sms * d.vg_ctas_per_sm_big;
        VitArgs vb = va;
        if (st_long) vb.warp_base = va.warp_base + ctas0 * VG_WARPS;   // its own slabs
        viterbi_generic_kernel<1><<<ctas, VG_WARPS * 32, sizeof(WarpSmem) * VG_WARPS, st_long ? st_long : st>>>(vb);
        LAUNCH_CHECK();
    }
    return BFA_OK;
}

}  // namespace

extern "C" {

int bfa_version(void) { return BFA_VERSION; }

const char* bfa_strerror(int code) {
    switch (code) {
        case BFA_OK: return "ok";
        case BFA_E_INVALID: return "invalid argument";
        case BFA_E_UNSUPPORTED: return "shape outside compiled limits (C > 256)";
        case BFA_E_WORKSPACE: return "workspace too small";
        case BFA_E_CUDA: return "CUDA runtime error";
        default: return "unknown error";
    }
}

const char* bfa_last_cuda_error(void) { return g_cuda_err; }
int bfa_sizeof_params(void) { return (int)sizeof(BfaParams); }
int64_t bfa_launch_count(void) { return (int64_t)g_launches.load(); }

void bfa_default_params(BfaParams* p, int32_t blank_id, int32_t silence_id) {
    memset(p, 0, sizeof(*p));
    p->blank_id = blank_id;
    p->silence_id = silence_id;
    p->silence_anchors = 10;   // AlignmentUtils.__init__ (forced_alignment.py:841)
    p->ignore_noise = 1;
    p->truly_forced = 1;
    p->boost_targets = 1;      // decode_alignments defaults (:858)
    p->enforce_minimum = 1;
    p->max_blanks = 10;        // assort_frames (:777)
    p->boost_factor = 5.0f;    // :29
    p->min_log_prob = logf(1e-8f);  // torch.log(torch.tensor(1e-8)) (:70)
    p->neg_inf = -1000.0f;     // :23
    p->sub_boost = 5.0f;       // :418
    p->boundary_pad = 3;       // :269
    p->min_speech_frames = 20; // :269
    p->mode = BFA_MODE_FULL;
}

size_t bfa_workspace_bytes(const BfaParams* p, const BfaShape* shape) {
    if (!p || !shape) return 0;
    DeviceInfo d;
    if (device_info(d) != BFA_OK) return 0;
    Layout L;
    if (make_layout(*p, *shape, d, L) != BFA_OK) return 0;
    return L.total;
}

static int align_impl(const BfaParams* p, const BfaShape* shape, const float* logp, const int64_t* row_off, const int32_t* T,
                      const int32_t* tgt, const int64_t* tgt_off, int32_t* frame_ph, int32_t* frame_idx,
                      const int64_t* frame_off, float* dp_final, int32_t* status, BfaStamp* stamps, float* conf,
                      int32_t* n_stamps, float* row_lse, void* workspace, size_t workspace_bytes, void* stream);

int bfa_align_batch(const BfaParams* p, const BfaShape* shape, const float* logp, const int64_t* row_off, const int32_t* T,
                    const int32_t* tgt, const int64_t* tgt_off, int32_t* frame_ph, int32_t* frame_idx,
                    const int64_t* frame_off, float* dp_final, int32_t* status, BfaStamp* stamps, float* conf,
                    int32_t* n_stamps, void* workspace, size_t workspace_bytes, void* stream) {
    return align_impl(p, shape, logp, row_off, T, tgt, tgt_off, frame_ph, frame_idx, frame_off, dp_final, status, stamps, conf, n_stamps,
                      nullptr, workspace, workspace_bytes, stream);
}

int bfa_align_batch_logits(const BfaParams* p, const BfaShape* shape, const float* logits, const int64_t* row_off, const int32_t* T,
                           const int32_t* tgt, const int64_t* tgt_off, int32_t* frame_ph, int32_t* frame_idx,
                           const int64_t* frame_off, float* dp_final, int32_t* status, BfaStamp* stamps, float* conf,
                           int32_t* n_stamps, float* row_lse, void* workspace, size_t workspace_bytes, void* stream) {
    if (!row_lse || !p) return BFA_E_INVALID;
    if (p->mode != BFA_MODE_FULL || !p->boost_targets) return BFA_E_UNSUPPORTED;   // without boosting nothing re-normalises the rows
    return align_impl(p, shape, logits, row_off, T, tgt, tgt_off, frame_ph, frame_idx, frame_off, dp_final, status, stamps, conf, n_stamps,
                      row_lse, workspace, workspace_bytes, stream);
}

static int align_impl(const BfaParams* p, const BfaShape* shape, const float* logp, const int64_t* row_off, const int32_t* T,
                      const int32_t* tgt, const int64_t* tgt_off, int32_t* frame_ph, int32_t* frame_idx,
                      const int64_t* frame_off, float* dp_final, int32_t* status, BfaStamp* stamps, float* conf,
                      int32_t* n_stamps, float* row_lse, void* workspace, size_t workspace_bytes, void* stream) {
    if (!p || !shape || !logp || !row_off || !T || !tgt_off || !frame_ph || !frame_idx || !frame_off || !status) return BFA_E_INVALID;
    if ((stamps == nullptr) != (n_stamps == nullptr)) return BFA_E_INVALID;
    if (p->blank_id < 0 || p->blank_id >= shape->C) return BFA_E_INVALID;
    if (shape->B == 0) return BFA_OK;
    if (!tgt && shape->max_N > 0) return BFA_E_INVALID;
    DeviceInfo d;
    int rc = device_info(d);
    if (rc) return rc;
    Layout L;
    rc = make_layout(*p, *shape, d, L);
    if (rc) return rc;
    if (!workspace || workspace_bytes < L.total) return BFA_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    const int B = shape->B, C = shape->C;
    const bool boost = p->boost_targets && p->mode == BFA_MODE_FULL;

    if (stamps && shape->max_stamps <= 0) return BFA_E_INVALID;
    uint32_t* tmask = (uint32_t*)(ws + L.off_tmask);
    double* sild = L.segmenting ? (double*)(ws + L.off_sild) : nullptr;
    float* path_lp = (stamps && conf && !(p->reserved & BFA_FLAG_UNFUSED_CONF)) ? (float*)(ws + L.off_pathlp) : nullptr;
    // counter slots: 0 items of the exact kernel (planner + retries), 1-2 its work counters (short / long class), 3 5 7 items of
    // the banded window classes (4 6 8 their frames), 10 the planner's exact-item count (side-stream pass), 11-12 work counters of the second exact
    // pass, 13 utterances the direct kernel handed back, 14 silence-scan work units, 15 work counter of the wide exact kernel
    int* counters = (int*)(ws + L.off_counters);
    g_last_counters = counters;
    g_last_direct_B = 0;
    const bool fast = (p->reserved & 1) == 0 && C <= B3_KK && d.band_ok;
    // The direct kernel takes every utterance that is one plain stride-4 DP problem and finishes it (frame labels, timestamps,
    // confidences); the planner chain below then only sees what it handed back (device-side list).  With BFA_FLAG_DIRECT_ONLY the
    // chain is not launched at all and such utterances are flagged instead.
    const bool use_direct = L.direct && fast;
    // bfa_align_batch_logits.  The planner chain can follow when its silence pass runs (it reads every row of the chain's utterances
    // and leaves their log-sum-exp behind for the stamp kernel); otherwise only the one-kernel pass runs and flags what it cannot finish.
    const bool logits = row_lse != nullptr;
    const bool chain_logits = logits && L.want_sil && shape->max_T > 0 && p->silence_id < C && (!(stamps && conf) || path_lp != nullptr);
    if (logits && !use_direct && !chain_logits) return BFA_E_UNSUPPORTED;
    const bool direct_only = use_direct && ((p->reserved & BFA_FLAG_DIRECT_ONLY) != 0 || (logits && !chain_logits));
    int* deferred = use_direct && !direct_only ? (int*)(ws + L.off_deferred) : nullptr;
    int32_t* uflag = use_direct && !direct_only ? (int32_t*)(ws + L.off_uflag) : nullptr;
    if (!direct_only) CUDA_TRY(cudaMemsetAsync(counters, 0, 64, st));
    if (use_direct) {
        g_last_direct_B = direct_only ? 0 : B;
        Band3Args da;
        memset(&da, 0, sizeof(da));
        da.p = *p; da.C = C; da.logp = logp; da.tgt = tgt;
        da.frame_ph = frame_ph; da.frame_idx = frame_idx; da.dp_final = dp_final;
        da.cls[0].bp_scratch = (uint32_t*)(ws + L.off_dslab);
        da.cls[0].bp_slab_words = L.d_slab_words;
        da.cls[0].smem_per_warp = L.d_smem_per_pair;
        da.cls[0].npairs = L.d_npairs;
        da.cls[0].region = L.d_region;
        da.B = B; da.row_off = (const long long*)row_off; da.T = T; da.tgt_off = (const long long*)tgt_off;
        da.frame_off = (const long long*)frame_off; da.status = status; da.stamps = stamps; da.conf = stamps ? conf : nullptr;
        da.n_stamps = n_stamps; da.max_stamps = shape->max_stamps;
        da.uflag = uflag; da.deferred = deferred; da.n_deferred = counters + 13;
        const bool spec = path_lp != nullptr && !(p->reserved & BFA_FLAG_NO_SPEC);
        da.pscr_lp = spec ? (float*)(ws + L.off_pscr_lp) : nullptr;
        da.pscr_gs = spec ? (unsigned char*)(ws + L.off_pscr_gs) : nullptr;
        da.tpitch = L.d_tpitch; da.ncap = L.d_ncap; da.nslots = d.nsmid; da.direct_only = direct_only ? 1 : 0;
        da.row_lse = row_lse;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        {
            std::lock_guard<std::mutex> lk(g_prof.mu);
            if (g_prof.on && (g_prof.calls++ % (unsigned)g_prof.every) == 0) { e0 = g_prof.get(); e1 = g_prof.get(); }
        }
        if (e0) cudaEventRecord(e0, st);
        direct_launch(da, !boost, d.sms, st, direct_only && (p->reserved & BFA_FLAG_PIPELINED) != 0, logits);
        LAUNCH_CHECK();
        if (e0) {
            cudaEventRecord(e1, st);
            std::lock_guard<std::mutex> lk(g_prof.mu);
            g_prof.pending.emplace_back(e0, e1);
        }
        if (direct_only) return BFA_OK;
    }

    // The planner's silence scan needs exp(modified_lp[t, silence_id]) of every row of the utterances whose target holds
    // silence_id (forced_alignment.py:297, :503-512): one streaming read of those utterances, left behind as running sums
    // (silscan.cuh).  The Viterbi kernels fuse boost + log_softmax + floor into their own row loads.
    if (L.want_sil && shape->max_T > 0) {
        SilArgs sa;
        sa.p = *p; sa.B = B; sa.C = C; sa.nst = L.sil_nst;
        sa.keep_in_l2 = (double)shape->total_frames * C * 4.0 < 48.0 * 1024 * 1024 ? 1 : 0;   // well inside the 126 MB L2
        sa.logp = logp; sa.row_off = (const long long*)row_off; sa.T = T; sa.tgt = tgt; sa.tgt_off = (const long long*)tgt_off;
        sa.frame_off = (const long long*)frame_off; sa.D = sild; sa.deferred = deferred; sa.n_deferred = counters + 13;
        sa.tmask = tmask; sa.units = (int2*)(ws + L.off_silunits); sa.n_units = counters + 14;
        sa.row_lse = chain_logits ? row_lse : nullptr; sa.list_all = chain_logits ? 1 : 0;
        launch_maybe_pdl(silunits_kernel, (B + 7) / 8, 256, 0, st, sa, use_direct);
        LAUNCH_CHECK();
        const long long units = (long long)shape->total_frames / SS_CHUNK + B;      // upper bound; the real count stays on the device
        const int grid = (int)std::max(1LL, std::min((long long)d.sms, (units + SS_WARPS - 1) / SS_WARPS));
        if (chain_logits) launch_maybe_pdl(silprob_kernel<true>, grid, SS_WARPS * 32, ss_smem_per_warp(C, sa.nst) * SS_WARPS, st, sa, false);
        else launch_maybe_pdl(silprob_kernel<false>, grid, SS_WARPS * 32, ss_smem_per_warp(C, sa.nst) * SS_WARPS, st, sa, false);
        LAUNCH_CHECK();
    }
    PlanArgs pa;
    pa.p = *p; pa.B = B; pa.C = C; pa.max_T = shape->max_T; pa.max_N = shape->max_N;
    pa.logp = logp; pa.row_off = (const long long*)row_off; pa.T = T; pa.tgt = tgt; pa.tgt_off = (const long long*)tgt_off;
    pa.frame_off = (const long long*)frame_off; pa.D = sild; pa.sil_ready = L.want_sil ? 1 : 0; pa.tmask = tmask;
    pa.frame_ph = frame_ph; pa.frame_idx = frame_idx; pa.dp_final = dp_final; pa.status = status;
    pa.item_cap = L.item_cap; pa.gmax = L.gmax; pa.amax = L.amax; pa.anchor_words = L.anchor_words;
    pa.items_local = (Item*)(ws + L.off_items_local); pa.items = (Item*)(ws + L.off_items);
    pa.n_items = counters; pa.lists = (int32_t*)(ws + L.off_lists); pa.list_ints = L.list_ints;
    pa.deferred = deferred; pa.n_deferred = counters + 13;
    for (int v = 0; v < BAND_NV; ++v) { pa.fast_items[v] = (Item*)(ws + L.off_fast[v]); pa.n_fast[v] = counters + 3 + 2 * v; pa.frames_fast[v] = counters + 4 + 2 * v; }
    pa.fast_enable = fast ? 1 : 0; pa.path_lp = path_lp;
    pa.cbase = (double*)(ws + L.off_cbase); pa.cb_pitch = L.cb_pitch; pa.anchors = (uint32_t*)(ws + L.off_anchors);
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_prof.mu);
        if (g_prof.on && g_prof.mode == 2) { pe0 = g_prof.get(); pe1 = g_prof.get(); }
    }
    if (pe0) cudaEventRecord(pe0, st);
    plan_kernel<<<(B + 7) / 8, 256, 0, st>>>(pa);
    LAUNCH_CHECK();
    if (pe0) {
        cudaEventRecord(pe1, st);
        std::lock_guard<std::mutex> lk(g_prof.mu);
        g_prof.tagged.push_back({0, {pe0, pe1}});
    }

    long long max_items = (long long)B * L.item_cap;
    const int max_items_i = (int)(max_items > (1 << 30) ? (1 << 30) : max_items);
    VitArgs va;
    va.p = *p; va.C = C; va.logp = logp; va.tmask = tmask; va.tgt = tgt;
    va.path = nullptr; va.true_idx = nullptr; va.anchors = pa.anchors; va.items = pa.items;
    va.n_items = counters; va.work_counter = counters + 1; va.first = nullptr; va.warp_base = 0;
    va.frame_ph = frame_ph; va.frame_idx = frame_idx; va.dp_final = dp_final; va.status = status; va.final_state = nullptr;
    va.path_lp = path_lp;
    va.bp_scratch = (uint32_t*)(ws + L.off_bp); va.bp_slab_words = L.slab_words;
    if (fast) {
        Band3Args ba;
        ba.p = *p; ba.C = C; ba.logp = logp; ba.tgt = tgt; ba.tmask = tmask;
        ba.retry_items = pa.items; ba.n_retry = counters;
        ba.frame_ph = frame_ph; ba.frame_idx = frame_idx; ba.dp_final = dp_final; ba.path_lp = path_lp;
        ba.guess_cls = (path_lp && !(p->reserved & BFA_FLAG_NO_SPEC)) ? (unsigned char*)(ws + L.off_gcls) : nullptr;
        for (int v = 0; v < BAND_NV; ++v) {
            ba.cls[v].items = pa.fast_items[v]; ba.cls[v].n_items = pa.n_fast[v]; ba.cls[v].n_frames = pa.frames_fast[v];
            ba.cls[v].bp_scratch = (uint32_t*)(ws + L.off_bp_band[v]);
            ba.cls[v].bp_slab_words = L.band_slab_words[v];
            ba.cls[v].smem_per_warp = L.band_smem_per_warp[v];
            ba.cls[v].npairs = band_warps(L.band_smem_per_warp[v]);
            ba.cls[v].region = (int)band3_stage_region(C, BAND_G[v]);
        }
        // The caller expects `hint` items for the exact kernel (utterances too dense for stride 4, ...): its first pass, over the
        // planner's list, runs on a side stream on a few SMs of its own while the banded kernel takes the rest; what the banded
        // kernel sends back is done afterwards.  One CTA of the exact kernel (8 items at a time) per reserved SM.
        const int hint = shape->reserved > 0 ? shape->reserved : 0;
        Fork* fk = hint > 0 ? fork_stream() : nullptr;
        // the events and side streams are shared by every call on this device: two host threads must not interleave their
        // record / wait sequences (thread A's side-stream pass would wait for thread B's planner); held until the join is enqueued
        std::unique_lock<std::mutex> fork_guard;
        if (fk) fork_guard = std::unique_lock<std::mutex>(fk->use);
        int band_grid = L.band_grid;
        if (fk) {
            const bool two = L.max_L > 256;      // short-path and long-path classes of the exact kernel side by side
            const int reserve = std::max(1, std::min(d.sms / (two ? 8 : 4), (hint + VG_WARPS - 1) / VG_WARPS));
            band_grid = d.sms - (two ? 2 : 1) * reserve;
            CUDA_TRY(cudaMemcpyAsync(counters + 10, counters, sizeof(int), cudaMemcpyDeviceToDevice, st));   // the planner's count
            CUDA_TRY(cudaEventRecord(fk->fork, st));
            CUDA_TRY(cudaStreamWaitEvent(fk->s, fk->fork, 0));
            if (two) CUDA_TRY(cudaStreamWaitEvent(fk->s2, fk->fork, 0));
            va.n_items = counters + 10;
            rc = launch_viterbi(va, max_items_i, L.max_L, d, fk->s, false, reserve, two ? fk->s2 : nullptr);
            if (rc) return rc;
            CUDA_TRY(cudaEventRecord(fk->join, fk->s));
            if (two) CUDA_TRY(cudaEventRecord(fk->join2, fk->s2));
            va.n_items = counters; va.first = counters + 10; va.work_counter = counters + 11;   // second pass: the retries
        }
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        {
            std::lock_guard<std::mutex> lk(g_prof.mu);
            if (!use_direct && g_prof.on && (g_prof.calls++ % (unsigned)g_prof.every) == 0) { e0 = g_prof.get(); e1 = g_prof.get(); }
        }
        if (e0) cudaEventRecord(e0, st);
        band_launch(ba, !boost, band_grid, st);       // all three window classes in one launch
        LAUNCH_CHECK();
        if (e0) {
            cudaEventRecord(e1, st);
            std::lock_guard<std::mutex> lk(g_prof.mu);
            g_prof.pending.emplace_back(e0, e1);
        }
        if (fk) {
            CUDA_TRY(cudaStreamWaitEvent(st, fk->join, 0));
            if (L.max_L > 256) CUDA_TRY(cudaStreamWaitEvent(st, fk->join2, 0));
        }
        if (fork_guard.owns_lock()) fork_guard.unlock();
        // the exact kernel: what the planner gave it (unless that ran on the side stream) plus what the banded kernel sent back
        rc = launch_viterbi(va, max_items_i, L.max_L, d, st, false);
        if (rc) return rc;
    } else {
        rc = launch_viterbi(va, max_items_i, L.max_L, d, st, true);
        if (rc) return rc;
    }
    if (L.wide_L) {          // what the exact kernel skipped: the planner's items of more than 1024 states, anywhere in its list
        VitArgs vw = va;
        vw.n_items = counters; vw.first = nullptr; vw.work_counter = counters + 15; vw.warp_base = 0;
        vw.bp_scratch = (uint32_t*)(ws + L.off_bp_wide); vw.bp_slab_words = L.wide_slab_words;
        viterbi_wide_kernel<<<L.wide_ctas, VW_WARPS * 32, sizeof(WarpSmem) * VW_WARPS, st>>>(vw);
        LAUNCH_CHECK();
    }

    if (stamps) {
        AssortArgs aa;
        aa.p = *p; aa.B = B; aa.C = C; aa.max_stamps = shape->max_stamps; aa.logp = logp; aa.row_off = (const long long*)row_off;
        aa.T = T; aa.frame_off = (const long long*)frame_off; aa.frame_ph = frame_ph; aa.frame_idx = frame_idx;
        aa.status = status; aa.stamps = stamps; aa.conf = conf; aa.n_stamps = n_stamps; aa.path_lp = path_lp;
        aa.uflag = uflag; aa.row_lse = chain_logits ? row_lse : nullptr;
        aa.ts = shape->max_N < 32768 ? assort_ts(shape->max_T) : 0;   // staged frames pack (idx, phoneme) into 16 + 16 bits
        aa.ss = assort_ss(shape->max_stamps);
        cudaEvent_t ae0 = nullptr, ae1 = nullptr;
        {
            std::lock_guard<std::mutex> lk(g_prof.mu);
            if (g_prof.on && g_prof.mode == 2) { ae0 = g_prof.get(); ae1 = g_prof.get(); }
        }
        if (ae0) cudaEventRecord(ae0, st);
        launch_pdl(assort_confidence_kernel, (B + ASSORT_WARPS - 1) / ASSORT_WARPS, ASSORT_WARPS * 32, assort_smem(aa.ts, aa.ss), st, aa);
        LAUNCH_CHECK();
        if (ae0) {
            cudaEventRecord(ae1, st);
            std::lock_guard<std::mutex> lk(g_prof.mu);
            g_prof.tagged.push_back({1, {ae0, ae1}});
        }
    }
    return BFA_OK;
}

size_t bfa_viterbi_paths_workspace_bytes(int32_t n_items, int32_t max_T, int32_t max_L) {
    DeviceInfo d;
    if (device_info(d) != BFA_OK) return 0;
    size_t warps = (size_t)d.sms * d.vg_ctas_per_sm * VG_WARPS;
    size_t slab = (size_t)(max_T + 2) * 32 * (max_L > 512 ? 2 : 1) * 4;
    return align_up((size_t)n_items * sizeof(Item)) + align_up(64) + align_up(warps * slab);
}

int bfa_viterbi_paths(const BfaParams* p, int32_t n_items, int32_t C, int32_t max_T, int32_t max_L, const float* logp,
                      const int64_t* row_off, const int32_t* T, const int32_t* path, const int32_t* true_idx,
                      const int64_t* path_off, const int32_t* Lp, const int32_t* band, int32_t* frame_ph, int32_t* frame_idx,
                      const int64_t* frame_off, float* dp_final, int32_t* final_state, void* workspace, size_t workspace_bytes,
                      void* stream) {
    if (!p || !logp || !row_off || !T || !path || !path_off || !Lp || !band || !frame_ph || !frame_idx || !frame_off) return BFA_E_INVALID;
    if (n_items <= 0) return n_items == 0 ? BFA_OK : BFA_E_INVALID;
    if (C <= 0 || C > BFA_MAX_C || max_L > VW_SPAN) return BFA_E_UNSUPPORTED;   // explicit paths: the per-warp exact kernel only
    if (p->blank_id < 0 || p->blank_id >= C) return BFA_E_INVALID;
    DeviceInfo d;
    int rc = device_info(d);
    if (rc) return rc;
    if (!workspace || workspace_bytes < bfa_viterbi_paths_workspace_bytes(n_items, max_T, max_L)) return BFA_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    Item* items = (Item*)ws;
    int* counters = (int*)(ws + align_up((size_t)n_items * sizeof(Item)));
    uint32_t* bp = (uint32_t*)((char*)counters + align_up(64));
    CUDA_TRY(cudaMemsetAsync(counters, 0, 64, st));
    items_from_arrays_kernel<<<(n_items + 127) / 128, 128, 0, st>>>(n_items, C, (const long long*)row_off, T, (const long long*)path_off,
                                                                    Lp, band, (const long long*)frame_off, items, counters);
    LAUNCH_CHECK();
    VitArgs va;
    va.p = *p; va.C = C; va.logp = logp; va.tmask = nullptr; va.tgt = nullptr;
    va.path = path; va.true_idx = true_idx; va.anchors = nullptr; va.items = items;
    va.n_items = counters; va.work_counter = counters + 1; va.first = nullptr; va.warp_base = 0;
    va.frame_ph = frame_ph; va.frame_idx = frame_idx; va.dp_final = dp_final; va.status = nullptr; va.final_state = final_state;
    va.path_lp = nullptr;
    va.bp_scratch = bp; va.bp_slab_words = (long long)(max_T + 2) * 32 * (max_L > 512 ? 2 : 1);
    return launch_viterbi(va, n_items, max_L, d, st);
}

int bfa_alignment_score_batch(int32_t B, int32_t C, const float* logp, const int64_t* row_off, const int32_t* T,
                              const int32_t* frame_ph, const int64_t* frame_off, double* score, void* stream) {
    if (!logp || !row_off || !T || !frame_ph || !frame_off || !score || C <= 0) return BFA_E_INVALID;
    if (B <= 0) return B == 0 ? BFA_OK : BFA_E_INVALID;
    alignment_score_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(B, C, logp, (const long long*)row_off, T, frame_ph,
                                                                          (const long long*)frame_off, score);
    LAUNCH_CHECK();
    return BFA_OK;
}

int bfa_stitch_log_softmax(int32_t B, int32_t n_windows, int32_t frames_per_window, int32_t C, int32_t total_frames,
                           const float* window_logits, int64_t in_pitch, const float* window_weights, float* logp_out,
                           int64_t out_pitch, void* stream) {
    if (!window_logits || !logp_out || C <= 0 || total_frames < 0 || frames_per_window < 0) return BFA_E_INVALID;
    if (C > BFA_MAX_C) return BFA_E_UNSUPPORTED;
    if (frames_per_window > 0 && (!window_weights || n_windows <= 0 || frames_per_window < 2)) return BFA_E_INVALID;
    if (B <= 0 || total_frames == 0) return B >= 0 ? BFA_OK : BFA_E_INVALID;
    DeviceInfo d;
    int rc = device_info(d);
    if (rc) return rc;
    StitchArgs a;
    a.B = B; a.W = n_windows; a.fpw = frames_per_window; a.C = C; a.total_frames = total_frames;
    a.sf = frames_per_window > 0 ? frames_per_window / 2 : 1;                     // stride_frames (windowing.py:128)
    a.in = window_logits; a.weights = window_weights; a.out = logp_out; a.in_pitch_b = in_pitch; a.out_pitch_b = out_pitch;
    const long long rows = (long long)B * total_frames;
    const int grid = (int)std::min((rows + STITCH_WARPS - 1) / STITCH_WARPS, (long long)d.sms * 8);
    if (frames_per_window == 0 && rows < (1LL << 31)) {
        const long long units = (rows + LS_ROWS - 1) / LS_ROWS;
        const int g2 = (int)std::min((units + STITCH_WARPS - 1) / STITCH_WARPS, (long long)d.sms * 8);
        if (C <= 32) log_softmax_rows_kernel<1><<<g2, STITCH_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
        else if (C <= 64) log_softmax_rows_kernel<2><<<g2, STITCH_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
        else if (C <= 96) log_softmax_rows_kernel<3><<<g2, STITCH_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
        else log_softmax_rows_kernel<MAX_WORDS><<<g2, STITCH_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    } else {
        if (C <= 32) stitch_log_softmax_kernel<1><<<grid, STITCH_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
        else if (C <= 64) stitch_log_softmax_kernel<2><<<grid, STITCH_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
        else if (C <= 96) stitch_log_softmax_kernel<3><<<grid, STITCH_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
        else stitch_log_softmax_kernel<MAX_WORDS><<<grid, STITCH_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    }
    LAUNCH_CHECK();
    return BFA_OK;
}

int bfa_confidence_batch_lse(int32_t B, int32_t C, const float* logits, const int64_t* row_off, const int32_t* T_conf,
                             const BfaStamp* stamps, const int32_t* n_stamps, int32_t max_stamps, float* conf,
                             const float* row_lse, const int64_t* lse_off, void* stream) {
    if (!logits || !row_off || !T_conf || !stamps || !n_stamps || !conf || C <= 0 || max_stamps <= 0) return BFA_E_INVALID;
    if ((row_lse == nullptr) != (lse_off == nullptr)) return BFA_E_INVALID;
    if (B <= 0) return B == 0 ? BFA_OK : BFA_E_INVALID;
    confidence_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(B, C, logits, (const long long*)row_off, T_conf, stamps, n_stamps,
                                                                     max_stamps, conf, row_lse, (const long long*)lse_off);
    LAUNCH_CHECK();
    return BFA_OK;
}

int bfa_confidence_batch(int32_t B, int32_t C, const float* logp, const int64_t* row_off, const int32_t* T_conf,
                         const BfaStamp* stamps, const int32_t* n_stamps, int32_t max_stamps, float* conf, void* stream) {
    return bfa_confidence_batch_lse(B, C, logp, row_off, T_conf, stamps, n_stamps, max_stamps, conf, nullptr, nullptr, stream);
}

int bfa_soft_boundaries_batch(int32_t B, int32_t C, const float* logp, const int64_t* row_off, const int32_t* T, BfaStamp* stamps,
                              const int32_t* n_stamps, int32_t max_stamps, int32_t boundary_softness, void* stream) {
    return bfa_soft_boundaries_batch_lse(B, C, logp, row_off, T, stamps, n_stamps, max_stamps, boundary_softness, nullptr, nullptr, stream);
}

int bfa_soft_boundaries_batch_lse(int32_t B, int32_t C, const float* logp, const int64_t* row_off, const int32_t* T, BfaStamp* stamps,
                                  const int32_t* n_stamps, int32_t max_stamps, int32_t boundary_softness,
                                  const float* row_lse, const int64_t* lse_off, void* stream) {
    if (!logp || !row_off || !T || !stamps || !n_stamps || C <= 0 || max_stamps <= 0) return BFA_E_INVALID;
    if ((row_lse == nullptr) != (lse_off == nullptr)) return BFA_E_INVALID;
    if (B <= 0) return B == 0 ? BFA_OK : BFA_E_INVALID;
    const size_t smem = (size_t)SOFT_WARPS * max_stamps * sizeof(double);
    if (smem > 200 * 1024) return BFA_E_UNSUPPORTED;
    DeviceInfo d;
    int rc = device_info(d);     // sets the kernel's shared-memory attribute once per DEVICE
    if (rc) return rc;
    const double t1 = pow(10.0, -1.0 * 3.0), t2 = pow(10.0, -1.0 * (double)boundary_softness);   // core.py:699-701
    soft_boundaries_kernel<<<(B + SOFT_WARPS - 1) / SOFT_WARPS, SOFT_WARPS * 32, smem, (cudaStream_t)stream>>>(
        B, C, logp, (const long long*)row_off, T, stamps, n_stamps, max_stamps, t1, t2, row_lse, (const long long*)lse_off);
    LAUNCH_CHECK();
    return BFA_OK;
}

// Development aid: per-phase warp-clock sums of the banded kernel (all zero unless built with -DBFA_PHASE_PROF).
// Development aid: the item counters of the most recent bfa_align_batch on this thread's device (blocks on the device):
// out[0] = items the exact kernel ran (ineligible + retried), out[1..3] = items of the banded kernels (24/40/64-group window).
int bfa_debug_item_counts(int32_t* out4) {
    if (!out4 || !g_last_counters) return BFA_E_INVALID;
    int h[16];
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(h, g_last_counters, sizeof(h), cudaMemcpyDeviceToHost));
    out4[0] = h[0]; out4[1] = h[3]; out4[2] = h[5]; out4[3] = h[7];
    if (g_last_direct_B > 0) out4[1] += g_last_direct_B - h[13];   // finished by the direct kernel (24-group window as well)
    return BFA_OK;
}

int bfa_debug_ctas(unsigned long long* out320, int reset) {   // development: per-CTA max DP task cycles + SM id of the banded kernel
#ifdef BFA_PHASE_PROF
    if (out320) CUDA_TRY(cudaMemcpyFromSymbol(out320, g_b3_cta, sizeof(unsigned long long) * 320));
    if (reset) { static unsigned long long z[320]; CUDA_TRY(cudaMemcpyToSymbol(g_b3_cta, z, sizeof(z))); }
#else
    if (out320) memset(out320, 0, sizeof(unsigned long long) * 320);
    (void)reset;
#endif
    return BFA_OK;
}

int bfa_debug_warps(unsigned long long* out32, int reset) {   // development: mean task cycles per warp id of the banded kernel
#ifdef BFA_PHASE_PROF
    if (out32) CUDA_TRY(cudaMemcpyFromSymbol(out32, g_b3_warp, sizeof(unsigned long long) * 32));
    if (reset) { unsigned long long z[32] = {0}; CUDA_TRY(cudaMemcpyToSymbol(g_b3_warp, z, sizeof(z))); }
#else
    if (out32) memset(out32, 0, sizeof(unsigned long long) * 32);
    (void)reset;
#endif
    return BFA_OK;
}

int bfa_debug_fin(unsigned long long* out32, int reset) {   // development: sub-phases of the direct kernel's finishing pass, [which * 16 + i]
#ifdef BFA_PHASE_PROF
    if (out32) CUDA_TRY(cudaMemcpyFromSymbol(out32, g_b3_fin, sizeof(unsigned long long) * 32));
    if (reset) { unsigned long long z[32] = {0}; CUDA_TRY(cudaMemcpyToSymbol(g_b3_fin, z, sizeof(z))); }
#else
    if (out32) memset(out32, 0, sizeof(unsigned long long) * 32);
    (void)reset;
#endif
    return BFA_OK;
}

int bfa_debug_phases(unsigned long long* out16, int reset) {   // 32 counters: [0,16) DP warps, [16,32) helper warps
#ifdef BFA_PHASE_PROF
    if (out16) CUDA_TRY(cudaMemcpyFromSymbol(out16, g_b3_phase, sizeof(unsigned long long) * 32));
    if (reset) { unsigned long long z[32] = {0}; z[15] = z[31] = ~0ull; CUDA_TRY(cudaMemcpyToSymbol(g_b3_phase, z, sizeof(z))); }
#else
    if (out16) memset(out16, 0, sizeof(unsigned long long) * 32);
    (void)reset;
#endif
    return BFA_OK;
}

// low byte 1: the dominant kernel; 2 (development): the planner and the stamp kernel as well.  Bits 8-15, when non-zero:
// sample the banded kernel of every n-th bfa_align_batch call only, starting with the next call.
void bfa_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    g_prof.on = (on & 0xff) != 0;
    g_prof.mode = on & 0xff;
    g_prof.every = ((on >> 8) & 0xff) ? ((on >> 8) & 0xff) : 1;
    g_prof.calls = 0;
}

// Development aid (mode 2): mean device time in ms of the planner (out2[0]) and of the stamp kernel (out2[1]) since the last read.
int bfa_profile_read_aux(float* out2) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    float sum[2] = {0.f, 0.f};
    int n[2] = {0, 0};
    for (auto& t : g_prof.tagged) {
        float ms = 0.f;
        CUDA_TRY(cudaEventSynchronize(t.second.second));
        CUDA_TRY(cudaEventElapsedTime(&ms, t.second.first, t.second.second));
        sum[t.first] += ms; ++n[t.first];
        g_prof.pool.push_back(t.second.first); g_prof.pool.push_back(t.second.second);
    }
    g_prof.tagged.clear();
    if (out2) { out2[0] = n[0] ? sum[0] / n[0] : 0.f; out2[1] = n[1] ? sum[1] / n[1] : 0.f; }
    return BFA_OK;
}

// Sum of the device time of the dominant kernel launches recorded since the last read; blocks on them.
int bfa_profile_read(float* dominant_ms, int32_t* n_launches) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    float total = 0.f;
    int n = 0;
    for (auto& pr : g_prof.pending) {
        float ms = 0.f;
        CUDA_TRY(cudaEventSynchronize(pr.second));
        CUDA_TRY(cudaEventElapsedTime(&ms, pr.first, pr.second));
        total += ms; ++n;
        g_prof.pool.push_back(pr.first); g_prof.pool.push_back(pr.second);
    }
    g_prof.pending.clear();
    if (dominant_ms) *dominant_ms = total;
    if (n_launches) *n_launches = n;
    return BFA_OK;
}

int bfa_assort_batch(const BfaParams* p, int32_t B, const int32_t* T, const int64_t* frame_off, const int32_t* frame_ph,
                     const int32_t* frame_idx, int32_t* status, BfaStamp* stamps, int32_t* n_stamps, int32_t max_stamps,
                     void* stream) {
    if (!p || !T || !frame_off || !frame_ph || !frame_idx || !status || !stamps || !n_stamps || max_stamps <= 0) return BFA_E_INVALID;
    if (B <= 0) return B == 0 ? BFA_OK : BFA_E_INVALID;
    AssortArgs aa;
    aa.p = *p; aa.B = B; aa.C = 0; aa.max_stamps = max_stamps; aa.logp = nullptr; aa.row_off = nullptr; aa.T = T;
    aa.frame_off = (const long long*)frame_off; aa.frame_ph = frame_ph; aa.frame_idx = frame_idx; aa.status = status;
    aa.stamps = stamps; aa.conf = nullptr; aa.n_stamps = n_stamps; aa.path_lp = nullptr; aa.uflag = nullptr; aa.row_lse = nullptr;
    aa.ts = 0; aa.ss = assort_ss(max_stamps);     // utterance lengths are only known on the device here: frames are read in place
    assort_confidence_kernel<<<(B + ASSORT_WARPS - 1) / ASSORT_WARPS, ASSORT_WARPS * 32, assort_smem(aa.ts, aa.ss), (cudaStream_t)stream>>>(aa);
    LAUNCH_CHECK();
    return BFA_OK;
}

}  // extern "C"
