// viterbi_band3.cuh -- the hot kernel: banded stride-4 Viterbi fill + back-trace on a reduced, provably
// equivalent lattice (3 states per phoneme), 4 utterances per warp.
//
// Same DP as viterbi_generic.cuh (forced_alignment.py:563-703) specialised for the shape that dominates real
// work: the structured stride-4 path  b,(p,b1,b2,b3)xN  (forced_alignment.py:153, :181-187) with a
// Sakoe-Chiba band (:190, :441) or a path short enough to fit the window.
//
// 1. Reduced lattice.  Transitions of the reference lattice (can_skip, :599-605): p<-{p,b3',b2'},
//    b1<-{b1,p}, b2<-{b2,b1,p}, b3<-{b3,b2} (' = previous group).  All blank states emit lp[t,blank], so two
//    paths with the same (phoneme/blank) output sequence have bit-identical fp32 scores (same addends, same
//    order).  Consequences, all in exact fp32 arithmetic (max(a,b)+e == max(a+e,b+e) by monotonicity):
//      * b2 >= b1 whenever b2 is alive, so b1 never wins strictly inside b2's arg-max (first max wins, :645)
//        and b1 is never on a back-trace path while b2 is in band; b1 is only ever needed while the upper band
//        edge sits between b1 and b2, where it behaves exactly like b2.  Hence b1 and b2 merge into one state
//        m = "b1 u b2" with transitions  m<-{m,p}  (ties -> stay = blank predecessor, the reference's order),
//        alive iff b1 is below the upper edge and b2 above the lower edge.
//      * p<-{p,b3',m'}, b3<-{b3,m} keep their candidate order, so every arg-max decision that can change the
//        output is taken on the same fp32 values in the same order as the reference.
//      * (1b) b3 <= m at every frame (induction: b3' = max(b3,m)+e = m+e <= max(m,p)+e = m', rounding is monotone), so
//        b3's arg-max is always "advance from m" unless b3 == m, which holds exactly when m itself stayed one frame
//        earlier: b3's decision bit at frame t is m's decision bit at frame t-1, and b3's value is m's stay candidate.
//        b3 therefore costs nothing in the frame loop; its decision word is m's word shifted by one frame, produced
//        when a record is flushed.  (State 0, the leading blank, lives in m of group 0: same values, same outputs.)
//      * (1c) for the same reason the best way into p from the previous group is always worth m' + e; it came "from b3'"
//        (the reference's first max) exactly when b3' == m'.  One compare decides stay / enter, m's history the rest.
//    A group then costs 6 FADD + 2 FMNMX + 2 SHF per frame (2 decision bits) instead of 10 + 6 + 6 on the full lattice.
//    (1b)/(1c) read "b3 == m" off m's decision instead of comparing the sums, so they differ from the reference when
//    b3 < m but the two sums round to the same float (a half-ulp tie; ~1e-8 per state and frame).  With the fused
//    log-softmax the emissions already differ from torch's by an ulp now and then, which flips near-ties far more
//    often; when the kernel consumes the caller's log-probs unchanged (no boost: simple / raw modes) every decision of
//    the full reduced lattice (p: 2, m, b3) is taken on the sums themselves: template parameter EXACT.
//
// 2. Lazy band.  The band mask (:650-653) is not applied state by state.  The window of W = 8*G groups follows
//    the lower band edge (it slides only at the start of an 8-frame chunk) and always covers the whole band, so
//    the lattice searched here is a SUPERSET of the reference's.  If the path found stays strictly inside the
//    band at every frame (checked during the back-trace, with a safety margin), it is a path of the reference
//    lattice with a score >= every reference path, its prefix scores equal the reference's dp values, and the
//    first-max argument shows the reference back-trace returns exactly this path.  Otherwise (or when the final
//    state is invalid, an emission is > 0, a row's log-sum-exp is not finite) the item is appended to the retry
//    list and re-run by the exact generic kernel.  No per-frame band bookkeeping is left in the frame loop.
//
// 3. Fusion and roles.  Target boost + log_softmax + floor (:121-129) are applied on the fly; rows are streamed exactly
//    once from HBM by 1-D bulk async copies (TMA engine, SASS UBLKCP; 8 rows per utterance per stage, 3 stages).
//    A task (4 utterances) is run by a PAIR of warps so that twice as many warps hide each other's latencies:
//      * the helper warp owns the data movement and the transcendental work: it issues the bulk copies, and for every
//        staged chunk each of its lanes reduces ONE row (lane = (utterance, row): 66 exps, no shuffles, no idle lanes;
//        per-utterance class weights come from a small shared table) to (log-sum-exp, blank emission);
//        while a row is on chip it also keeps the raw log-prob of the frame-wise best class, the confidence pass's input
//        wherever the path agrees with that guess (path_lp / guess_cls);
//      * the DP warp runs the frame loop (registers only + one shared load per group per frame) and flushes one decision
//        record per 32 frames, stored as the cell image the back-trace walks.
//    They meet on three mbarrier rings: full (TMA -> helper), ready (helper -> DP), free (DP -> helper).
//
// 4. Back-trace, again split over the pair.  Records come back by bulk copy straight into place (4 deep); the DP warp walks one
//    RUN of the path per iteration and hands the visited cells of a 32-frame block to the helper warp (kready / kfree
//    barriers, two buffers), which writes frame_phonemes / frame_phonemes_idx, checks that the path kept B3_MARGIN states
//    away from the band edges and fetches the confidence inputs the fill guessed wrong.
//
// 5. One launch serves the three window classes (G = 3, 5, 8 groups per lane = 24, 40, 64 groups per window): every CTA works
//    through the classes in turn (band3_run), re-initialising its barriers in between.
#pragma once
#include <type_traits>

#include "assort.cuh"
#include "bfa_common.cuh"

namespace bfa {

constexpr int B3_LPU = 8;          // lanes per utterance
constexpr int B3_UPW = 4;          // utterances per warp
constexpr int B3_NST = 3;          // pipeline stages (chunk c is consumed while c+1 is being reduced and c+2 is in flight)
constexpr int B3_ROWS = 8;         // rows per stage per utterance (8*C*4 bytes is always a multiple of 16)
constexpr int B3_PAIRS = 7;        // (DP warp, helper warp) pairs per CTA, one CTA per SM
constexpr int B3_KK = 72;          // floats per utterance in the class-weight table (C <= 72)
constexpr int B3_KK2 = 2 * B3_KK;  // per utterance: the table, then a copy shifted by B3_KSH floats (odd class counts: rows that
constexpr int B3_KSH = 3;          // start on an odd float pair up classes (1,2),(3,4),..; the copy puts class 1 on a 16-byte boundary)
constexpr int B3_STP = 9;          // float2 pitch of the per-utterance (lnS, eb) array (bank spreading)
constexpr int B3_SLOT_PAD = 4;      // floats of slack per staged (utterance, chunk) slot: room for an item's lead-in (Item::lead)
constexpr int B3_NMAX = 128;       // phonemes per item on this path (byte-sized class table in shared memory)
constexpr int B3_MARGIN = 3;       // states of safety margin of the band-legality check (a frame spent in m may be reported as b3: 2 states)
constexpr int B3_NREC = 4;         // decision records of 32-frame blocks in flight from the slab during the back-trace
constexpr int B3_GRING = 4;        // 32-frame blocks of confidence gathers kept in flight during the back-trace

struct Band3Args {
    BfaParams p;
    int C;
    const float* logp;
    const int32_t* tgt;
    const uint32_t* tmask;     // [B][MAX_WORDS]
    Item* retry_items;         // generic list: items this kernel could not finish are appended
    int* n_retry;
    int32_t* frame_ph;
    int32_t* frame_idx;
    float* dp_final;
    float* path_lp;            // [total_frames] raw log-prob of the assigned class per frame (confidence input), or null
    unsigned char* guess_cls;  // [total_frames] class whose raw log-prob the fill left in path_lp (frame-wise best class), or null
    struct Class {             // one per window class (24 / 40 / 64 groups)
        const Item* items;
        const int* n_items;
        const int* n_frames;   // frames of the class's items (the planner's sum; list mode with several classes: the CTA shares)
        uint32_t* bp_scratch;  // decision-record slabs, one per resident pair
        long long bp_slab_words;
        int smem_per_warp;     // bytes of shared memory per (DP, helper) pair
        int npairs;            // pairs per CTA that fit
        int region;            // bytes of the pair's staging region (stage ring / back-trace staging), see band3_off_*
    } cls[3];
    // ---- direct mode (viterbi_band3_direct_kernel): the kernel plans its utterances itself (task q = utterances 4q .. 4q+3),
    //      emits timestamps and confidences from its own back-trace, and needs no item list.  cls[0] describes its window class.
    int B;
    const long long* row_off;
    const int32_t* T;
    const long long* tgt_off;
    const long long* frame_off;
    int32_t* status;
    BfaStamp* stamps;          // may be null (frames only)
    float* conf;               // may be null
    int32_t* n_stamps;
    int max_stamps;
    int32_t* uflag;            // [B] 1: finished here, stamps included; 0: left to the planner chain.  Null in direct-only mode
    int* deferred;             // utterances left to the planner chain (null in direct-only mode: status = BFA_ST_DEFERRED instead)
    int* n_deferred;
    float* pscr_lp;            // per-SM, per-pair scratch [slot][pair][UPW][tpitch]: raw log-prob of the frame-wise best class ...
    unsigned char* pscr_gs;    // ... and that class (the confidence inputs wherever the path agrees with the guess)
    int tpitch;                // frames per utterance in the scratch / shared staging arrays (multiple of 16, >= max_T)
    int ncap;                  // phonemes per utterance the shared event / stamp arrays hold (>= max_N of the batch)
    int nslots;                // SM slots the per-SM scratch and slabs were sized for (> every %smid)
    int direct_only;           // no planner chain follows: utterances that do not qualify get status BFA_ST_DEFERRED
    float* row_lse;            // logits mode (LOGITS instantiations): [total_frames] log(sum(exp(row))) of every row of the utterances run here
};

template <int G>
struct Band3Shape {
    static constexpr int W = B3_LPU * G;    // groups in the window
    static constexpr int ACC = 4 * G;       // decision accumulators per lane
    static constexpr int REC = 6 * G + 1;   // words per lane of a record: its 3G cells as (w0, w1) pairs + the window-slide word
    static constexpr int CELLS = 3 * W;     // back-trace cells per utterance
};

// Smallest window (in groups) that covers band(t-1) U band(t) for every frame of every 8-frame chunk when the
// window base is the group of (a conservative fp32 estimate, at most one state low, of) the lower band edge at
// the frame before the chunk (scripts/check_window_fit.py verifies the closed form by brute force).
__host__ __device__ inline int band3_window_need(int N, int T, int L, int band) {
    if (band <= 0 || T < 2 || L < 2) return N + 1;
    const int adv = (L - 1 < (1 << 27))                                           // ceil(8 * pace), 32-bit whenever it fits
                        ? (int)(((unsigned)B3_ROWS * (unsigned)(L - 1) + (unsigned)(T - 2)) / (unsigned)(T - 1))
                        : (int)(((long long)B3_ROWS * (L - 1) + (T - 2)) / (T - 1));
    const int w = (2 * band + adv + 6) / 4 + 1;
    return w < N + 1 ? w : N + 1;
}

// The stage ring doubles as the back-trace staging area (cells + visited-cell buffer) once the fill is done.
__host__ __device__ inline size_t band3_stage_region(int C, int G) {
    size_t st = (size_t)B3_NST * B3_UPW * (B3_ROWS * C + B3_SLOT_PAD) * 4;
    // back-trace: B3_NREC record buffers (a record IS the cell array of its block: [UPW][3*8G] x 8 B, + 32 slide words), two
    // visited-cell buffers [4][32] words, B3_GRING gather blocks [4][32] floats, per-utterance verdict + final score
    size_t bt = (size_t)B3_NREC * (6 * G + 1) * 128 + (size_t)2 * 4 * 128 + (size_t)B3_GRING * 4 * 128 + (size_t)B3_UPW * 8;
    size_t b = st > bt ? st : bt;
    return (b + 15) / 16 * 16;
}
// word offset of the visited-cell buffers inside the back-trace staging area (after the record buffers)
__host__ __device__ inline int band3_bt_keep_words(int G) { return B3_NREC * (6 * G + 1) * 32; }
// shared memory of one pair: staging region (stage ring, later the back-trace staging) | class weights | (lnS, eb) per staged row |
// target classes | helper flags + utterance states | the pair's work items (direct mode) | mbarriers.  R = bytes of the region.
enum : int { B3_BAR_FULL = 0, B3_BAR_READY = B3_NST, B3_BAR_FREE = 2 * B3_NST, B3_BAR_REC = 3 * B3_NST, B3_BAR_KREADY = 3 * B3_NST + B3_NREC,
             B3_BAR_KFREE = 3 * B3_NST + B3_NREC + 2, B3_BAR_PLAN = 3 * B3_NST + B3_NREC + 4, B3_BAR_PAIR = 3 * B3_NST + B3_NREC + 5, B3_NBARS = 3 * B3_NST + B3_NREC + 6 };
// arrivals that complete a phase of barrier i: every utterance's copy on the stage-full barriers, both warps on the pair rendezvous, one elsewhere
__host__ __device__ inline uint32_t band3_bar_count(int i) { return (i >= B3_BAR_FULL && i < B3_BAR_FULL + B3_NST) ? B3_UPW : (i == B3_BAR_PAIR ? 2u : 1u); }
__host__ __device__ inline size_t band3_off_kk(size_t R) { return R; }
__host__ __device__ inline size_t band3_off_stats(size_t R) { return band3_off_kk(R) + (size_t)B3_UPW * B3_KK2 * 4; }
__host__ __device__ inline size_t band3_off_cls(size_t R) { return band3_off_stats(R) + (size_t)B3_NST * B3_UPW * B3_STP * 8; }
__host__ __device__ inline size_t band3_off_flags(size_t R) { return band3_off_cls(R) + (size_t)B3_UPW * B3_NMAX; }
__host__ __device__ inline size_t band3_off_items(size_t R) { return band3_off_flags(R) + 32; }
__host__ __device__ inline size_t band3_off_bars(size_t R) { return band3_off_items(R) + (size_t)B3_UPW * sizeof(Item); }
__host__ __device__ inline size_t band3_smem_per_pair(size_t R) { return (band3_off_bars(R) + (size_t)B3_NBARS * 8 + 127) / 128 * 128; }
__host__ __device__ inline size_t band3_smem_per_warp(int C, int G) { return band3_smem_per_pair(band3_stage_region(C, G)); }   // per PAIR of warps

// Direct mode: what the back-trace keeps in the staging region (G = 3).  Byte offsets:
//   events [UPW][evcap] | (verdict, final score) [UPW] + event counts [UPW] | (start, end) per phoneme [UPW][ncap2] |
//   per warp of the pair: run-boundary bits (tpitch / 32 + 1 words) + fetched-frame nibbles (tpitch / 4 bytes) | record buffers of the walk |
//   per utterance: lp_s [tpitch] floats + gs_s [tpitch] bytes (confidence inputs kept by the fill)
__host__ __device__ inline int band3_evcap(int ncap) { return (3 * ncap + 8 + 3) & ~3; }
__host__ __device__ inline size_t band3_d_off_ev() { return 0; }
__host__ __device__ inline size_t band3_d_off_fin(int ncap) { return (size_t)B3_UPW * band3_evcap(ncap) * 4; }
__host__ __device__ inline size_t band3_d_off_se(int ncap) { return band3_d_off_fin(ncap) + 64; }
__host__ __device__ inline size_t band3_d_off_bits(int ncap) { return band3_d_off_se(ncap) + (size_t)B3_UPW * ((ncap + 1) & ~1) * 8; }
__host__ __device__ inline size_t band3_d_off_rec(int ncap, int tpitch) { return (band3_d_off_bits(ncap) + (size_t)8 * (tpitch / 32 + 1) * 4 + 127) / 128 * 128; }
__host__ __device__ inline size_t band3_d_off_lp(int ncap, int tpitch) {
    return band3_d_off_rec(ncap, tpitch) + (size_t)B3_NREC * (6 * 3 + 1) * 128;
}
__host__ __device__ inline size_t band3_direct_region(int C, int tpitch, int ncap) {
    const size_t st = band3_stage_region(C, 3), bt = band3_d_off_lp(ncap, tpitch) + (size_t)B3_UPW * 5 * tpitch;
    return ((st > bt ? st : bt) + 15) / 16 * 16;
}

__device__ __forceinline__ float b3_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float b3_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// packed fp32 pairs (sm_100: FADD2 / FFMA2, one issue slot for two IEEE operations)
__device__ __forceinline__ unsigned long long b3_pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void b3_unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long b3_add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long b3_fma2(unsigned long long a, float s, unsigned long long c) {   // a * (s, s) + c
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b3_pack2(s, s)), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long b3_fma2v(unsigned long long a, unsigned long long b, unsigned long long c) {   // a * b + c, lane by lane
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// 4-byte asynchronous global->shared copies (LDGSTS): completion is tracked by cp.async groups, not by the register
// scoreboard, so nothing that follows stalls on them until the matching wait_group
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// push (later > earlier) into acc: the sign bit of (earlier - later) is 1 iff later > earlier (ties -> 0)
__device__ __forceinline__ void b3_push(uint32_t& acc, float earlier, float later) {
    acc = __funnelshift_l(__float_as_uint(earlier - later), acc, 1);
}

// Optional phase timers (development only, -DBFA_PHASE_PROF): warp-clock cycles per phase, summed over lane 0 of the DP warps.
#ifdef BFA_PHASE_PROF
__device__ unsigned long long g_b3_phase[32];
__device__ unsigned long long g_b3_warp[32];
__device__ unsigned long long g_b3_cta[160 * 2];   // [cta]: max DP task cycles, [160 + cta]: physical SM id   // [warp id]: summed task cycles, [16 + warp id]: tasks
__device__ unsigned long long g_b3_fin[32];   // band3_direct_finish sub-phases: [which * 16 + i]
#define FIN_DECL long long fin_last = clock64(); long long fin_acc[10] = {0,0,0,0,0,0,0,0,0,0}
#define FIN_T(i) do { const long long fin_now = clock64(); fin_acc[i] += fin_now - fin_last; fin_last = fin_now; } while (0)
#define FIN_FLUSH do { if (lane == 0) { for (int i = 0; i < 10; ++i) atomicAdd(&g_b3_fin[which * 16 + i], (unsigned long long)fin_acc[i]); } } while (0)
#define PH_DECL long long ph_last = clock64(); long long ph_acc[16] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0}; const int ph_base = 0
#define PH_DECL_H long long ph_last = clock64(); long long ph_acc[16] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0}; const int ph_base = 16
#define PH_RESET ph_last = clock64()
#define PH_T(i) do { const long long ph_now = clock64(); ph_acc[i] += ph_now - ph_last; ph_last = ph_now; } while (0)
#define PH_FLUSH do { if (lane == 0) { long long ph_tot = 0; for (int i = 0; i < 14; ++i) { ph_tot += ph_acc[i]; atomicAdd(&g_b3_phase[ph_base + i], (unsigned long long)ph_acc[i]); } \
    atomicMax(&g_b3_phase[ph_base + 14], (unsigned long long)ph_tot); atomicMin(&g_b3_phase[ph_base + 15], (unsigned long long)ph_tot); \
    atomicAdd(&g_b3_warp[threadIdx.x >> 5], (unsigned long long)ph_tot); atomicAdd(&g_b3_warp[16 + (threadIdx.x >> 5)], 1ull); \
    if (ph_base == 0 && blockIdx.x < 160) { unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); atomicMax(&g_b3_cta[blockIdx.x], (unsigned long long)ph_tot); g_b3_cta[160 + blockIdx.x] = smid; } } } while (0)
#else
#define FIN_DECL
#define FIN_T(i)
#define FIN_FLUSH
#define PH_DECL
#define PH_DECL_H
#define PH_RESET
#define PH_T(i)
#define PH_FLUSH
#endif

// What both warps of a pair derive from the item list (uniform within an 8-lane segment).  `my_item` is this lane's segment's
// item: an entry of the global list (list mode) or the copy the helper warp planned into shared memory (direct mode); `on` says
// whether the segment has work.
template <int G, int CT>
struct Band3Task {
    int seg, l8, C;
    bool seg_on;
    const Item* it;
    int T, Tmax, n_chunks, seg_stride, stage_floats, lead;
    float* stage_buf;
    float* kk;
    float2* stats;
    unsigned char* cls8;
    int* hflags;
    uint32_t bar0;
    const float* my_src;
    bool use_stats, warp_stats;
    __device__ __forceinline__ Band3Task(const Band3Args& a, const Item* my_item, bool on, unsigned char* smem_pair, size_t R, int lane) {
        seg = lane >> 3; l8 = lane & 7;
        C = CT ? CT : a.C;
        seg_on = on;
        it = my_item;
        T = seg_on ? it->T : 0;
        Tmax = T;
#pragma unroll
        for (int d = 8; d < 32; d <<= 1) Tmax = max(Tmax, __shfl_xor_sync(FULL, Tmax, d));
        n_chunks = (Tmax + B3_ROWS - 1) / B3_ROWS;
        seg_stride = B3_ROWS * C + B3_SLOT_PAD;
        lead = seg_on ? it->lead : 0;
        stage_floats = B3_UPW * seg_stride;
        stage_buf = reinterpret_cast<float*>(smem_pair);
        kk = reinterpret_cast<float*>(smem_pair + band3_off_kk(R));                   // [UPW][B3_KK]
        stats = reinterpret_cast<float2*>(smem_pair + band3_off_stats(R));            // [NST][UPW][B3_STP]
        cls8 = smem_pair + band3_off_cls(R);                                          // [UPW][B3_NMAX]
        hflags = reinterpret_cast<int*>(smem_pair + band3_off_flags(R));              // [UPW] helper verdict: rows not sane; then [UPW] utterance states (direct mode)
        bar0 = smem_u32(smem_pair + band3_off_bars(R));
        my_src = a.logp + it->lp_off;
        use_stats = seg_on && (it->flags & ITEM_STATS) != 0;
        warp_stats = __any_sync(FULL, use_stats);   // every item of a call shares the mode
    }
};
// the segment's item and whether it is on: list mode (items of the window class) / direct mode (planned into shared memory)
__device__ __forceinline__ const Item* band3_list_item(const Item* items, int first, int n_valid, int lane, bool& on) {
    const int seg = lane >> 3;
    on = seg < n_valid;
    return &items[first + (on ? seg : 0)];
}
__device__ __forceinline__ const Item* band3_direct_item(unsigned char* smem_pair, size_t R, int lane, bool& on) {
    const Item* it = reinterpret_cast<const Item*>(smem_pair + band3_off_items(R)) + (lane >> 3);
    on = it->T > 0;
    return it;
}
// utterance states of a direct-mode task (shared memory, after the helper flags)
enum : int { B3_U_RUN = 0, B3_U_NONE = 1, B3_U_DEFER = 2 };

// Direct mode: the helper warp plans the four utterances of task `task` (what plan_kernel's single-item branch does,
// forced_alignment.py:153-190 / :963-976) and leaves their items in shared memory.  An utterance is taken here when it is a single
// stride-4 DP problem that fits the 24-group window and its targets are plain phoneme classes; when silence anchoring is on and
// the target holds silence_id (a segmentation attempt, :133), or anything else is unusual, it is left to the planner chain.
template <int CT>
__device__ __noinline__ void band3_direct_plan(const Band3Args& a, int task, unsigned char* smem_pair, size_t R, int lane) {
    constexpr float LOG2E = 1.4426950408889634f;
    const int seg = lane >> 3, l8 = lane & 7;
    const int C = CT ? CT : a.C;
    const BfaParams& p = a.p;
    Item* items = reinterpret_cast<Item*>(smem_pair + band3_off_items(R));
    int* ustate = reinterpret_cast<int*>(smem_pair + band3_off_flags(R)) + B3_UPW;
    unsigned char* cls8 = smem_pair + band3_off_cls(R) + seg * B3_NMAX;
    float* kk = reinterpret_cast<float*>(smem_pair + band3_off_kk(R)) + seg * B3_KK2;
    const int u = task * B3_UPW + seg;
#ifdef BFA_PHASE_PROF
    long long pp_last = clock64();
#define PP_T(i) do { const long long pp_now = clock64(); if (lane == 0) atomicAdd(&g_b3_phase[16 + (i)], (unsigned long long)(pp_now - pp_last)); pp_last = pp_now; } while (0)
#else
#define PP_T(i)
#endif
    int state = B3_U_NONE;
    Item it;
    it.lp_off = 0; it.stat_off = 0; it.out_off = 0; it.out_lim = 0; it.seq_off = 0;
    it.T = 0; it.L = 0; it.band = 0; it.stride = 4; it.n = 0; it.idx0 = 0; it.trim = 0; it.n_out = 0;
    it.utt = u; it.anchor_off = 0; it.flags = 0; it.lead = 0;
    int T = 0, N = 0;
    const bool boost = p.boost_targets && p.mode == BFA_MODE_FULL;
    if (u < a.B) {
        state = B3_U_DEFER;
        T = a.T[u];
        const long long t0 = a.tgt_off[u], t1 = a.tgt_off[u + 1];
        const long long ro = a.row_off[u], fo0 = a.frame_off[u], fo1 = a.frame_off[u + 1];
        N = (int)min(t1 - t0, (long long)(1 << 20));
        const int L = 4 * N + 1;
        bool ok = N >= 1 && N <= B3_NMAX && N <= a.ncap && T >= 2 && T <= a.tpitch;
        if (p.mode == BFA_MODE_SIMPLE) ok = ok && !((double)L > (double)T * 0.9) && !((double)L > (double)T * 0.8);   // :963-968: stride stays 4 (both tests see stride 4)
        else ok = ok && L <= T;                                                          // :153-157
        const int band = (L > 60) ? max(L / 4, 20) : 0;                                   // :190 / :976
        ok = ok && L <= T && band3_window_need(N, T, L, band) <= B3_LPU * 3;
        const int lead = (int)(((unsigned long long)(a.logp + ro) & 15ull) >> 2);
        if (lead != 0 && (ro < lead || (C == 66 && (lead & 1)))) ok = false;     // same rule as the planner's fast_class (C = 66 reads 8-byte pairs from an even float)
        it.lp_off = ro; it.stat_off = fo0; it.out_off = fo0; it.out_lim = fo1; it.seq_off = t0;
        it.L = L; it.band = band; it.n = N; it.n_out = T; it.lead = lead;
        it.flags = ITEM_FINAL;
        if (p.mode == BFA_MODE_FULL) it.flags |= (p.boost_targets ? ITEM_STATS : 0) | (p.enforce_minimum ? ITEM_FLOOR : 0);
        if (ok) state = B3_U_RUN;
    }
    PP_T(10);
    // targets: byte-sized class table + the class weights of the fused log-sum-exp, exp(x + boost*[c in targets] - boost) =
    // 2^(x*log2e + kk[c]); every id must be a plain class, and none may be silence_id while silence anchoring is on
    if (boost) {
        for (int c = l8; c < B3_KK; c += B3_LPU) {
            kk[c] = c < C ? -p.boost_factor * LOG2E : -INFINITY;
            if (c + B3_KSH < B3_KK) kk[B3_KK + B3_KSH + c] = kk[c];
        }
    }
    __syncwarp();
    PP_T(11);
    bool tbad = false;
    if (state == B3_U_RUN) {
        const bool segmenting = p.mode == BFA_MODE_FULL && p.silence_anchors > 0 && p.silence_id >= 0;
        const int32_t* seq = a.tgt + it.seq_off;
        int cv[B3_NMAX / B3_LPU];                  // all of this lane's targets in flight at once (one memory round trip)
#pragma unroll
        for (int i = 0; i < B3_NMAX / B3_LPU; ++i) {
            const int j = l8 + B3_LPU * i;
            cv[i] = j < N ? seq[j] : 0;
        }
#pragma unroll
        for (int i = 0; i < B3_NMAX / B3_LPU; ++i) {
            const int j = l8 + B3_LPU * i;
            if (j < N) {
                const int c = cv[i];
                if (c < 0 || c >= C || c == p.blank_id || (segmenting && c == p.silence_id)) tbad = true;
                else {
                    cls8[j] = (unsigned char)c;
                    if (boost) {
                        kk[c] = 0.0f;
                        if (c + B3_KSH < B3_KK) kk[B3_KK + B3_KSH + c] = 0.0f;
                    }
                }
            }
        }
    }
    const unsigned bm = __ballot_sync(FULL, tbad);
    PP_T(12);
    if ((bm >> (seg * B3_LPU)) & 0xffu) state = B3_U_DEFER;
    if (l8 == 0) {
        it.T = state == B3_U_RUN ? T : 0;
        items[seg] = it;
        ustate[seg] = state;
    }
    __syncwarp();
    PP_T(13);
#undef PP_T
}

// ------------------------------------------------------------------------------------------------------------
// Helper warp: tables, bulk copies, row statistics.
// ------------------------------------------------------------------------------------------------------------
// both warps of a pair meet here (named barrier 1 + pair, 64 threads): hand-overs of the direct mode
// A rendezvous on the pair's own mbarrier (two arrivals per phase, release / acquire like the other hand-overs): unlike a named
// hardware barrier it involves nobody but the two warps, whatever the rest of the CTA is doing or has already left.
__device__ __forceinline__ void band3_pair_sync(uint32_t bar0, uint32_t& phase, int lane) {
    const uint32_t bar = bar0 + 8u * B3_BAR_PAIR;
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
    mbar_wait(bar, (phase >> B3_BAR_PAIR) & 1u);
    phase ^= 1u << B3_BAR_PAIR;
    __syncwarp();
}

template <int CT, bool LOGITS>
__device__ __noinline__ void band3_direct_finish(const Band3Args& a, unsigned char* smem_pair, size_t R, int seg0, int which, int lane, bool have_spec);

// LOGITS (direct mode only): the rows hold un-normalised logits.  Boosting re-normalises every row (:51-54), so emissions, path,
// timestamps and DP score are the same whatever constant a row is shifted by; only the confidences (utils.py:81, exp of the
// ORIGINAL log-probabilities) need the row's own log-sum-exp.  The reduction carries it along: with e_c = 2^(x_c log2e + kk_c) the
// boosted sum is S = sum(e_c) and the plain one sum(e_c * 2^(-kk_c)), and 2^(-kk_c) is 1 for a target class and exp(boost) otherwise,
// i.e. 1 + kk_c * MSC: the plain sum is S + MSC * sum(e_c * kk_c) -- one more packed FMA per pair of classes.  It is written out per
// frame (Band3Args::row_lse).
template <int G, int CT, bool DIRECT, bool LOGITS = false>
__device__ void band3_helper(const Band3Args& a, const Band3Args::Class& kc, int first, int n_valid, unsigned char* smem_pair, uint32_t& phase, bool not_first,
                             int lane, uint64_t pol, int pair, float* pscr_lp, unsigned char* pscr_gs) {
    constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    const size_t R = (size_t)kc.region;
    // This warp finished the previous task last (it wrote that task's outputs; direct mode: both warps met at the pair barrier),
    // so tables and staging area are free; the area was last written through the generic proxy and is about to be written by
    // bulk copies.
    PH_DECL_H;
    if (not_first) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (DIRECT) band3_direct_plan<CT>(a, first, smem_pair, R, lane);      // `first` is the task number
    PH_T(8);
    bool on;
    const Item* my_item = DIRECT ? band3_direct_item(smem_pair, R, lane, on) : band3_list_item(kc.items, first, n_valid, lane, on);
    const Band3Task<G, CT> k(a, my_item, on, smem_pair, R, lane);
    const int seg = k.seg, l8 = k.l8, C = k.C, T = k.T;
    const int blank = a.p.blank_id;
    const float boostv = a.p.boost_factor;
    const bool warp_stats = k.warp_stats;
    static_assert(!LOGITS || DIRECT, "logits are taken by the direct kernel only");
    const float MSC = (LOGITS && boostv != 0.0f) ? (1.0f - expf(boostv)) / (boostv * LOG2E) : 0.0f;   // 2^(-kk) = 1 + kk * MSC for kk in {0, -boost log2e}

    // ---- per-utterance tables (built before the first bulk copies are issued: behind them these small loads would queue
    //      for microseconds): target classes (bytes) and the class weights of the fused log-sum-exp:
    //      exp(x + boost*[c in targets] - boost) = 2^(x*log2e + kk[c]) ----   (direct mode: band3_direct_plan built them)
    if (!DIRECT) {
        if (k.seg_on) {
            const int32_t* seq = a.tgt + k.it->seq_off;
            const int N = k.it->n;
            for (int j = l8; j < N; j += B3_LPU) k.cls8[seg * B3_NMAX + j] = (unsigned char)seq[j];
        }
        if (warp_stats) {
            const int utt = k.it->utt;
            for (int c = l8; c < B3_KK; c += B3_LPU) {
                const bool ok = c < C;
                const bool tg = ok && k.seg_on && ((a.tmask[(size_t)utt * MAX_WORDS + (c >> 5)] >> (c & 31)) & 1u);
                const float kv = ok ? (tg ? 0.0f : -boostv * LOG2E) : -INFINITY;
                k.kk[seg * B3_KK2 + c] = kv;
                if (c + B3_KSH < B3_KK) k.kk[seg * B3_KK2 + B3_KK + B3_KSH + c] = kv;
            }
        }
    }
    __syncwarp();
    if (DIRECT && lane == 0) mbar_arrive(k.bar0 + 8u * B3_BAR_PLAN);      // the DP warp may read the items and the class table
    PH_T(9);
    // lanes with l8 == 0 issue their own utterance's copy and arrive once per chunk on the stage's full barrier (count = UPW)
    const uint32_t dst0 = smem_u32(k.stage_buf + seg * k.seg_stride);
    const uint32_t full_bytes = (uint32_t)B3_ROWS * C * 4;
    const uint32_t lead_b = 4u * (uint32_t)k.lead;      // the copy starts this many bytes before the chunk (16-byte aligned source)
    auto issue = [&](int c, int st) {
        if (l8 == 0) {
            const int rows = T - c * B3_ROWS;
            const uint32_t bar = k.bar0 + 8u * (B3_BAR_FULL + st);
            const uint32_t dst = dst0 + (uint32_t)st * k.stage_floats * 4u;
            const float* s = k.my_src + (size_t)c * B3_ROWS * C - k.lead;
            if (rows > B3_ROWS || (rows == B3_ROWS && lead_b == 0)) {
                const uint32_t nb = (full_bytes + lead_b + 15u) & ~15u;   // may take in the first floats of the next chunk
                mbar_expect_tx(bar, nb);
                bulk_g2s_hint(dst, s, nb, bar, pol);
            } else if (rows > 0) {                                        // last chunk: never read past the item's last row
                const uint32_t bytes = (uint32_t)rows * C * 4 + lead_b, bulk = bytes & ~15u;
                float* d = k.stage_buf + st * k.stage_floats + seg * k.seg_stride;
                for (uint32_t w = bulk >> 2; w < (bytes >> 2); ++w) d[w] = s[w];   // < 4 tail floats
                mbar_expect_tx(bar, bulk);
                if (bulk) bulk_g2s_hint(dst, s, bulk, bar, pol);
            } else {
                mbar_arrive(bar);
            }
        }
    };
#pragma unroll
    for (int c = 0; c < B3_NST; ++c)
        if (c < k.n_chunks) issue(c, c);


    // Speculative confidence inputs: the confidence pass needs lp[f, class of the path at f] (utils.py:89-103), which is
    // only known after the back-trace, when the row has long left the chip.  While a row is in shared memory this warp
    // stores the raw log-prob of its frame-wise best (boosted) class together with that class; the back-trace then only
    // has to fetch the frames where the path disagrees with the guess.
    const bool spec = DIRECT ? (warp_stats && pscr_lp != nullptr) : (warp_stats && a.path_lp != nullptr && a.guess_cls != nullptr);
    const int o_trim = k.it->trim;
    const long long o_out = k.it->out_off;
    int rel_hi = 0;                  // frames [trim, trim + rel_hi) of the item are written out (:447-448, :465-467)
    if (k.seg_on) {
        const long long room = k.it->out_lim - o_out;
        rel_hi = (int)min((long long)k.it->n_out, max(room, 0LL));
        rel_hi = min(rel_hi, T - o_trim);
    }
    // indexed by the item's frame number; direct mode: the pair's private scratch, read back by this task's own back-trace
    float* const plp_row = DIRECT ? (pscr_lp ? pscr_lp + seg * a.tpitch : nullptr) : (a.path_lp ? a.path_lp + (o_out - o_trim) : nullptr);
    unsigned char* const gcl_row = DIRECT ? (pscr_gs ? pscr_gs + seg * a.tpitch : nullptr) : (a.guess_cls ? a.guess_cls + (o_out - o_trim) : nullptr);

    float emax = -INFINITY;          // raw mode: emissions must be <= 0 (log-probabilities)
    float lse_chk = 0.f;             // running sum of the rows' log-sum-exp (finite <=> all rows sane)
    int st = 0;
    for (int c = 0; c < k.n_chunks; ++c) {
        mbar_wait(k.bar0 + 8u * (B3_BAR_FULL + st), (phase >> (B3_BAR_FULL + st)) & 1u);
        phase ^= 1u << (B3_BAR_FULL + st);
        __syncwarp();                                  // tail floats written by the issuing lane become visible
        PH_T(5);

        // ---- row statistics: lane (seg, l8) owns row l8 of its utterance: log-sum-exp of the boosted row (:51-54)
        //      and the blank emission; no cross-lane traffic ----
        {
            const float* rowp = k.stage_buf + st * k.stage_floats + seg * k.seg_stride + k.lead + l8 * C;
            const int t_row = c * B3_ROWS + l8;
            float lnS = 0.f;
            float lnS0 = 0.f;                        // logits mode: log-sum-exp of the raw row
            if (warp_stats) {
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                float z0 = 0.f, z1 = 0.f, z2 = 0.f, z3 = 0.f;      // logits mode: the plain sums
                float best = -INFINITY;                  // running max of the boosted values, class index in the low 7 mantissa bits
                // the mask must live in a register for (v & mask) | class to be ONE three-input logic op (the class is the
                // immediate); deriving it from a run-time value keeps ptxas from folding it back into a second immediate
                const uint32_t tagmask = 0xffffff80u | ((uint32_t)a.C >> 16);
                auto tag = [&](float v, int c) { return __uint_as_float((__float_as_uint(v) & tagmask) | (uint32_t)c); };
                const float* kp = k.kk + seg * B3_KK2;
                auto term = [&](float x, float kc, int c, float& acc, float& zacc) {
                    const float v = fmaf(x, LOG2E, kc);
                    const float e = b3_ex2(v);
                    acc += e;
                    if constexpr (LOGITS) zacc = fmaf(e, kc, zacc);
                    best = fmaxf(best, tag(v, c));
                };
                if (CT != 0 && (CT & 1) == 1) {
                    // Odd compiled class count (67 = the reference's phoneme head, 17 = its group head): a row of CT floats starts
                    // on an even or an odd float, alternately.  Classes are still taken two per instruction: a row on an even
                    // float pairs (0,1),(2,3),.. and leaves class CT-1 single, a row on an odd float leaves class 0 single and
                    // pairs (1,2),(3,4),.. (aligned again), with the class weights read from the shifted copy of the table.
                    // Pair elements carry the tag "position in the pair sequence" (class = tag + first paired class), the
                    // single class the tag 127.
                    const int odd = (k.lead + l8) & 1;         // parity of the row's first float (stage slot and seg pitch are even)
                    const unsigned long long* x64 = reinterpret_cast<const unsigned long long*>(rowp + odd);
                    const float* kq = odd ? kp + B3_KK + B3_KSH + 1 : kp;                  // weights of the first paired class, 16-byte aligned
                    const ulonglong2* k128 = reinterpret_cast<const ulonglong2*>(kq);
                    unsigned long long sA = 0ull, sB = 0ull, zA = 0ull, zB = 0ull;
                    auto pair_term = [&](unsigned long long x, unsigned long long kc, int c, unsigned long long& acc2, unsigned long long& zacc2) {
                        const unsigned long long v = b3_fma2(x, LOG2E, kc);
                        float v0, v1;
                        b3_unpack2(v, v0, v1);
                        const unsigned long long e2 = b3_pack2(b3_ex2(v0), b3_ex2(v1));
                        acc2 = b3_add2(acc2, e2);
                        if constexpr (LOGITS) zacc2 = b3_fma2v(e2, kc, zacc2);
                        best = fmaxf(best, fmaxf(tag(v0, c), tag(v1, c + 1)));
                    };
                    constexpr int NP = (CT - 1) / 2;           // pairs
#pragma unroll
                    for (int i = 0; i < NP / 2; ++i) {
                        const ulonglong2 kv = k128[i];
                        pair_term(x64[2 * i], kv.x, 4 * i, sA, zA);
                        pair_term(x64[2 * i + 1], kv.y, 4 * i + 2, sB, zB);
                    }
                    if (NP & 1) pair_term(x64[NP - 1], reinterpret_cast<const unsigned long long*>(kq)[NP - 1], 2 * (NP - 1), sA, zA);
                    b3_unpack2(sA, s0, s1);
                    b3_unpack2(sB, s2, s3);
                    if constexpr (LOGITS) { b3_unpack2(zA, z0, z1); b3_unpack2(zB, z2, z3); }
                    {   // the single class: CT-1 (row on an even float) or 0 (row on an odd float)
                        const int cs1 = odd ? 0 : CT - 1;
                        const float v = fmaf(rowp[cs1], LOG2E, kp[cs1]);
                        const float e = b3_ex2(v);
                        s0 += e;
                        if constexpr (LOGITS) z0 = fmaf(e, kp[cs1], z0);
                        best = fmaxf(best, tag(v, 127));
                    }
                    // tag -> class
                    {
                        const int tg = (int)(__float_as_uint(best) & 127u);
                        const int cls_best = tg == 127 ? (odd ? 0 : CT - 1) : tg + odd;
                        best = __uint_as_float((__float_as_uint(best) & ~127u) | (uint32_t)cls_best);
                    }
                } else if (CT != 0 && (CT & 1) == 0) {
                    // two classes per instruction where the ISA allows it (FFMA2 / FADD2, same IEEE results lane by lane):
                    // the issue slots, not the FP pipe, are what this warp competes for
                    const unsigned long long* x64 = reinterpret_cast<const unsigned long long*>(rowp);
                    const ulonglong2* k128 = reinterpret_cast<const ulonglong2*>(kp);
                    unsigned long long sA = 0ull, sB = 0ull, zA = 0ull, zB = 0ull;          // (s0, s1), (s2, s3)
                    auto pair_term = [&](unsigned long long x, unsigned long long kc, int c, unsigned long long& acc2, unsigned long long& zacc2) {
                        const unsigned long long v = b3_fma2(x, LOG2E, kc);
                        float v0, v1;
                        b3_unpack2(v, v0, v1);
                        const unsigned long long e2 = b3_pack2(b3_ex2(v0), b3_ex2(v1));
                        acc2 = b3_add2(acc2, e2);
                        if constexpr (LOGITS) zacc2 = b3_fma2v(e2, kc, zacc2);
                        best = fmaxf(best, fmaxf(tag(v0, c), tag(v1, c + 1)));
                    };
#pragma unroll
                    for (int i = 0; i < CT / 4; ++i) {
                        const ulonglong2 kv = k128[i];
                        pair_term(x64[2 * i], kv.x, 4 * i, sA, zA);
                        pair_term(x64[2 * i + 1], kv.y, 4 * i + 2, sB, zB);
                    }
                    if (CT % 4) pair_term(x64[CT / 2 - 1], reinterpret_cast<const unsigned long long*>(kp)[CT / 2 - 1], CT - 2, sA, zA);
                    b3_unpack2(sA, s0, s1);
                    b3_unpack2(sB, s2, s3);
                    if constexpr (LOGITS) { b3_unpack2(zA, z0, z1); b3_unpack2(zB, z2, z3); }
                } else {
                    int i = 0;
                    for (; i + 4 <= C; i += 4) {
                        term(rowp[i], kp[i], i, s0, z0);
                        term(rowp[i + 1], kp[i + 1], i + 1, s1, z1);
                        term(rowp[i + 2], kp[i + 2], i + 2, s2, z2);
                        term(rowp[i + 3], kp[i + 3], i + 3, s3, z3);
                    }
                    for (; i < C; ++i) term(rowp[i], kp[i], i, s0, z0);
                }
                lnS = b3_lg2((s0 + s1) + (s2 + s3)) * LN2;          // log sum exp(x + b - boost)
                if (k.use_stats && t_row < T) lse_chk += lnS;       // any zero / overflowing / NaN sum leaves a non-finite trace
                if constexpr (LOGITS) {
                    // sum(e_c 2^(-kk_c)) = sum(e_c (1 + kk_c MSC)) = S + MSC sum(e_c kk_c): ONE more packed FMA per pair of classes
                    lnS0 = b3_lg2(fmaf(MSC, (z0 + z1) + (z2 + z3), (s0 + s1) + (s2 + s3))) * LN2;     // log sum exp(x)
                    const int rel = t_row - o_trim;
                    if (k.use_stats && t_row < T) {
                        lse_chk += lnS0;
                        if (rel >= 0 && rel < rel_hi) a.row_lse[o_out - o_trim + t_row] = lnS0;
                    }
                }
                if (spec) {
                    const int cs = min((int)(__float_as_uint(best) & 127u), C - 1);
                    const int rel = t_row - o_trim;
                    if (rel >= 0 && rel < rel_hi) {
                        plp_row[t_row] = rowp[cs] - lnS0;            // (logits mode: the class's log-probability)
                        gcl_row[t_row] = (unsigned char)cs;
                    }
                }
            } else {
                float m = -INFINITY;
                for (int i = 0; i < C; ++i) m = fmaxf(m, rowp[i]);
                if (t_row < T) emax = fmaxf(emax, m);
            }
            // blank is never a target: x - lse with lse = lnS + boost; phoneme classes are boosted targets: x - lnS
            const float eb = rowp[blank] - (warp_stats ? lnS + boostv : 0.f);
            k.stats[(st * B3_UPW + seg) * B3_STP + l8] = make_float2(lnS, eb);
        }
        if (c == k.n_chunks - 1) {                     // verdict on the rows of this task, published with the last chunk
            const bool badrow = (emax > 0.f) || (k.use_stats && !(fabsf(lse_chk) < 3.0e38f));
            const unsigned m = __ballot_sync(FULL, badrow);
            if (l8 == 0) k.hflags[seg] = ((m >> (seg * B3_LPU)) & 0xffu) != 0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(k.bar0 + 8u * (B3_BAR_READY + st));      // release: the DP warp may consume chunk c
        PH_T(6);

        // refill the stage the DP warp has finished with (chunk c-1) with chunk c+2
        if (c >= 1 && c + 2 < k.n_chunks) {
            const int st_p = (st == 0) ? B3_NST - 1 : st - 1;
            mbar_wait(k.bar0 + 8u * (B3_BAR_FREE + st_p), (phase >> (B3_BAR_FREE + st_p)) & 1u);
            phase ^= 1u << (B3_BAR_FREE + st_p);
            issue(c + 2, st_p);
        }
        PH_T(7);
        st = (st + 1 == B3_NST) ? 0 : st + 1;
    }

    if (a.p.reserved & BFA_FLAG_FILL_ONLY) return;   // measurement switch, see bfa_b200.h
    if constexpr (DIRECT) {
        // direct mode: the DP warp walks the path by itself (transitions only); then the two warps turn the transitions of two
        // utterances each into frame labels, timestamps and confidences (band3_direct_finish)
        // While the DP warp walks: bring the confidence inputs the fill kept (raw log-prob of the frame-wise best class + that
        // class, all four utterances) back from the pair's scratch with bulk copies and turn them into probabilities, frame by
        // frame; the finishing pass then only has to fetch the frames where the path disagrees with the guess.
        if (k.n_chunks > 0 && spec) {
            asm volatile("fence.proxy.async.global;" ::: "memory");   // the scratch this warp wrote with ordinary stores is read back by bulk copies
            mbar_wait(k.bar0 + 8u * B3_BAR_KFREE, (phase >> B3_BAR_KFREE) & 1u);   // the DP warp has left the fill: the stage ring is free
            phase ^= 1u << B3_BAR_KFREE;
            unsigned char* lpb = smem_pair + band3_d_off_lp(a.ncap, a.tpitch);
            const uint32_t cbar = k.bar0 + 8u * (B3_BAR_KREADY + 1);
            int Tq[B3_UPW];
            uint32_t nb_all = 0;
#pragma unroll
            for (int sg = 0; sg < B3_UPW; ++sg) {
                Tq[sg] = __shfl_sync(FULL, T, sg * B3_LPU);
                nb_all += (uint32_t)((Tq[sg] + 3) & ~3) * 4u + (uint32_t)((Tq[sg] + 15) & ~15);
            }
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(cbar, nb_all);           // one arrival (the barrier counts 1), all eight copies on it
#pragma unroll
                for (int sg = 0; sg < B3_UPW; ++sg)
                    if (Tq[sg] > 0) {
                        bulk_g2s(smem_u32(lpb + (size_t)sg * 5 * a.tpitch), pscr_lp + (size_t)sg * a.tpitch, (uint32_t)((Tq[sg] + 3) & ~3) * 4u, cbar);
                        bulk_g2s(smem_u32(lpb + (size_t)sg * 5 * a.tpitch + (size_t)4 * a.tpitch), pscr_gs + (size_t)sg * a.tpitch,
                                 (uint32_t)((Tq[sg] + 15) & ~15), cbar);
                    }
            }
            mbar_wait(cbar, (phase >> (B3_BAR_KREADY + 1)) & 1u);
            phase ^= 1u << (B3_BAR_KREADY + 1);
#pragma unroll 1
            for (int sg = 0; sg < B3_UPW; ++sg) {
                const int Ts = Tq[sg];
                float4* v = reinterpret_cast<float4*>(lpb + (size_t)sg * 5 * a.tpitch);
                for (int i = lane; i < (Ts + 3) / 4; i += 32) {
                    float4 x = v[i];
                    x.x = expf(x.x); x.y = expf(x.y); x.z = expf(x.z); x.w = expf(x.w);
                    v[i] = x;
                }
            }
            __syncwarp();
        }
        PH_RESET;
        band3_pair_sync(k.bar0, phase, lane);                  // the walk is done
        PH_T(0);
        band3_direct_finish<CT, LOGITS>(a, smem_pair, R, 2, 1, lane, pscr_lp != nullptr);
        PH_T(1);
        band3_pair_sync(k.bar0, phase, lane);                  // both warps are done with the task
        PH_T(2);
        PH_FLUSH;
        return;
    } else {
    // ---- back-trace, output side.  The DP warp walks the path one 32-frame block at a time and hands over the visited
    //      cells (8 * absolute cell of frames 32b + 8i + l8, lane for lane); this warp turns them into frame_phonemes /
    //      frame_phonemes_idx (:695-703), checks that the path stayed inside the band, and fetches the confidence inputs
    //      lp[f, phoneme of the stamp] (utils.py:89-103) the fill did not guess right, with 4-byte cp.async copies, one
    //      commit group per block, written out B3_GRING-1 blocks later so that no miss latency is ever waited for ----
    {
        const Item& it = *k.it;
        const int L = it.L, band = it.band, idx0 = it.idx0;
        const bool use_band = band > 0 && T > 1 && L > 1;
        const float pace_f = use_band ? (float)((double)(L - 1) / (double)(T - 1)) : 0.0f;
        const float lim_f = use_band ? (float)(band - B3_MARGIN) : INFINITY;
        const unsigned char* my_cls = k.cls8 + seg * B3_NMAX;
        const int nblk = (k.Tmax + 31) >> 5;
        const uint32_t* keepbuf = reinterpret_cast<const uint32_t*>(smem_pair) + band3_bt_keep_words(G);
        const float* gbuf = reinterpret_cast<const float*>(keepbuf + 2 * 128);
        const float* finv = gbuf + B3_GRING * 128;                 // [UPW] (verdict, final score) pairs
        const uint32_t g_s = smem_u32(gbuf) + 4u * lane;
        // frames [tf_lo, tf_hi) of the item are written; everything is indexed by the item's frame number
        const int tf_lo = o_trim, tf_hi = o_trim + rel_hi;
        int32_t* const fph_row = a.frame_ph + (o_out - o_trim);
        int32_t* const fidx_row = a.frame_idx + (o_out - o_trim);
        const bool gather_blank = !a.p.ignore_noise;
        const bool want_lp = a.path_lp != nullptr;
        PH_RESET;
        bool walk = false, illegal = false;
        uint32_t gm = 0;                                           // fetched-slot bits of the last B3_GRING blocks
        auto write_gathered = [&](int bb, uint32_t bits) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (bits & (1u << i)) plp_row[bb * 32 + 8 * i + l8] = gbuf[(bb % B3_GRING) * 128 + i * 32 + lane];
        };
        auto load_guess = [&](int bb, int (&gq)[4]) {              // the fill's guesses for block bb (255: none)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int tf = bb * 32 + 8 * i + l8;
                gq[i] = (spec && bb >= 0 && tf >= tf_lo && tf < tf_hi) ? (int)__ldcg(gcl_row + tf) : 255;
            }
        };
        int gq[4], gq_next[4];
        load_guess(nblk - 1, gq);
        for (int b = nblk - 1; b >= 0; --b) {
            load_guess(b - 1, gq_next);
            mbar_wait(k.bar0 + 8u * (B3_BAR_KREADY + (b & 1)), (phase >> (B3_BAR_KREADY + (b & 1))) & 1u);
            phase ^= 1u << (B3_BAR_KREADY + (b & 1));
            PH_T(0);
            if (b == nblk - 1) walk = k.seg_on && T > 0 && __float_as_uint(finv[2 * seg]) == 0u;
            uint32_t kv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) kv[i] = keepbuf[(b & 1) * 128 + i * 32 + lane];
            __syncwarp();
            if (lane == 0 && b >= 2) mbar_arrive(k.bar0 + 8u * (B3_BAR_KFREE + (b & 1)));   // block b-2 may overwrite the buffer
            uint32_t gbits = 0;
            int gi[4], pc[4];
            bool ph[4], on[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {                           // decode: branch-free, the four slots in parallel
                const int tf = b * 32 + 8 * i + l8;
                const uint32_t cabs = kv[i] >> 3;
                gi[i] = (int)((cabs * 43691u) >> 17);               // cabs / 3 (cabs < 2^15)
                const int kc = (int)cabs - 3 * gi[i];
                ph[i] = kc == 0;
                on[i] = walk && tf >= tf_lo && tf < tf_hi;
                // band legality of the state at frame tf (:650-653), conservative: centre of the (merged) state
                // must be at least B3_MARGIN states inside the band
                const float s_c = (float)(4 * gi[i]) - (kc == 0 ? 3.0f : (kc == 1 ? 1.5f : 0.0f));
                illegal |= walk && tf < T && fabsf(s_c - (float)tf * pace_f) > lim_f;
                pc[i] = ph[i] ? (int)my_cls[min(max(gi[i] - 1, 0), B3_NMAX - 1)] : blank;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int tf = b * 32 + 8 * i + l8;
                if (on[i]) {
                    fph_row[tf] = pc[i];
                    fidx_row[tf] = ph[i] ? idx0 + gi[i] - 1 : -1;
                    if (want_lp && (ph[i] || gather_blank) && pc[i] != gq[i]) {
                        cp_async4(g_s + (uint32_t)(b % B3_GRING) * 512u + 128u * i, k.my_src + (long long)tf * C + pc[i]);
                        gbits |= 1u << i;
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            PH_T(1);
            gm = (gm << 4) | gbits;
            asm volatile("cp.async.wait_group %0;" ::"n"(B3_GRING - 1) : "memory");
            PH_T(2);
            if (b + B3_GRING - 1 <= nblk - 1) write_gathered(b + B3_GRING - 1, (gm >> (4 * (B3_GRING - 1))) & 15u);
#pragma unroll
            for (int i = 0; i < 4; ++i) gq[i] = gq_next[i];
            PH_T(3);
        }
        cp_async_wait_all();
#pragma unroll
        for (int d = B3_GRING - 2; d >= 0; --d)
            if (d <= nblk - 1) write_gathered(d, (gm >> (4 * d)) & 15u);
        // ---- verdict: a path that left (or came too close to) the band is not provably the reference's -> exact kernel ----
        const unsigned m = __ballot_sync(FULL, illegal);
        if (k.seg_on && l8 == 0) {
            const bool bad = __float_as_uint(finv[2 * seg]) != 0u || ((m >> (seg * B3_LPU)) & 0xffu) != 0;
            if (bad) {
                const int slot = atomicAdd(a.n_retry, 1);
                a.retry_items[slot] = it;
            } else if ((it.flags & ITEM_FINAL) && a.dp_final) {
                a.dp_final[it.utt] = finv[2 * seg + 1];
            }
        }
        __syncwarp();
        PH_T(4);
        PH_FLUSH;
    }
    }   // !DIRECT
}

// ------------------------------------------------------------------------------------------------------------
// DP warp: frame loop, decision records, back-trace, outputs.
// ------------------------------------------------------------------------------------------------------------
template <int G, int CT, bool EXACT, bool DIRECT, bool LOGITS = false>
__device__ void band3_dp(const Band3Args& a, const Band3Args::Class& kc, int first, int n_valid, unsigned char* smem_pair, uint32_t* slab, uint32_t& phase,
                         int lane, int pair, const float* pscr_lp, const unsigned char* pscr_gs) {
    using S = Band3Shape<G>;
    PH_DECL;
    const size_t R = (size_t)kc.region;
    if (DIRECT) {                                      // the helper warp has planned the task: items and class table are in shared memory
        const uint32_t pb = smem_u32(smem_pair + band3_off_bars(R)) + 8u * B3_BAR_PLAN;
        mbar_wait(pb, (phase >> B3_BAR_PLAN) & 1u);
        phase ^= 1u << B3_BAR_PLAN;
        PH_T(3);
    }
    bool on;
    const Item* my_item = DIRECT ? band3_direct_item(smem_pair, R, lane, on) : band3_list_item(kc.items, first, n_valid, lane, on);
    const Band3Task<G, CT> k(a, my_item, on, smem_pair, R, lane);
    if constexpr (DIRECT) {
        if (k.n_chunks == 0) {                         // nothing of this task runs here (no such utterances, or all left to the planner chain)
            if (a.p.reserved & BFA_FLAG_FILL_ONLY) return;
            band3_pair_sync(k.bar0, phase, lane);
            band3_direct_finish<CT, LOGITS>(a, smem_pair, R, 0, 0, lane, pscr_lp != nullptr);
            band3_pair_sync(k.bar0, phase, lane);
            return;
        }
    }
    const int seg = k.seg, l8 = k.l8, C = k.C, T = k.T, Tmax = k.Tmax, n_chunks = k.n_chunks;
    const bool seg_on = k.seg_on;
    const Item& it = *k.it;
    const float NEG = a.p.neg_inf;
    const int blank = a.p.blank_id;
    const uint32_t bar0 = k.bar0;
    const int N = it.n, L = it.L, band = it.band, flags = it.flags;
    const bool use_band = band > 0 && T > 1 && L > 1;                          // :586
    const float pace_f = use_band ? (float)((double)(L - 1) / (double)(T - 1)) : 0.0f;   // :587
    const float min_lp = (flags & ITEM_FLOOR) ? a.p.min_log_prob : -INFINITY;
    const int base_max = max(0, N + 1 - S::W);
    const float* my_src = k.my_src;
    const unsigned char* my_cls = k.cls8 + seg * B3_NMAX;
    auto group_class = [&](int gi) { return (seg_on && gi >= 1 && gi <= N) ? (int)my_cls[gi - 1] : blank; };

    // chunk 0 is ready: the helper has filled the tables and reduced the first 8 rows
    mbar_wait(bar0 + 8u * B3_BAR_READY, (phase >> B3_BAR_READY) & 1u);
    phase ^= 1u << B3_BAR_READY;

    // ---- per-lane window state: groups base + l8*G + g, g = 0..G-1; -inf = invalid ----
    int base = 0;
    float P[G], M[G], B3[G];
    int cls[G];
    uint32_t acc[S::ACC];            // per group: decision words of p (2) and m, and m's word of the previous 32-frame block
    uint32_t slide_acc = 0;
#pragma unroll
    for (int i = 0; i < S::ACC; ++i) acc[i] = 0;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        P[g] = M[g] = B3[g] = -INFINITY;
        cls[g] = group_class(l8 * G + g);
    }
    if (l8 == 0) {                    // virtual frame -1: only state 0 is alive, with score 0 (:594-596)
        if (EXACT) B3[0] = 0.0f;      // state 0 = b3 of group 0
        else M[0] = 0.0f;             // reduced form: it lives in m of group 0 (b3 of group 0 follows it one frame later)
    }
    int next_cls = group_class(S::W);   // class of the group that enters at the next slide (last lane of the segment)
    const bool seg_first = l8 == 0, seg_last = l8 == B3_LPU - 1;

    bool bad = false;                // needs the exact path
    float fin_val = NEG;
    int fin_cell = 0, fin_base = 0;
    PH_T(0);

    int st = 0;                                        // stage of chunk c
    int nslide_next = 0;
    for (int c = 0; c < n_chunks; ++c) {
        const int t0 = c * B3_ROWS;

        // ---- slide the window to the group of the lower band edge at frame t0-1 (fp32 estimate, never above the
        //      exact :651 value, at most one state below it) ----
        int nslide = nslide_next;                      // computed during the previous chunk
        slide_acc = (slide_acc << 4) | (uint32_t)nslide;
        while (__any_sync(FULL, nslide > 0)) {
            const float nP = __shfl_down_sync(FULL, P[0], 1), nM = __shfl_down_sync(FULL, M[0], 1);
            const float n3 = EXACT ? __shfl_down_sync(FULL, B3[0], 1) : 0.f;
            const int nc = __shfl_down_sync(FULL, cls[0], 1);
            uint32_t nA[4];                            // the decision words of the current block move with their group
#pragma unroll
            for (int i = 0; i < 4; ++i) nA[i] = (EXACT || i != 1) ? __shfl_down_sync(FULL, acc[i], 1) : 0u;   // word 1 is unused in the reduced form
            if (nslide > 0) {
#pragma unroll
                for (int i = 0; i < S::ACC - 4; ++i) acc[i] = acc[i + 4];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[S::ACC - 4 + i] = seg_last ? 0u : nA[i];
#pragma unroll
                for (int g = 0; g + 1 < G; ++g) { P[g] = P[g + 1]; M[g] = M[g + 1]; B3[g] = B3[g + 1]; cls[g] = cls[g + 1]; }
                P[G - 1] = seg_last ? -INFINITY : nP;
                M[G - 1] = seg_last ? -INFINITY : nM;
                B3[G - 1] = seg_last ? -INFINITY : n3;
                cls[G - 1] = seg_last ? next_cls : nc;
                base += 1;
                if (seg_last) next_cls = group_class(base + S::W);
            }
            --nslide;
        }
        if (use_band) {                                // slide of the next chunk: lower edge at frame t0+7
            const float lo_est = floorf(fmaf((float)(t0 + B3_ROWS - 1), pace_f, -(float)band - 0.01f));
            nslide_next = min(max(((int)lo_est + 3) >> 2, 0), base_max) - base;
        }
        PH_T(1);

        if (c > 0) {                                   // the helper has reduced chunk c (which implies that it has landed)
            mbar_wait(bar0 + 8u * (B3_BAR_READY + st), (phase >> (B3_BAR_READY + st)) & 1u);
            phase ^= 1u << (B3_BAR_READY + st);
        }
        PH_T(2);

        const float* seg_rows = k.stage_buf + st * k.stage_floats + seg * k.seg_stride + k.lead;   // the 8 staged rows of this utterance
        const int fin_r = T - 1 - t0;                                      // row of the last frame if it is in this chunk
        const bool fin_here = __any_sync(FULL, fin_r >= 0 && fin_r < B3_ROWS);
        const float* xg[G];                                                // row 0 address of each group's class
#pragma unroll
        for (int g = 0; g < G; ++g) xg[g] = seg_rows + cls[g];
        const float2* stp = k.stats + (st * B3_UPW + seg) * B3_STP;
        float xp_n[G];                                                     // raw emissions are fetched one frame ahead
#pragma unroll
        for (int g = 0; g < G; ++g) xp_n[g] = xg[g][0];
        float2 st_nx = stp[0];

        // The four decision words per group that go into a record: p's two (from b3' / from m'), m's, b3's.
        // Reduced form: b3's word is m's one frame later (slot 3 holds m's word of the previous block), and p's single
        // "entered" word splits by the b3 word of the group to its left (b3' == m' exactly when m' stayed).
        auto record_words = [&](uint32_t (&w)[S::ACC]) {
            if (EXACT) {
#pragma unroll
                for (int i = 0; i < S::ACC; ++i) w[i] = acc[i];
            } else {
                uint32_t w3[G];
#pragma unroll
                for (int g = 0; g < G; ++g) w3[g] = __funnelshift_r(acc[4 * g + 2], acc[4 * g + 3], 1);
                uint32_t left = __shfl_up_sync(FULL, w3[G - 1], 1);
                if (seg_first) left = 0u;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const uint32_t y = g > 0 ? w3[g > 0 ? g - 1 : 0] : left, x = acc[4 * g];
                    w[4 * g] = x & ~y;
                    w[4 * g + 1] = x & y;
                    w[4 * g + 2] = acc[4 * g + 2];
                    w[4 * g + 3] = w3[g];
                }
            }
        };

        // A record is stored as the back-trace reads it: per cell one (w0, w1) pair -- p: (from b3', from m'), m: (from p, 0),
        // b3: (from m, 0) -- in cell order, so the bulk copy that brings it back IS the staging; then one slide word per lane.
        auto write_record = [&](uint32_t* rec, const uint32_t (&w)[S::ACC], int sh, uint32_t slide_word) {
            uint2* cells = reinterpret_cast<uint2*>(rec) + seg * S::CELLS + l8 * G * 3;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                cells[3 * g + 0] = make_uint2(w[4 * g] << sh, w[4 * g + 1] << sh);
                cells[3 * g + 1] = make_uint2(w[4 * g + 2] << sh, 0u);
                // b3 advanced from m at frame t exactly when m came from p at t-1 (reduced form), and m and b3 produce the same
                // output (blank, -1): attribute frame t-1 to b3 as well and hop straight to p -- one walk iteration per
                // phoneme less.  (First frame of a block: the ordinary one-cell hop, the bit cannot move into the previous record.)
                const uint32_t w3 = w[4 * g + 3] << sh;
                cells[3 * g + 2] = EXACT ? make_uint2(w3, 0u) : make_uint2(w3 & 0x80000000u, w3 << 1);
            }
            rec[B3_UPW * S::CELLS * 2 + lane] = slide_word;
        };

        auto frame = [&](const int r, auto check_fin) {
            constexpr bool CHECK = decltype(check_fin)::value;
            const float lnS = st_nx.x, eb = st_nx.y;
            float ep[G];
#pragma unroll
            for (int g = 0; g < G; ++g) ep[g] = fmaxf(xp_n[g] - lnS, min_lp);
            if (CHECK || r + 1 < B3_ROWS) {
                const int rn = (r + 1 < B3_ROWS) ? r + 1 : r;
                st_nx = stp[rn];
#pragma unroll
                for (int g = 0; g < G; ++g) xp_n[g] = xg[g][rn * C];
            }

            // ---- DP update, right-most group first so that left neighbours are still frame t-1 ----
            float lm = __shfl_up_sync(FULL, M[G - 1], 1), l3 = EXACT ? __shfl_up_sync(FULL, B3[G - 1], 1) : 0.f;
            if (seg_first) { lm = -INFINITY; l3 = -INFINITY; }   // nothing (or a dropped group, dead for every legal path) to the left
#pragma unroll
            for (int g = G - 1; g >= 0; --g) {
                const float Lm = (g > 0) ? M[(g > 0) ? g - 1 : 0] : lm;
                const float SP = P[g] + eb, SM = M[g] + eb;
                uint32_t* A = &acc[4 * g];
                if (EXACT) {
                    const float L3 = (g > 0) ? B3[(g > 0) ? g - 1 : 0] : l3;
                    const float c0 = P[g] + ep[g], c1 = L3 + ep[g], c2 = Lm + ep[g];   // stay / advance from b3' / skip from b2' (= m')
                    const float S3 = B3[g] + eb;
                    // p  : first max of (c0, c1, c2)                      (:645)
                    const float m01 = fmaxf(c0, c1);
                    b3_push(A[0], c0, c1);
                    b3_push(A[1], m01, c2);
                    P[g] = fmaxf(m01, c2);
                    // b3 : (stay S3, advance SM)
                    b3_push(A[3], S3, SM);
                    B3[g] = fmaxf(S3, SM);
                } else {
                    // p  : b3' <= m' (header, 1b), so the best way in is always worth c2 and one decision bit is enough:
                    //      A[0] = "entered from the previous group"; from b3' or from m' is decided by m's history (flush)
                    const float c0 = P[g] + ep[g], c2 = Lm + ep[g];
                    b3_push(A[0], c0, c2);
                    P[g] = fmaxf(c0, c2);
                    if (CHECK) B3[g] = SM;     // b3's value is m's stay candidate; only the final-state rule looks at it
                }
                // m  : (stay SM, from p SP)  -- b1<-{b1,p} / b2<-{b2,b1,p} merged
                b3_push(A[2], SM, SP);
                M[g] = fmaxf(SM, SP);
            }

            if (CHECK) {
                const bool fin = (r == fin_r);
                const unsigned fin_mask = __ballot_sync(FULL, fin);
                uint32_t wrec[S::ACC];
                if (fin_mask != 0u) record_words(wrec);
                if (fin) {
                    const int t = t0 + r;
                    // last record of this utterance, left-aligned so that frame 32b+q sits at bit 31-q
                    const int sh = 31 - (t & 31);
                    write_record(slab + (size_t)(t >> 5) * S::REC * 32, wrec, sh, slide_acc << (4 * (3 - (c & 3))));
                    // ---- final state (:656-682) from the window at frame T-1; cells are 3*(group - base) + {0:p, 1:m, 2:b3} ----
                    float bv = -INFINITY;
                    int bs = -1;
                    if (!a.p.truly_forced) {
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            const int gi = base + l8 * G + g;
                            const int cell = 3 * (l8 * G + g);
                            // states 4gi-3 (p), 4gi-2 / 4gi-1 (m), 4gi (b3) must exist: 0 <= s < L
                            if (gi >= 1 && gi <= N && P[g] > NEG && (bs < 0 || P[g] > bv)) { bv = P[g]; bs = cell; }
                            if (gi >= 1 && gi <= N && M[g] > NEG && (bs < 0 || M[g] > bv)) { bv = M[g]; bs = cell + 1; }
                            if (gi >= 0 && gi <= N && B3[g] > NEG && (bs < 0 || B3[g] > bv)) { bv = B3[g]; bs = cell + 2; }
                        }
                    } else {
                        // L-1 = 4N is b3 of group N, L-2 its b2 (= m)
                        const int w = N - base;
                        const int lw = w / G, sl = w - lw * G;
                        if (w >= 0 && w < S::W && lw == l8) {
#pragma unroll
                            for (int g = 0; g < G; ++g)
                                if (g == sl) {
                                    if (B3[g] > NEG) { bv = B3[g]; bs = 3 * w + 2; }
                                    else if (L >= 2 && M[g] > NEG) { bv = M[g]; bs = 3 * w + 1; }
                                }
                        }
                    }
                    // segment reduction: max value, ties -> lower state
#pragma unroll
                    for (int d = 1; d < B3_LPU; d <<= 1) {
                        const float ov = __shfl_xor_sync(fin_mask, bv, d);
                        const int os = __shfl_xor_sync(fin_mask, bs, d);
                        if (os >= 0 && (bs < 0 || ov > bv || (ov == bv && os < bs))) { bv = ov; bs = os; }
                    }
                    if (bs < 0) bad = true;   // nothing valid at the end: degenerate -> exact path
                    fin_val = bv;
                    fin_cell = bs < 0 ? 0 : bs;
                    fin_base = base;
                }
            }
        };

        if (!fin_here) {
#pragma unroll
            for (int r = 0; r < B3_ROWS; ++r) frame(r, std::false_type{});
        } else {
#pragma unroll 1
            for (int r = 0; r < B3_ROWS; ++r) frame(r, std::true_type{});
        }
        PH_T(4);
        // ---- flush one full 32-frame record (utterances that end inside this chunk flushed at their last frame)
        if ((c & 3) == 3) {
            uint32_t wrec[S::ACC];
            record_words(wrec);
            if (T - 1 > t0 + B3_ROWS - 1) {
                write_record(slab + (size_t)(c >> 2) * S::REC * 32, wrec, 0, slide_acc);
            }
            if (!EXACT) {
#pragma unroll
                for (int g = 0; g < G; ++g) acc[4 * g + 3] = acc[4 * g + 2];   // m's word of the block just finished
            }
        }
        __syncwarp();                                  // every lane is done with stage st
        if (c + B3_NST < n_chunks && lane == 0) mbar_arrive(bar0 + 8u * (B3_BAR_FREE + st));   // the helper may refill it
        PH_T(11);
        st = (st + 1 == B3_NST) ? 0 : st + 1;
    }
    PH_T(5);
    __syncwarp();
    if (seg_on && k.hflags[seg]) bad = true;           // helper: an emission > 0 (raw mode) or a non-finite log-sum-exp
    {   // make `bad` uniform per segment
        const unsigned m = __ballot_sync(FULL, bad);
        bad = ((m >> (seg * B3_LPU)) & 0xffu) != 0;
    }

    if (a.p.reserved & BFA_FLAG_FILL_ONLY) {   // measurement switch: time the fill by itself (no outputs)
        PH_FLUSH;
        return;
    }
    if constexpr (DIRECT) {
        if (lane == 0) mbar_arrive(bar0 + 8u * B3_BAR_KFREE);   // the stage ring is free: the helper may stage the confidence inputs there
        // ---- back-trace (:686-703), direct mode: the walk only notes where the path ENTERS a cell (one event per run:
        //      cell and first frame, latest first); band3_direct_finish turns the events into everything else.  Same records,
        //      same run-per-iteration walk as the list mode below, without the per-frame hand-over to the helper warp ----
        const bool walk = seg_on && !bad && T > 0;
        const int last_blk = (T - 1) >> 5;
        uint32_t* recbuf = reinterpret_cast<uint32_t*>(smem_pair + band3_d_off_rec(a.ncap, a.tpitch));   // [B3_NREC][REC][32]
        const int evcap = band3_evcap(a.ncap);
        uint32_t* ev = reinterpret_cast<uint32_t*>(smem_pair + band3_d_off_ev()) + seg * evcap;
        float* finv = reinterpret_cast<float*>(smem_pair + band3_d_off_fin(a.ncap));           // [UPW] (verdict, final score)
        int* nevs = reinterpret_cast<int*>(smem_pair + band3_d_off_fin(a.ncap) + 32);          // [UPW] events per utterance
        constexpr uint32_t RECB = (uint32_t)S::REC * 128u;
        const uint32_t rec_s = smem_u32(recbuf);
        const uint32_t cellbase = rec_s + (uint32_t)seg * S::CELLS * 8u;
        uint32_t A = cellbase + 8u * (uint32_t)fin_cell;
        uint32_t K = 24u * (uint32_t)fin_base - cellbase;                                      // A + K = 8 * cabs
        const int nblk = (Tmax + 31) >> 5;
        const uint32_t rbar0 = bar0 + 8u * B3_BAR_REC;
        auto issue_rec = [&](int blk) {
            if (lane == 0) {
                const uint32_t bar = rbar0 + 8u * (blk % B3_NREC);
                mbar_expect_tx(bar, RECB);
                bulk_g2s(rec_s + (uint32_t)(blk % B3_NREC) * RECB, slab + (size_t)blk * S::REC * 32, RECB, bar);
            }
        };
        asm volatile("fence.proxy.async.global;" ::: "memory");   // the slab was written with ordinary stores
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the stage ring was read through the generic proxy
        __syncwarp();
#pragma unroll
        for (int d = 1; d <= B3_NREC; ++d)
            if (nblk - d >= 0) issue_rec(nblk - d);
        uint32_t rphase = phase >> B3_BAR_REC;
        int nev = 0;
        int last_f = -1;
        PH_T(6);
        for (int b = nblk - 1; b >= 0; --b) {
            mbar_wait(rbar0 + 8u * (b % B3_NREC), (rphase >> (b % B3_NREC)) & 1u);
            rphase ^= 1u << (b % B3_NREC);
            const bool live = walk && b <= last_blk;
            const uint32_t boff = (uint32_t)(b % B3_NREC) * RECB;
            const uint32_t sfw = live ? recbuf[(b % B3_NREC) * S::REC * 32 + B3_UPW * S::CELLS * 2 + lane] : 0u;
            const uint32_t nsl = (sfw & 15u) + ((sfw >> 4) & 15u) + ((sfw >> 8) & 15u) + ((sfw >> 12) & 15u);
            PH_T(7);
            A += boff;
            K -= boff;
            uint32_t mask = live ? 0xffffffffu : 0u;
            uint32_t c0, c1, p0, p1, q0, q1;
            auto lds2 = [](uint32_t addr, uint32_t& x, uint32_t& y) {
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(addr) : "memory");
            };
            lds2(A, c0, c1);
            lds2(A - 8u, p0, p1);
            lds2(A - 16u, q0, q1);
            const uint32_t fbase = (uint32_t)b * 32u;
            while (mask != 0u) {
                const uint32_t stop = (c0 | c1) & mask;
                const uint32_t bit = stop & (0u - stop);      // frame whose decision ends the run (0: none left in this block)
                const uint32_t maskn = bit * 0xfffffffeu;     // frames strictly before it
                if (bit != 0u) {                              // the path entered cell (A + K) / 8 at frame fbase + clz(bit)
                    last_f = (int)(fbase + (uint32_t)__clz(bit));
                    if (l8 == 0 && nev < evcap) ev[nev] = ((A + K) << 19) | (uint32_t)last_f;
                    ++nev;
                }
                const bool m1 = (c0 & bit) != 0u, m2 = (c1 & bit) != 0u;
                A = m2 ? A - 16u : (m1 ? A - 8u : A);
                c0 = m2 ? q0 : (m1 ? p0 : c0);
                c1 = m2 ? q1 : (m1 ? p1 : c1);
                mask = maskn;
                lds2(A - 8u, p0, p1);
                lds2(A - 16u, q0, q1);
            }
            __syncwarp();                          // every lane is done with this record buffer
            if (b >= B3_NREC) issue_rec(b - B3_NREC);
            A += 24u * nsl - boff;
            K -= 24u * nsl - boff;
            PH_T(8);
        }
        // the run the utterance starts in has no entry decision: note it as entered at frame 0
        if (walk && last_f != 0) {
            if (l8 == 0 && nev < evcap) ev[nev] = ((A + K) << 19);
            ++nev;
        }
        if (nev > evcap) bad = true;               // cannot happen on a monotone path (<= 3 events per phoneme); defensive
        if (l8 == 0) {
            finv[2 * seg] = __uint_as_float((bad || !seg_on) ? 1u : 0u);
            finv[2 * seg + 1] = fin_val;
            nevs[seg] = nev;
        }
        phase = (phase & ~(((1u << B3_NREC) - 1u) << B3_BAR_REC)) | ((rphase & ((1u << B3_NREC) - 1u)) << B3_BAR_REC);
        __syncwarp();
        band3_pair_sync(k.bar0, phase, lane);                     // events and verdicts are in shared memory
        PH_T(9);
        band3_direct_finish<CT, LOGITS>(a, smem_pair, R, 0, 0, lane, pscr_lp != nullptr);
        PH_T(10);
        band3_pair_sync(k.bar0, phase, lane);                     // both warps are done with the task
        PH_T(12);
        PH_FLUSH;
        return;
    } else {
    // ---- back-trace (:686-703), walking side ----
    // A cell is addressed by its window-relative index ci = 3*(group - base) + k.  Staging lays a record out as
    // bt2[seg][ci] = (first decision word, second decision word or 0), so one 64-bit shared load per run yields
    // (b0, b1) and every transition is  ci -= b1 ? 2 : b0  (p: from m' = -2 / from b3' = -1, m: from p = -1,
    // b3: from m = -1).  The decision words of a block are aligned to the window position at the END of the block (the
    // fill shifts its accumulators together with the window), so inside a block the walk never sees a slide; between
    // blocks the cell index moves by 3 * (groups slid during the block).  cabs = 3*base + ci is slide-invariant.
    // Records of frames past the end of an utterance are zero ("stay"); the cell is kept as a shared-memory byte address
    // A so the loop-carried chain is LDS -> mask -> lowest set bit -> 2 selects.  The visited cells go to the helper warp
    // through a double-buffered shared array; it writes the outputs while this warp walks the next block.
    const bool walk = seg_on && !bad && T > 0;
    const int last_blk = (T - 1) >> 5;
    // Shared staging (aliases the stage ring): B3_NREC record buffers | two visited-cell buffers | fetch ring | verdicts.
    // A record comes back from the slab by bulk async copy (TMA) on its own mbarrier, B3_NREC blocks ahead, and is already
    // laid out as the cells of its block (see write_record): nothing to stage.
    uint32_t* recbuf = reinterpret_cast<uint32_t*>(smem_pair);                                // [B3_NREC][REC][32]
    uint32_t* keepbuf = reinterpret_cast<uint32_t*>(smem_pair) + band3_bt_keep_words(G);     // [2][4][32]
    float* finv = reinterpret_cast<float*>(keepbuf + 2 * 128 + B3_GRING * 128);              // [UPW] (verdict, final score)
    constexpr uint32_t RECB = (uint32_t)S::REC * 128u;                                         // bytes of one record
    const uint32_t rec_s = smem_u32(recbuf);
    const uint32_t cellbase = rec_s + (uint32_t)seg * S::CELLS * 8u;                           // this utterance's cells in record buffer 0
    uint32_t A = cellbase + 8u * (uint32_t)fin_cell;                                           // cell address, kept relative to buffer 0
    uint32_t K = 24u * (uint32_t)fin_base - cellbase;                                          // A + K = 8 * cabs
    const int nblk = (Tmax + 31) >> 5;
    const uint32_t rbar0 = bar0 + 8u * B3_BAR_REC;                                             // the record barriers
    auto issue_rec = [&](int blk) {
        if (lane == 0) {
            const uint32_t bar = rbar0 + 8u * (blk % B3_NREC);
            mbar_expect_tx(bar, RECB);
            bulk_g2s(rec_s + (uint32_t)(blk % B3_NREC) * RECB, slab + (size_t)blk * S::REC * 32, RECB, bar);
        }
    };
    asm volatile("fence.proxy.async.global;" ::: "memory");   // the slab was written with ordinary stores, the bulk copies read it through the async proxy
    __syncwarp();
#pragma unroll
    for (int d = 1; d <= B3_NREC; ++d)
        if (nblk - d >= 0) issue_rec(nblk - d);
    if (l8 == 0) {                             // what the helper needs for the verdict (published by the first KREADY arrive)
        finv[2 * seg] = __uint_as_float(bad ? 1u : 0u);
        finv[2 * seg + 1] = fin_val;
    }
    uint32_t rphase = phase >> B3_BAR_REC;     // phase bits of the record barriers
    PH_T(6);
    for (int b = nblk - 1; b >= 0; --b) {
        mbar_wait(rbar0 + 8u * (b % B3_NREC), (rphase >> (b % B3_NREC)) & 1u);     // record of block b has landed
        rphase ^= 1u << (b % B3_NREC);
        const bool live = walk && b <= last_blk;   // later blocks of a shorter utterance hold stale records: not walked
        const uint32_t boff = (uint32_t)(b % B3_NREC) * RECB;
        const uint32_t sfw = live ? recbuf[(b % B3_NREC) * S::REC * 32 + B3_UPW * S::CELLS * 2 + lane] : 0u;
        const uint32_t nsl = (sfw & 15u) + ((sfw >> 4) & 15u) + ((sfw >> 8) & 15u) + ((sfw >> 12) & 15u);   // groups slid during the block
        PH_T(7);
        if (b + 2 <= nblk - 1) {               // the helper has read the visited cells of block b+2 out of this buffer
            mbar_wait(bar0 + 8u * (B3_BAR_KFREE + (b & 1)), (phase >> (B3_BAR_KFREE + (b & 1))) & 1u);
            phase ^= 1u << (B3_BAR_KFREE + (b & 1));
        }
        PH_T(12);
        A += boff;                             // same cell in the buffer this block's record landed in
        K -= boff;

        // ---- walk the 32 frames of the block, one RUN per iteration (bit 31-f belongs to frame f): the path stays in
        //      its cell until a decision bit of that cell is set.  Utterances leave the loop on their own ----
        //      The decision words of the two cells the path can move to are fetched one iteration ahead, so the shared-memory
        //      latency is off the loop-carried chain (mask -> lowest set bit -> selects).
        uint32_t keep[4] = {0u, 0u, 0u, 0u};             // 8 * cabs of frames 32b + 8i + l8
        uint32_t mask = live ? 0xffffffffu : 0u;          // frames not yet walked (none for an utterance that is not walked)
        uint32_t c0, c1, p0, p1, q0, q1;                  // words of the cells A, A - 1 cell, A - 2 cells
        auto lds2 = [](uint32_t addr, uint32_t& x, uint32_t& y) {
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(addr) : "memory");
        };
        lds2(A, c0, c1);
        lds2(A - 8u, p0, p1);
        lds2(A - 16u, q0, q1);
        do {
            const uint32_t stop = (c0 | c1) & mask;
            const uint32_t bit = stop & (0u - stop);      // frame whose decision ends the run (0: none left in this block)
            const uint32_t maskn = bit * 0xfffffffeu;     // frames strictly before it  (= -(2*bit); 0 when bit is 0 or bit 31)
            const uint32_t run = mask & ~maskn;           // frames of this run
            const uint32_t AK = A + K;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (run & (0x80000000u >> (8 * i + l8))) keep[i] = AK;
            const bool m1 = (c0 & bit) != 0u, m2 = (c1 & bit) != 0u;   // from b3' / from m' (m' wins)
            A = m2 ? A - 16u : (m1 ? A - 8u : A);
            c0 = m2 ? q0 : (m1 ? p0 : c0);
            c1 = m2 ? q1 : (m1 ? p1 : c1);
            mask = maskn;
            lds2(A - 8u, p0, p1);
            lds2(A - 16u, q0, q1);
        } while (mask != 0u);
        __syncwarp();                          // every lane is done with this record buffer
        if (b >= B3_NREC) issue_rec(b - B3_NREC);
        PH_T(8);
#pragma unroll
        for (int i = 0; i < 4; ++i) keepbuf[(b & 1) * 128 + i * 32 + lane] = keep[i];
        A += 24u * nsl - boff;                 // the record of block b-1 is aligned to the window before these slides
        K -= 24u * nsl - boff;
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 8u * (B3_BAR_KREADY + (b & 1)));
        PH_T(9);
    }
    phase = (phase & ~(((1u << B3_NREC) - 1u) << B3_BAR_REC)) | ((rphase & ((1u << B3_NREC) - 1u)) << B3_BAR_REC);
    PH_T(10);
    PH_FLUSH;
    }   // !DIRECT
}

// Direct mode, last phase.  One warp of the pair finishes two utterances (seg0, seg0 + 1), one after the other, `which` = 0 for
// the DP warp, 1 for the helper warp:
//   1. the events of the walk (cell, first frame; latest first) become one (start, end) pair per phoneme -- on a monotone path
//      every phoneme state is entered exactly once, and its run ends where the next event begins -- the band-legality check of
//      the lazy band (header, 2.) runs on the two end frames of every run (the distance to the band centre is linear in between),
//      and both ends of every phoneme run are marked in a bit map over the frames;
//   2. one sweep over the frames, 32 per step (lane = frame, nothing data-dependent in the control flow): the number of marks at
//      or before a frame says whether it lies in a phoneme run and in which; frame_phonemes / frame_phonemes_idx (:695-703) go
//      straight to global memory, coalesced; where the fill's guess (the frame-wise best class, whose probability the helper warp
//      has already left in lp_s) is not the frame's phoneme, lp[f, phoneme] is fetched from the posteriors (4-byte asynchronous
//      copies, all in flight together) and exponentiated afterwards;
//   3. one lane per phoneme emits its timestamp (assort_frames, :777-834: with ignore_noise the stamps are exactly the phoneme
//      runs, blanks always separate two phonemes) and its confidence (utils.py:70-113, sequential fp32 like the reference).
// Utterances that cannot be finished here (not planned, illegal path, degenerate end) are left to the planner chain, or flagged
// BFA_ST_DEFERRED when no chain follows.
// Not inlined on purpose: this is run-once code, and one copy shared by all fourteen warps of the CTA (which arrive here at about
// the same time) keeps its instruction-cache misses to one warp's worth.
template <int CT, bool LOGITS>
__device__ __noinline__ void band3_direct_finish(const Band3Args& a, unsigned char* smem_pair, size_t R, int seg0, int which, int lane, bool have_spec) {
    const int C = CT ? CT : a.C;
    const Item* items = reinterpret_cast<const Item*>(smem_pair + band3_off_items(R));
    const int* ustate = reinterpret_cast<const int*>(smem_pair + band3_off_flags(R)) + B3_UPW;
    const unsigned char* cls_all = smem_pair + band3_off_cls(R);
    const int evcap = band3_evcap(a.ncap), ncap2 = (a.ncap + 1) & ~1;
    const uint32_t* ev_all = reinterpret_cast<const uint32_t*>(smem_pair + band3_d_off_ev());
    const float* finv = reinterpret_cast<const float*>(smem_pair + band3_d_off_fin(a.ncap));
    const int* nevs = reinterpret_cast<const int*>(smem_pair + band3_d_off_fin(a.ncap) + 32);
    int2* se_all = reinterpret_cast<int2*>(smem_pair + band3_d_off_se(a.ncap));
    const int wpitch = a.tpitch / 32 + 1;
    uint32_t* bnd = reinterpret_cast<uint32_t*>(smem_pair + band3_d_off_bits(a.ncap)) + (size_t)which * 4 * wpitch;   // run-boundary bits
    unsigned char* mnib = reinterpret_cast<unsigned char*>(bnd + wpitch);               // [quad of frames] which of its frames were fetched
    unsigned char* lpb = smem_pair + band3_d_off_lp(a.ncap, a.tpitch);
    const int blank = a.p.blank_id;
    const bool want_stamps = a.stamps != nullptr;
    const bool want_conf = want_stamps && a.conf != nullptr;
    bool waited = false;                       // griddepcontrol.wait executed (before the first write another grid could see)
    FIN_DECL;
    for (int q = 0; q < 2; ++q) {
        const int seg = seg0 + q;
        const int state = ustate[seg];
        if (state == B3_U_NONE) continue;      // no such utterance (uniform over the warp)
        const Item& it = items[seg];
        const int u = it.utt, T = it.T, N = it.n;
        bool bad = state != B3_U_RUN || __float_as_uint(finv[2 * seg]) != 0u;
        // the helper warp has left exp(raw log-prob of the frame-wise best class) and that class here (see band3_helper)
        const bool spec = want_conf && have_spec && (it.flags & ITEM_STATS) != 0;
        float* lp_s = reinterpret_cast<float*>(lpb + (size_t)seg * 5 * a.tpitch);
        const unsigned char* gs_s = lpb + (size_t)seg * 5 * a.tpitch + (size_t)4 * a.tpitch;
        int2* se = se_all + seg * ncap2;
        const int nw = (T + 31) >> 5;
        if (!bad) {
            // ---- 1. events -> (start, end) per phoneme, legality, run-boundary bits ----
            for (int i = lane; i < nw; i += 32) bnd[i] = 0u;
            for (int g = lane; g < N; g += 32) se[g] = make_int2(0, 0);
            __syncwarp();
            const uint32_t* ev = ev_all + seg * evcap;
            const int nev = nevs[seg];
            const int L = it.L, band = it.band;
            const bool use_band = band > 0 && T > 1 && L > 1;
            const float pace_f = use_band ? (float)((double)(L - 1) / (double)(T - 1)) : 0.0f;
            const float lim_f = use_band ? (float)(band - B3_MARGIN) : INFINITY;
            bool illegal = false;
            int pcount = 0;
            for (int i = lane; i < nev; i += 32) {
                const uint32_t e = ev[i];
                const uint32_t cabs = e >> 22;
                const int f = (int)(e & 0x3fffffu);
                const int f_end = i == 0 ? T : (int)(ev[i - 1] & 0x3fffffu);         // the run ends where the next (later) one begins
                const int g = (int)((cabs * 43691u) >> 17);                            // cabs / 3 (cabs < 2^15)
                const int kc = (int)cabs - 3 * g;
                // band legality (:650-653), conservative: the centre of the (merged) state stays B3_MARGIN states inside the band
                const float s_c = (float)(4 * g) - (kc == 0 ? 3.0f : (kc == 1 ? 1.5f : 0.0f));
                illegal |= f_end <= f || f_end > T || fabsf(s_c - (float)f * pace_f) > lim_f || fabsf(s_c - (float)(f_end - 1) * pace_f) > lim_f;
                if (kc == 0) {
                    if (g >= 1 && g <= N && f_end <= T) {
                        se[g - 1] = make_int2(f, f_end);
                        ++pcount;
                        atomicOr(&bnd[f >> 5], 1u << (f & 31));
                        if (f_end < T) atomicOr(&bnd[f_end >> 5], 1u << (f_end & 31));
                    } else illegal = true;
                }
            }
            pcount = __reduce_add_sync(FULL, pcount);
            bad = __any_sync(FULL, illegal) || pcount != N;
            __syncwarp();
            FIN_T(1);
        }
        if (!waited) { pdl_wait(); waited = true; }
        FIN_T(4);
        if (!bad) {
            // ---- 2. frame sweep: every lane takes four consecutive frames per step (128 frames per step of the warp) ----
            const unsigned char* my_cls = cls_all + seg * B3_NMAX;
            const float* src = a.logp + it.lp_off;
            int32_t* fph = a.frame_ph + it.out_off;
            int32_t* fidx = a.frame_idx + it.out_off;
            const int room = (int)min((long long)T, max(it.out_lim - it.out_off, 0LL));
            const bool vec_ok = (((unsigned long long)fph | (unsigned long long)fidx) & 15ull) == 0;
            const uint32_t lp_sa = smem_u32(lp_s);
            const int idx0 = it.idx0;
            bool my_miss = false;
            int carry = 0;                                       // marks before the current group of 32 words
            for (int w0 = 0; w0 < nw; w0 += 32) {
                // exclusive prefix of the marks per word over this group of 32 words (lane = word)
                const int mine = (w0 + lane < nw) ? __popc(bnd[w0 + lane]) : 0;
                int inc = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = __shfl_up_sync(FULL, inc, d);
                    if (lane >= d) inc += o;
                }
                const int excl = carry + inc - mine;
                carry += __shfl_sync(FULL, inc, 31);
                const int qn = (min(32, nw - w0) + 3) >> 2;      // steps of 4 words = 128 frames in this group
#pragma unroll 2
                for (int i = 0; i < qn; ++i) {
                    const int wj = 4 * i + (lane >> 3);          // this lane's word within the group
                    const int f0 = (w0 + wj) * 32 + 4 * (lane & 7);
                    const int bo = 4 * (lane & 7);
                    const uint32_t word = (w0 + wj < nw) ? bnd[w0 + wj] : 0u;
                    int cnt = __shfl_sync(FULL, excl, wj) + __popc(word & ((1u << bo) - 1u));
                    int pcv[4], ixv[4];
                    uint32_t nib = 0;
                    const uint32_t gw = (spec && f0 < T) ? *reinterpret_cast<const uint32_t*>(gs_s + f0) : 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        cnt += (int)((word >> (bo + k)) & 1u);
                        const bool inside = (cnt & 1) != 0 && f0 + k < T;
                        const int g = (cnt + 1) >> 1;            // 1-based phoneme of the run the frame lies in
                        pcv[k] = blank; ixv[k] = -1;
                        if (inside) {
                            pcv[k] = my_cls[g - 1];
                            ixv[k] = idx0 + g - 1;
                            bool miss = want_conf && !(spec && (int)((gw >> (8 * k)) & 255u) == pcv[k]);
#ifdef BFA_DBG_NOMISS
                            miss = false;
#endif
                            if (miss) {
                                cp_async4(lp_sa + 4u * (uint32_t)(f0 + k), src + (long long)(f0 + k) * C + pcv[k]);
                                nib |= 1u << k;
                            }
                        }
                    }
#ifndef BFA_DBG_NOSTORE
                    if (vec_ok && f0 + 3 < room) {
                        *reinterpret_cast<int4*>(fph + f0) = make_int4(pcv[0], pcv[1], pcv[2], pcv[3]);
                        *reinterpret_cast<int4*>(fidx + f0) = make_int4(ixv[0], ixv[1], ixv[2], ixv[3]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (f0 + k < room) { fph[f0 + k] = pcv[k]; fidx[f0 + k] = ixv[k]; }
                    }
#endif
                    if (f0 < T) mnib[f0 >> 2] = (unsigned char)nib;
                    my_miss |= nib != 0u;
                }
            }
            FIN_T(5);
            if (__any_sync(FULL, my_miss)) {                     // what arrived is a log-prob: exponentiate it in place
                cp_async_wait_all();
                FIN_T(6);
                __syncwarp();                                    // every lane's copies have landed
                const float* rl = LOGITS ? a.row_lse + it.out_off : nullptr;   // logits mode: what arrived is a logit; the helper warp left the row's log-sum-exp
                for (int q4 = 0; q4 < (T + 3) >> 2; q4 += 32) {
                    const int qq = q4 + lane;
                    if (qq < (T + 3) >> 2) {
                        const uint32_t nib = mnib[qq];
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if ((nib >> k) & 1u) lp_s[4 * qq + k] = expf(lp_s[4 * qq + k] - (LOGITS ? rl[4 * qq + k] : 0.0f));
                    }
                }
            }
            __syncwarp();
            FIN_T(7);
            // ---- 3. timestamps + confidences, one phoneme per lane ----
            int st = BFA_ST_OK;
            if (want_stamps) {
                BfaStamp* out = a.stamps + (size_t)u * a.max_stamps;
                const bool out16 = ((unsigned long long)out & 15ull) == 0;
                for (int g = lane; g < N && g < a.max_stamps; g += 32) {
                    const int2 r = se[g];
                    BfaStamp sv;
                    sv.phoneme = my_cls[g]; sv.start = r.x; sv.end = r.y; sv.target_idx = it.idx0 + g;
                    if (out16) *reinterpret_cast<int4*>(out + g) = make_int4(sv.phoneme, sv.start, sv.end, sv.target_idx);
                    else out[g] = sv;
                    if (want_conf) a.conf[(size_t)u * a.max_stamps + g] = stamp_confidence_prob(lp_s, T, r.x, r.y);
                }
                if (N > a.max_stamps) st |= BFA_ST_STAMP_OVERFLOW;
            }
            if (lane == 0) {
                if (want_stamps) a.n_stamps[u] = min(N, a.max_stamps);
                a.status[u] = st;
                if (a.dp_final) a.dp_final[u] = finv[2 * seg + 1];
                if (a.uflag) a.uflag[u] = 1;
            }
        } else if (lane == 0) {
            if (a.deferred) {
                a.deferred[atomicAdd(a.n_deferred, 1)] = u;
                if (a.uflag) a.uflag[u] = 0;
            } else {
                a.status[u] = BFA_ST_DEFERRED;
                if (want_stamps) a.n_stamps[u] = 0;
                if (a.dp_final) a.dp_final[u] = 0.0f;
            }
        }
        __syncwarp();
        FIN_T(8);
    }
    FIN_FLUSH;
}

// Role and pair of a warp.  A DP warp issues about 1.45x the instructions of a helper warp, and warp w issues on scheduler w % 4.
// With 14 warps the schedulers hold 4, 4, 3, 3 of them; giving the two 4-warp schedulers one DP warp + three helpers each and
// the 3-warp schedulers 2 + 1 and 3 + 0 keeps the busiest scheduler ~9 % lighter than any contiguous split.
__device__ __forceinline__ void band3_roles(int warp, int npairs, bool& is_dp, int& pair, bool& idle) {
    is_dp = warp < npairs;                    // default: warps [0, npairs) run the DP, [npairs, 2 npairs) are their helpers
    pair = is_dp ? warp : warp - npairs;
    idle = warp >= 2 * npairs;
#ifndef BFA_ROLES_CONTIG
    if (npairs == 7) {
#ifdef BFA_DP_WARPS
        constexpr uint32_t DP_WARPS = BFA_DP_WARPS;
#else
        constexpr uint32_t DP_WARPS = (1u << 0) | (1u << 1) | (1u << 2) | (1u << 3) | (1u << 6) | (1u << 7) | (1u << 11);
#endif
        is_dp = (DP_WARPS >> warp) & 1u;
        pair = __popc((is_dp ? DP_WARPS : ~DP_WARPS) & ((1u << warp) - 1u));
        idle = warp >= 14;
    }
#endif
}

// cta / ncta: this CTA's rank among the CTAs that work on the class, and their number (uniform over the CTA); rev: hand the list
// out back to front.
template <int G, int CT, bool EXACT>
__device__ void band3_run(const Band3Args& a, const Band3Args::Class& k, unsigned char* smem_raw, int warp, int lane, int cta, int ncta, bool rev) {
    const int n_items = *k.n_items;
    if (n_items == 0) return;                 // nothing of this window class in the batch (uniform over the CTA)
    const int npairs = min((int)(blockDim.x >> 6), k.npairs);   // as many pairs as the shared memory of one SM holds
    bool is_dp, idle;
    int pair;
    band3_roles(warp, npairs, is_dp, pair, idle);
    unsigned char* smem_pair = smem_raw + (size_t)pair * k.smem_per_warp;
    __syncthreads();                          // the previous class is done with the shared memory
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // its generic-proxy writes are ordered before this class's bulk copies
    if (!idle && is_dp && lane == 0) {
        // No zero fill: slots that are never loaded (rows past the end of an utterance, unused segments) may hold
        // anything, NaN included; whatever is computed from them is never looked at.
        const uint32_t b0 = smem_u32(smem_pair + band3_off_bars((size_t)k.region));
        for (int i = 0; i < B3_NBARS; ++i) mbar_init(b0 + 8u * i, band3_bar_count(i));
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    const int n_tasks = (n_items + B3_UPW - 1) / B3_UPW;
    // static deal, round by round: a pair is slot (cta + ncta * pair) of n_slots; even rounds hand their tasks out in slot order,
    // odd rounds in reverse (a snake): when the list is ordered by length -- a length-bucketing caller, handed out longest first --
    // the slot that got the longest task of one round gets the shortest of the next, and every slot ends up with about the same
    // number of frames.  With one CTA per SM consecutive slots are different SMs, so a round spreads evenly over the SMs.
    const int n_slots = ncta * npairs, slot = cta + ncta * pair;
    if (idle) {
    } else if (is_dp) {
        uint32_t* slab = k.bp_scratch + (size_t)(blockIdx.x * npairs + pair) * k.bp_slab_words;
        for (int r = 0; r * n_slots < n_tasks; ++r) {
            const int q = r * n_slots + ((r & 1) ? n_slots - 1 - slot : slot);
            if (q >= n_tasks) continue;
            const int j = rev ? n_tasks - 1 - q : q;
            band3_dp<G, CT, EXACT, false>(a, k, j * B3_UPW, min(B3_UPW, n_items - j * B3_UPW), smem_pair, slab, phase, lane, pair, nullptr, nullptr);
        }
    } else {
        const uint64_t pol = policy_evict_first();
        bool not_first = false;
        for (int r = 0; r * n_slots < n_tasks; ++r) {
            const int q = r * n_slots + ((r & 1) ? n_slots - 1 - slot : slot);
            if (q >= n_tasks) continue;
            const int j = rev ? n_tasks - 1 - q : q;
            band3_helper<G, CT, false>(a, k, j * B3_UPW, min(B3_UPW, n_items - j * B3_UPW), smem_pair, phase, not_first, lane, pol, pair, nullptr, nullptr);
            not_first = true;
        }
    }
    __syncthreads();                          // every pair of this CTA is done with this class
    if (!idle && is_dp && lane == 0) {        // the barrier words are about to become ordinary shared memory again
        const uint32_t b0 = smem_u32(smem_pair + band3_off_bars((size_t)k.region));
        for (int i = 0; i < B3_NBARS; ++i) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(b0 + 8u * i) : "memory");
    }
}

// The launch: all three window classes (24 / 40 / 64 groups per window), one after the other inside every CTA.
template <int CT, bool EXACT>
__global__ void __launch_bounds__(B3_PAIRS * 64, 1) viterbi_band3_kernel(Band3Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_release();
    pdl_wait();                               // the planner's item lists
    // One class in the batch (the usual case): every CTA works on it.  Several: the classes run SIDE BY SIDE, each on a share of
    // the CTAs proportional to its work (frames x cost of a chunk at that window width / pairs an SM holds at that width) --
    // run one after the other each class costs about one longest task, however few tasks it has.
    const int n0 = *a.cls[0].n_items, n1 = *a.cls[1].n_items, n2 = *a.cls[2].n_items;
    const int G_ = (int)gridDim.x, c_ = (int)blockIdx.x;
    // every list is handed out from its longer end (a caller that orders utterances by length, either way, gets longest-first)
    auto from_back = [](const Band3Args::Class& k, int n) { return n > 1 && k.items[0].T < k.items[n - 1].T; };
    const bool r0 = from_back(a.cls[0], n0), r1 = from_back(a.cls[1], n1), r2 = from_back(a.cls[2], n2);
    if ((n0 > 0) + (n1 > 0) + (n2 > 0) <= 1 || G_ < 6) {
        band3_run<3, CT, EXACT>(a, a.cls[0], smem_raw, warp, lane, c_, G_, r0);
        band3_run<5, CT, EXACT>(a, a.cls[1], smem_raw, warp, lane, c_, G_, r1);
        band3_run<8, CT, EXACT>(a, a.cls[2], smem_raw, warp, lane, c_, G_, r2);
        return;
    }
    const int np = (int)(blockDim.x >> 6);
    const float w0 = n0 ? fmaxf(1.f, (float)*a.cls[0].n_frames) * 1.0f / (float)min(np, a.cls[0].npairs) : 0.f;
    const float w1 = n1 ? fmaxf(1.f, (float)*a.cls[1].n_frames) * 1.6f / (float)min(np, a.cls[1].npairs) : 0.f;
    const float w2 = n2 ? fmaxf(1.f, (float)*a.cls[2].n_frames) * 2.5f / (float)min(np, a.cls[2].npairs) : 0.f;
    const float ws = w0 + w1 + w2;
    int g0 = n0 ? max(1, (int)(G_ * w0 / ws + 0.5f)) : 0;
    int g1 = n1 ? max(1, (int)(G_ * w1 / ws + 0.5f)) : 0;
    int g2 = n2 ? max(1, (int)(G_ * w2 / ws + 0.5f)) : 0;
    // make the shares add up: take from / give to the largest
    int* gl = (g0 >= g1 && g0 >= g2) ? &g0 : (g1 >= g2 ? &g1 : &g2);
    *gl += G_ - (g0 + g1 + g2);
    if (c_ < g0) band3_run<3, CT, EXACT>(a, a.cls[0], smem_raw, warp, lane, c_, g0, r0);
    else if (c_ < g0 + g1) band3_run<5, CT, EXACT>(a, a.cls[1], smem_raw, warp, lane, c_ - g0, g1, r1);
    else band3_run<8, CT, EXACT>(a, a.cls[2], smem_raw, warp, lane, c_ - g0 - g1, g2, r2);
}

// Direct mode: ONE kernel per batch on the common path.  Task q = utterances 4q .. 4q+3, planned by the helper warp itself,
// 24-group window; fill, walk, frame labels, timestamps and confidences all happen here; whatever does not qualify is handed to
// the planner chain (or flagged).  The kernel reads nothing another kernel of the call produces, so it never waits for its
// predecessor in the stream before the fill: launched with programmatic stream serialization it starts, SM by SM, while the
// previous launch (normally the previous batch's instance of this kernel) drains, and executes griddepcontrol.wait only before
// its first write another grid could see (band3_direct_finish).  The fill's private scratch (decision records, confidence
// inputs) is indexed by the PHYSICAL SM: a CTA of this kernel owns its SM's shared memory, so two launches never use the same
// SM's scratch at the same time.
template <int CT, bool EXACT, bool LOGITS = false>
__global__ void __launch_bounds__(B3_PAIRS * 64, 1) viterbi_band3_direct_kernel(const __grid_constant__ Band3Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_release();
    const Band3Args::Class& k = a.cls[0];
    const int npairs = min((int)(blockDim.x >> 6), k.npairs);
    bool is_dp, idle;
    int pair;
    band3_roles(warp, npairs, is_dp, pair, idle);
    unsigned char* smem_pair = smem_raw + (size_t)pair * k.smem_per_warp;
    const int n_tasks = (a.B + B3_UPW - 1) / B3_UPW;
    if (!idle && is_dp && lane == 0) {
        const uint32_t b0 = smem_u32(smem_pair + band3_off_bars((size_t)k.region));
        for (int i = 0; i < B3_NBARS; ++i) mbar_init(b0 + 8u * i, band3_bar_count(i));
        fence_mbar_init();
    }
    __syncthreads();
    if (idle) return;
    unsigned smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    if ((int)smid >= a.nslots) __trap();      // the host sized the per-SM scratch from %nsmid
    const size_t slot = (size_t)smid * npairs + pair;
    uint32_t* slab = k.bp_scratch + slot * k.bp_slab_words;
    float* pscr_lp = a.pscr_lp ? a.pscr_lp + slot * B3_UPW * a.tpitch : nullptr;
    unsigned char* pscr_gs = a.pscr_lp ? a.pscr_gs + slot * B3_UPW * a.tpitch : nullptr;
    uint32_t phase = 0;
    if (is_dp) {
        for (int q = blockIdx.x + gridDim.x * pair; q < n_tasks; q += gridDim.x * npairs)
            band3_dp<3, CT, EXACT, true, LOGITS>(a, k, q, B3_UPW, smem_pair, slab, phase, lane, pair, pscr_lp, pscr_gs);
    } else {
        const uint64_t pol = policy_evict_first();
        bool not_first = false;
        for (int q = blockIdx.x + gridDim.x * pair; q < n_tasks; q += gridDim.x * npairs) {
            band3_helper<3, CT, true, LOGITS>(a, k, q, B3_UPW, smem_pair, phase, not_first, lane, pol, pair, pscr_lp, pscr_gs);
            not_first = true;
        }
    }
}

}  // namespace bfa
