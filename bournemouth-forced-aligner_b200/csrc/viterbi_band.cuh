// viterbi_band.cuh -- the hot kernel: banded stride-4 Viterbi fill + back-trace, 4 utterances per warp.
//
// Same DP as viterbi_generic.cuh (forced_alignment.py:563-703) specialised for the shape that
// dominates real work: structured stride-4 path  b,(p,b,b,b)xN  (forced_alignment.py:153, :181-187)
// with a Sakoe-Chiba band (:190, :441) or a path short enough to fit the window.
//
// Mapping.  States are grouped as g' = (s+3)/4 -> (p, b1, b2, b3); group 0 holds only state 0 in its
// b3 slot.  An 8-lane segment owns one utterance; lane l of the segment holds G consecutive groups
// of a sliding window of W = 8*G groups that follows the band.  A warp therefore advances 4
// utterances in lock-step; per frame a lane needs only 2 shuffles (left neighbour's b2,b3), one
// blank emission and G phoneme emissions.  Because all candidates of a state share its emission,
// a group costs 7 FADD + 6 FMNMX; each of its 6 arg-max decisions is one FADD (sign of the
// difference) + one funnel shift that pushes the sign bit into a per-decision accumulator, so
// back-pointers are 6 bits per group per frame and are flushed as whole words every 32 frames.
// Frame 0 is not special: the DP starts from a virtual frame -1 with dp[state 0] = 0, everything
// else invalid, which yields dp[0][0] = lp[0][blank], dp[0][1] = lp[0][path[1]] (:594-596).
//
// Band.  The reference re-masks every out-of-band state to exactly -1000 on every frame (:650-653).
// Here out-of-band states only have to be INVALID: they are set to -inf, which is absorbing under
// "+ emission" and loses every comparison, so an invalid state stays invalid without being re-masked
// and can never beat a valid candidate.  (A path that dips to <= -1000 without being masked is
// "invalid" for the reference but alive here; emissions are <= 0, so such a path also ENDS <= -1000
// and is caught by the final validity test.)  That makes masking event-driven: when the lower edge
// passes a state it is killed once; when the upper edge admits a new state everything above the old
// edge is killed first.  The band schedule (fp64 centre, fp32 limits, exactly as :651-652) is
// evaluated once per 8 frames, one frame per lane, and turned into per-frame event bit masks, so a
// frame without an event pays one uniform bit test.  Anything that ends on an invalid state
// (degenerate input, where the reference's exact -1000 bookkeeping matters) is not finished here:
// the item is appended to the retry list and re-run by the exact generic kernel.
//
// Fusion.  The target boost + log_softmax + floor of :121-129 is applied on the fly: the 8 lanes of a
// segment compute the row's log-sum-exp from the staged raw row (fixed shift = boost instead of the
// row max -- valid for log-probability inputs; a non-finite sum sends the item to the exact path), so
// the posteriors are read from HBM exactly once, through 1-D bulk async copies (TMA engine, SASS
// UBLKCP) of 8 rows per utterance per stage, double buffered, one mbarrier per (stage, utterance).
#pragma once
#include <type_traits>

#include "bfa_common.cuh"

namespace bfa {

// LPU lanes own one utterance (template parameter: 16 -> 2 utterances per warp, 8 -> 4).  A warp-task is latency
// bound (one task takes the same time whether the SM holds 2 or 14 of them), so the default is the small task:
// LPU = 16 halves the instructions per warp-frame and doubles the number of independent warps.
constexpr int BK_NST = 2;         // pipeline stages
constexpr int BK_ROWS = 8;        // rows per stage per utterance (8*C*4 bytes is always a multiple of 16)
constexpr int BK_PAD = 128;       // zeroed floats after the last stage buffer (LSE lanes may read past a row)

struct BandArgs {
    BfaParams p;
    int C;
    const float* logp;
    const int32_t* tgt;
    const uint32_t* tmask;     // [B][MAX_WORDS]
    const Item* items;         // fast list
    const int* n_items;
    Item* retry_items;         // generic list: items this kernel could not finish are appended
    int* n_retry;
    int32_t* frame_ph;
    int32_t* frame_idx;
    float* dp_final;
    float* path_lp;            // [total_frames] raw log-prob of the assigned class per frame (confidence input), or null
    uint32_t* bp_scratch;
    long long bp_slab_words;
    int seg_stride;            // floats per (stage, utterance) buffer = BK_ROWS * C
    int smem_per_warp;         // bytes
};

template <int LPU, int G>
struct BandShape {
    static constexpr int UPW = 32 / LPU;    // utterances per warp
    static constexpr int W = LPU * G;       // groups in the window
    static constexpr int ACC = 6 * G;       // decision accumulators per lane
    static constexpr int REC = ACC + 1;     // + window-shift flag word
    static constexpr int FPL = 32 / LPU;    // frames of a 32-frame block each lane writes out
    static constexpr int WARPS = (LPU == 16) ? 14 : 8;   // warps per CTA, one CTA per SM
};

// window eligibility: groups spanned by band(t-1) U band(t) must fit (see header comment)
__host__ __device__ inline bool band_window_fits(int N, int band, int W) {
    if (band <= 0) return N + 1 <= W;
    return (2 * band + 5) / 4 + 1 <= W || N + 1 <= W;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// push (later > earlier) into acc: the sign bit of (earlier - later) is 1 iff later > earlier (ties -> 0)
__device__ __forceinline__ void push_gt(uint32_t& acc, float earlier, float later) {
    acc = __funnelshift_l(__float_as_uint(earlier - later), acc, 1);
}

template <int LPU, int G, int NI>
__device__ void band_task(const BandArgs& a, int first, int n_valid, unsigned char* smem_warp, uint32_t* slab, uint32_t& phase,
                          int lane, uint64_t pol) {
    using S = BandShape<LPU, G>;
    constexpr int BK_LPU = LPU, BK_UPW = S::UPW;
    constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    constexpr int SENT = 1 << 29;
    const int seg = lane / LPU, l8 = lane % LPU;   // l8: lane inside the segment (0..LPU-1)
    const int C = a.C;
    const float NEG = a.p.neg_inf;
    const int blank = a.p.blank_id;

    float* stage_buf = reinterpret_cast<float*>(smem_warp);
    const int stage_floats = BK_UPW * a.seg_stride;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(stage_buf + BK_NST * stage_floats + BK_PAD);   // [NST][UPW]
    uint32_t* bt = reinterpret_cast<uint32_t*>(bars + BK_NST * BK_UPW);                                               // [REC][32]

    // ---- per-segment item description (uniform within a segment) ----
    const bool seg_on = seg < n_valid;
    const Item& it = a.items[first + (seg_on ? seg : 0)];
    const int T = seg_on ? it.T : 0;
    const int N = it.n, L = it.L, band = it.band, flags = it.flags, utt = it.utt;
    const int trim = it.trim, n_out = it.n_out, idx0 = it.idx0;
    const long long out_off = it.out_off, out_lim = it.out_lim;
    const bool use_band = band > 0 && T > 1 && L > 1;                          // :586
    const double pace = use_band ? (double)(L - 1) / (double)(T - 1) : 0.0;   // :587
    const double bandd = (double)band;
    const bool use_stats = (flags & ITEM_STATS) != 0;
    const bool warp_stats = __any_sync(FULL, use_stats);
    const float min_lp = (flags & ITEM_FLOOR) ? a.p.min_log_prob : -INFINITY;
    const int32_t* seq = a.tgt + it.seq_off;
    const int base_max = max(0, N + 1 - S::W);
    const float* my_src = a.logp + it.lp_off;
    int Tq[BK_UPW];   // frames of each of the four utterances (warp-uniform)
    int Tmax = 0;
#pragma unroll
    for (int q = 0; q < BK_UPW; ++q) {
        Tq[q] = __shfl_sync(FULL, T, q * BK_LPU);
        Tmax = max(Tmax, Tq[q]);
    }
    const int n_chunks = (Tmax + BK_ROWS - 1) / BK_ROWS;
    const int C4 = C * 4;

    // lanes with l8 == 0 issue their own utterance's copy; every lane of the segment waits on its barrier
    auto issue = [&](int c) {
        const int rows = min(BK_ROWS, T - c * BK_ROWS);
        if (l8 == 0 && rows > 0) {
            const int st = c & 1;
            const uint32_t bar = smem_u32(&bars[st * BK_UPW + seg]);
            float* dst = stage_buf + st * stage_floats + seg * a.seg_stride;
            const float* s = my_src + (size_t)c * BK_ROWS * C;
            const uint32_t bytes = (uint32_t)rows * C4, bulk = bytes & ~15u;
            mbar_expect_tx(bar, bulk);
            if (bulk) bulk_g2s_hint(smem_u32(dst), s, bulk, bar, pol);
            for (uint32_t w = bulk >> 2; w < (bytes >> 2); ++w) dst[w] = s[w];   // < 4 tail floats of a partial last chunk
        }
    };

    // ---- per-lane window state ----
    int base = 0;                      // first group of the window
    float P[G], B1[G], B2[G], B3[G];   // dp of the lane's groups; -inf = invalid
    int crel[G];                       // (class of the group) - l8 : float offset relative to the lane's row pointer
    uint32_t acc[S::ACC];
    uint32_t shift_acc = 0;
#pragma unroll
    for (int i = 0; i < S::ACC; ++i) acc[i] = 0;
    auto group_class = [&](int gi) { return (gi >= 1 && gi <= N) ? seq[gi - 1] : blank; };
#pragma unroll
    for (int g = 0; g < G; ++g) {
        P[g] = B1[g] = B2[g] = B3[g] = -INFINITY;
        crel[g] = (seg_on ? group_class(l8 * G + g) : blank) - l8;
    }
    if (l8 == 0) B3[0] = 0.0f;        // virtual frame -1: only state 0 is alive, with score 0
    int next_crel = (seg_on ? group_class(S::W) : blank) - l8;   // class of the group that enters at the next shift (lane 7)
    const int brel = blank - l8;
    const bool seg_first = l8 == 0;

    // ---- log-sum-exp constants: this lane sums classes l8 + LPU*i; out-of-range classes get weight 0 ----
    float kk[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int c = l8 + LPU * i;
        const bool ok = c < C;
        const bool tg = ok && seg_on && ((a.tmask[(size_t)utt * MAX_WORDS + (c >> 5)] >> (c & 31)) & 1u);
        kk[i] = ok ? (tg ? 0.0f : -a.p.boost_factor * LOG2E) : -INFINITY;   // (x + b - boost) * log2(e)
    }
    const float boostv = a.p.boost_factor;

    issue(0);

    int carry_lo1 = -SENT, carry_lo2 = -SENT, carry_hi1 = SENT;   // s_lo(t0-1), s_lo(t0-2), s_hi(t0-1)
    bool bad = false;                                              // needs the exact path
    float fin_val = NEG;
    int fin_g = 0, fin_k = 3, fin_base = 0;
    float emax = -INFINITY;                                        // raw mode: emissions must be <= 0 (log-probabilities)
    float lse_chk = 0.f;                                           // running sum of the rows' log-sum-exp (finite <=> all rows sane)

    for (int c = 0; c < n_chunks; ++c) {
        const int t0 = c * BK_ROWS;
        __syncwarp();                                  // every lane is done with the other stage (chunk c-1)
        if (c + 1 < n_chunks) issue(c + 1);
        // Every lane polls all four utterance barriers: the wait stays warp-uniform, so the segments never split
        // into independently scheduled sub-warps (which would quadruple the issue slots of everything below).
#pragma unroll
        for (int q = 0; q < BK_UPW; ++q) {
            if (t0 < Tq[q]) {                          // utterance q has rows in chunk c
                const int bi = (c & 1) * BK_UPW + q;
                mbar_wait(smem_u32(&bars[bi]), (phase >> bi) & 1u);
                phase ^= 1u << bi;
            }
        }
        __syncwarp();                                  // tail floats written by the issuing lane become visible

        // ---- band schedule of frames t0..t0+7 (:650-652), frame t0+l8 on lane l8, and the event masks ----
        const int r8 = l8 & 7;            // lanes 8..15 of a 16-lane segment mirror lanes 0..7 (their bits are masked out)
        const int j = t0 + r8;
        int s_lo = -SENT, s_hi = SENT;
        if (use_band) {
            const double center = (double)j * pace;
            s_lo = __float2int_ru((float)(center - bandd));   // smallest s with (float)s >= lo
            s_hi = __float2int_rd((float)(center + bandd));   // largest  s with (float)s <= hi
        }
        int p1_lo = __shfl_up_sync(FULL, s_lo, 1), p2_lo = __shfl_up_sync(FULL, s_lo, 2), p1_hi = __shfl_up_sync(FULL, s_hi, 1);
        if (r8 == 0) { p1_lo = carry_lo1; p2_lo = carry_lo2; p1_hi = carry_hi1; }
        if (r8 == 1) p2_lo = carry_lo1;
        // kill before frame j: the lower edge moved at j-1, or the upper edge rises at j
        const bool kill_j = use_band && j >= 1 && ((j >= 2 && p1_lo > p2_lo) || s_hi > p1_hi);
        // window base used for frame j = group of s_lo(j-1), clamped
        const int want_j = (use_band && j >= 1) ? min(max((p1_lo + 3) >> 2, 0), base_max) : 0;
        const int want_m = (use_band && j >= 2) ? min(max((p2_lo + 3) >> 2, 0), base_max) : 0;
        const bool shift_j = want_j > want_m;
        const unsigned kb = __ballot_sync(FULL, kill_j), sb = __ballot_sync(FULL, shift_j);
        const unsigned my_kill8 = (kb >> (seg * LPU)) & 0xffu, my_shift8 = (sb >> (seg * LPU)) & 0xffu;
        unsigned ev8 = 0;                                          // frames of this chunk in which any segment has an event
#pragma unroll
        for (int q = 0; q < BK_UPW; ++q) ev8 |= ((kb | sb) >> (q * LPU)) & 0xffu;
        shift_acc = (shift_acc << 8) | (__brev(my_shift8) >> 24);  // frame order: t0 in the highest of the 8 bits
        // s_lo(t-1) / s_hi(t-1) as seen by frame r of this chunk live on lane r of the segment (p1_*)
        carry_lo1 = __shfl_sync(FULL, s_lo, seg * LPU + 7);
        carry_lo2 = __shfl_sync(FULL, s_lo, seg * LPU + 6);
        carry_hi1 = __shfl_sync(FULL, s_hi, seg * LPU + 7);

        const float* rowp0 = stage_buf + (c & 1) * stage_floats + seg * a.seg_stride + l8;   // lane's pointer into row 0
        const int fin_r = T - 1 - t0;                                                         // row of the last frame if it is in this chunk
        const bool fin_here = __any_sync(FULL, fin_r >= 0 && fin_r < BK_ROWS);

        // ---- fused boost + log_softmax statistics (:51-54) of the 8 rows, up front: 72 independent exps per lane,
        //      tree sums and batched shuffles instead of one dependent chain per frame ----
        float lnS8[BK_ROWS];
#pragma unroll
        for (int r = 0; r < BK_ROWS; ++r) lnS8[r] = 0.f;
        if (warp_stats) {   // uniform: every item of a call shares the mode
#pragma unroll
            for (int r = 0; r < BK_ROWS; ++r) {
                const float* rp = rowp0 + r * C;
                float e[NI];
#pragma unroll
                for (int i = 0; i < NI; ++i) e[i] = ex2_approx(fmaf(rp[LPU * i], LOG2E, kk[i]));
#pragma unroll
                for (int w = 1; w < NI; w <<= 1)
#pragma unroll
                    for (int i = 0; i + w < NI; i += 2 * w) e[i] += e[i + w];
                lnS8[r] = e[0];
            }
#pragma unroll
            for (int d = 1; d < BK_LPU; d <<= 1)
#pragma unroll
                for (int r = 0; r < BK_ROWS; ++r) lnS8[r] += __shfl_xor_sync(FULL, lnS8[r], d);
#pragma unroll
            for (int r = 0; r < BK_ROWS; ++r) {
                lnS8[r] = lg2_approx(lnS8[r]) * LN2;   // log sum exp(x + b - boost)
                lse_chk += lnS8[r];                    // any zero / overflowing / NaN sum leaves a non-finite trace here
            }
        }
        // raw emissions of the next frame are fetched one frame ahead
        float xb_n = rowp0[brel], xp_n[G];
#pragma unroll
        for (int g = 0; g < G; ++g) xp_n[g] = rowp0[crel[g]];

        auto frame = [&](const int r, auto check_fin) {
            constexpr bool CHECK = decltype(check_fin)::value;
            const float* rowp = rowp0 + r * C;
            const bool active = CHECK ? (r <= fin_r) : (t0 < T);
            float lnS = lnS8[0];
            if (CHECK) {
#pragma unroll
                for (int q = 1; q < BK_ROWS; ++q) lnS = (r == q) ? lnS8[q] : lnS;
            } else {
                lnS = lnS8[r];
            }
            const float lse = warp_stats ? lnS + boostv : 0.f;
            // emissions: blank is never a target (x - lse); phoneme classes are boosted targets (x + boost - lse = x - lnS)
            const float eb = xb_n - lse;
            float ep[G];
#pragma unroll
            for (int g = 0; g < G; ++g) ep[g] = fmaxf(xp_n[g] - lnS, min_lp);
            if (r + 1 < BK_ROWS) {
                xb_n = rowp[C + brel];
#pragma unroll
                for (int g = 0; g < G; ++g) xp_n[g] = rowp[C + crel[g]];
            }
            if (!warp_stats) {
                float m = eb;
#pragma unroll
                for (int g = 0; g < G; ++g) m = fmaxf(m, ep[g]);
                if (active) emax = fmaxf(emax, m);
            }

            // ---- events (rare): kill cells outside band(t-1); slide the window to the group of s_lo(t-1) ----
            if (ev8 & (1u << r)) {
                const int slo_p = __shfl_sync(FULL, p1_lo, seg * LPU + r);
                const int shi_p = __shfl_sync(FULL, p1_hi, seg * LPU + r);
                if ((my_kill8 >> r) & 1u) {
                    const unsigned span = (unsigned)(shi_p - slo_p);
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const int s0 = 4 * (base + l8 * G + g) - 3 - slo_p;
                        if ((unsigned)(s0 + 0) > span) P[g] = -INFINITY;
                        if ((unsigned)(s0 + 1) > span) B1[g] = -INFINITY;
                        if ((unsigned)(s0 + 2) > span) B2[g] = -INFINITY;
                        if ((unsigned)(s0 + 3) > span) B3[g] = -INFINITY;
                    }
                }
                const bool shift_me = (my_shift8 >> r) & 1u;
                const float nP = __shfl_down_sync(FULL, P[0], 1), n1 = __shfl_down_sync(FULL, B1[0], 1);
                const float n2 = __shfl_down_sync(FULL, B2[0], 1), n3 = __shfl_down_sync(FULL, B3[0], 1);
                const int nc = __shfl_down_sync(FULL, crel[0], 1) + 1;   // neighbour's offset is relative to lane l8+1
                if (shift_me) {
#pragma unroll
                    for (int g = 0; g + 1 < G; ++g) { P[g] = P[g + 1]; B1[g] = B1[g + 1]; B2[g] = B2[g + 1]; B3[g] = B3[g + 1]; crel[g] = crel[g + 1]; }
                    const bool last = l8 == BK_LPU - 1;
                    P[G - 1] = last ? -INFINITY : nP; B1[G - 1] = last ? -INFINITY : n1;
                    B2[G - 1] = last ? -INFINITY : n2; B3[G - 1] = last ? -INFINITY : n3;
                    crel[G - 1] = last ? next_crel : nc;
                    base += 1;
                    if (last) next_crel = group_class(base + S::W) - l8;
#pragma unroll
                    for (int g = 0; g < G; ++g) {   // emissions follow their groups (this frame and the prefetched next one)
                        ep[g] = fmaxf(rowp[crel[g]] - lnS, min_lp);
                        if (r + 1 < BK_ROWS) xp_n[g] = rowp[C + crel[g]];
                    }
                }
            }

            // ---- DP update, right-most group first so that left neighbours are still frame t-1 ----
            float l2 = __shfl_up_sync(FULL, B2[G - 1], 1), l3 = __shfl_up_sync(FULL, B3[G - 1], 1);
            if (seg_first) { l2 = -INFINITY; l3 = -INFINITY; }   // nothing (or a dropped, invalid group) to the left
#pragma unroll
            for (int g = G - 1; g >= 0; --g) {
                const float Lb2 = (g > 0) ? B2[(g > 0) ? g - 1 : 0] : l2;
                const float Lb3 = (g > 0) ? B3[(g > 0) ? g - 1 : 0] : l3;
                const float c0 = P[g] + ep[g], c1 = Lb3 + ep[g], c2 = Lb2 + ep[g];   // stay / advance / skip into p
                const float PB = P[g] + eb, S1 = B1[g] + eb, S2 = B2[g] + eb, S3 = B3[g] + eb;
                uint32_t* A = &acc[6 * g];
                // p   : first max of (c0, c1, c2)           (:645)
                const float m01 = fmaxf(c0, c1);
                push_gt(A[0], c0, c1);
                push_gt(A[1], m01, c2);
                P[g] = fmaxf(m01, c2);
                // b1  : (stay S1, advance PB); the skip candidate is masked (blank == blank, :604)
                push_gt(A[2], S1, PB);
                B1[g] = fmaxf(S1, PB);
                // b2  : (stay S2, advance S1, skip PB)
                const float m21 = fmaxf(S2, S1);
                push_gt(A[3], S2, S1);
                push_gt(A[4], m21, PB);
                B2[g] = fmaxf(m21, PB);
                // b3  : (stay S3, advance S2)
                push_gt(A[5], S3, S2);
                B3[g] = fmaxf(S3, S2);
            }

            if (CHECK) {
                const bool fin = (r == fin_r);
                const unsigned fin_mask = __ballot_sync(FULL, fin);
                if (fin) {
                    const int t = t0 + r;
                    // last record of this utterance, left-aligned so that frame 32b+q sits at bit 31-q
                    const int sh = 31 - (t & 31);
                    uint32_t* rec = slab + (size_t)(t >> 5) * S::REC * 32 + lane;
#pragma unroll
                    for (int i = 0; i < S::ACC; ++i) rec[i * 32] = acc[i] << sh;
                    rec[S::ACC * 32] = shift_acc << (24 - 8 * (c & 3));
                    // ---- final state (:656-682) from the window at frame T-1 ----
                    float bv = -INFINITY;
                    int bs = -1;
                    if (!a.p.truly_forced) {
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            const int s0 = 4 * (base + l8 * G + g) - 3;
                            if (s0 >= 0 && s0 < L && P[g] > NEG && (bs < 0 || P[g] > bv)) { bv = P[g]; bs = s0; }
                            if (s0 + 1 >= 0 && s0 + 1 < L && B1[g] > NEG && (bs < 0 || B1[g] > bv)) { bv = B1[g]; bs = s0 + 1; }
                            if (s0 + 2 >= 0 && s0 + 2 < L && B2[g] > NEG && (bs < 0 || B2[g] > bv)) { bv = B2[g]; bs = s0 + 2; }
                            if (s0 + 3 >= 0 && s0 + 3 < L && B3[g] > NEG && (bs < 0 || B3[g] > bv)) { bv = B3[g]; bs = s0 + 3; }
                        }
                    } else {
                        // L-1 = 4N is b3 of group N, L-2 its b2
                        const int w = N - base;
                        const int lw = w / G, sl = w - lw * G;
                        if (w >= 0 && w < S::W && lw == l8) {
#pragma unroll
                            for (int g = 0; g < G; ++g)
                                if (g == sl) {
                                    if (B3[g] > NEG) { bv = B3[g]; bs = L - 1; }
                                    else if (L >= 2 && B2[g] > NEG) { bv = B2[g]; bs = L - 2; }
                                }
                        }
                    }
                    // segment reduction: max value, ties -> lower state
#pragma unroll
                    for (int d = 1; d < BK_LPU; d <<= 1) {
                        const float ov = __shfl_xor_sync(fin_mask, bv, d);
                        const int os = __shfl_xor_sync(fin_mask, bs, d);
                        if (os >= 0 && (bs < 0 || ov > bv || (ov == bv && os < bs))) { bv = ov; bs = os; }
                    }
                    if (bs < 0) bad = true;   // nothing valid at the end: degenerate -> exact path
                    fin_val = bv;
                    fin_g = (bs + 3) >> 2;
                    fin_k = (bs + 3) & 3;
                    fin_base = base;
                }
            }
        };

        if (!fin_here) {
#pragma unroll
            for (int r = 0; r < BK_ROWS; ++r) frame(r, std::false_type{});
        } else {
#pragma unroll 1
            for (int r = 0; r < BK_ROWS; ++r) frame(r, std::true_type{});
        }
        // ---- flush one full 32-frame record (utterances that end inside this chunk flushed at their last frame)
        if ((c & 3) == 3 && T - 1 > t0 + BK_ROWS - 1) {
            uint32_t* rec = slab + (size_t)(c >> 2) * S::REC * 32 + lane;
#pragma unroll
            for (int i = 0; i < S::ACC; ++i) rec[i * 32] = acc[i];
            rec[S::ACC * 32] = shift_acc;
        }
    }
    __syncwarp();
    if (emax > 0.f) bad = true;
    if (use_stats && !(fabsf(lse_chk) < 3.0e38f)) bad = true;
    {   // make `bad` uniform per segment
        const unsigned m = __ballot_sync(FULL, bad);
        bad = ((m >> (seg * BK_LPU)) & ((LPU == 32) ? 0xffffffffu : ((1u << LPU) - 1u))) != 0;
    }
    if (seg_on && l8 == 0) {
        if (bad) {
            const int slot = atomicAdd(a.n_retry, 1);
            a.retry_items[slot] = it;
        } else if ((flags & ITEM_FINAL) && a.dp_final) {
            a.dp_final[utt] = fin_val;
        }
    }

    // ---- back-trace (:686-703) ----
    // A cell is addressed by its window-relative state ci = 4*(g' - base) + k.  Staging lays the record
    // out as bt2[seg][ci] = (first decision word, second decision word or 0), so one 64-bit shared load
    // per frame yields (b0, b1) and every transition is  ci -= b1 ? 2 : b0  (p: skip = -2 / advance = -1,
    // b1: advance = -1, b2: skip = -2 / advance = -1, b3: advance = -1); a window shift adds 4.
    const bool walk = seg_on && !bad && T > 0;
    int ci = 4 * (fin_g - fin_base) + fin_k;
    int sabs0 = 4 * fin_base - 3;          // state of ci = 0 at the current frame
    long long pend_o[S::FPL];   // outputs of the previous 32-frame block, stored one block late
    int pend_cls[S::FPL], pend_idx[S::FPL];
    float pend_x[S::FPL];
    bool pend_gather[S::FPL];
#pragma unroll
    for (int i = 0; i < S::FPL; ++i) { pend_o[i] = -1; pend_cls[i] = 0; pend_idx[i] = 0; pend_x[i] = 0.f; pend_gather[i] = false; }
    uint2* bt2 = reinterpret_cast<uint2*>(bt);                 // [UPW][W*4] (b0 word, b1 word)
    uint32_t* sfw_s = bt + BK_UPW * S::W * 4 * 2;              // [32] shift-flag words
    const int nblk = (Tmax + 31) >> 5;
    for (int b = nblk - 1; b >= 0; --b) {
        __syncwarp();
        {
            const uint32_t* rec = slab + (size_t)b * S::REC * 32 + lane;
            uint32_t w[S::REC];
#pragma unroll
            for (int i = 0; i < S::REC; ++i) w[i] = rec[i * 32];
            uint4* dst = reinterpret_cast<uint4*>(bt2 + seg * S::W * 4 + l8 * G * 4);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                dst[2 * g + 0] = make_uint4(w[6 * g + 0], w[6 * g + 1], w[6 * g + 2], 0u);   // p: (A0, A1)   b1: (A2, 0)
                dst[2 * g + 1] = make_uint4(w[6 * g + 3], w[6 * g + 4], w[6 * g + 5], 0u);   // b2: (A3, A4)  b3: (A5, 0)
            }
            sfw_s[lane] = w[S::ACC];
        }
        __syncwarp();
        const uint32_t sfw = sfw_s[lane];
        const uint2* cell = bt2 + seg * S::W * 4;
        const int qhi = walk ? min(31, T - 1 - b * 32) : -1;   // last frame of this utterance inside the block
        int keep4[S::FPL];                                      // states of frames 32b + LPU*i + l8
#pragma unroll
        for (int i = 0; i < S::FPL; ++i) keep4[i] = 0;
#pragma unroll
        for (int q = 31; q >= 0; --q) {
            if (q <= qhi) {
                if ((q % LPU) == l8) keep4[q / LPU] = sabs0 + ci;
                if (q > 0 || b > 0) {                      // frame 0 has no predecessor
                    const uint2 wv = cell[ci];
                    const uint32_t bit = 1u << (31 - q);
                    const int d = (wv.y & bit) ? 2 : ((wv.x & bit) ? 1 : 0);
                    ci -= d;
                    if (sfw & bit) { ci += 4; sabs0 -= 4; }   // the window slid before this frame was computed
                }
            }
        }
        // ---- output of the block, software-pipelined by one block: the loads it needs (target ids, gathered
        //      log-probs for the confidences) are issued now, independent chains per lane, and consumed
        //      after the next block's walk, so their latency never stalls the warp ----
#pragma unroll
        for (int i = 0; i < S::FPL; ++i)
            if (pend_o[i] >= 0) {
                a.frame_ph[pend_o[i]] = pend_cls[i];
                a.frame_idx[pend_o[i]] = pend_idx[i];
                if (pend_gather[i]) a.path_lp[pend_o[i]] = pend_x[i];
            }
#pragma unroll
        for (int i = 0; i < S::FPL; ++i) {
            pend_o[i] = -1;
            const int tf = b * 32 + LPU * i + l8;
            if (walk && tf < T) {
                const int rel = tf - trim;
                const long long o = out_off + rel;
                if (rel >= 0 && rel < n_out && o < out_lim) {
                    const bool ph = ((keep4[i] + 3) & 3) == 0;      // state s = 4g'-3+k is a phoneme state iff k == 0
                    const int gi = (keep4[i] + 3) >> 2;
                    pend_o[i] = o;
                    pend_cls[i] = ph ? seq[gi - 1] : blank;
                    pend_idx[i] = ph ? idx0 + gi - 1 : -1;
                    pend_gather[i] = a.path_lp && (ph || !a.p.ignore_noise);
                }
            }
        }
        // confidences (utils.py:89-103) read lp[f, phoneme of the stamp]: gather it here, once per frame
#pragma unroll
        for (int i = 0; i < S::FPL; ++i)
            if (pend_o[i] >= 0 && pend_gather[i]) pend_x[i] = __ldg(my_src + (long long)(b * 32 + LPU * i + l8) * C + pend_cls[i]);
    }
#pragma unroll
    for (int i = 0; i < S::FPL; ++i)
        if (pend_o[i] >= 0) {
            a.frame_ph[pend_o[i]] = pend_cls[i];
            a.frame_idx[pend_o[i]] = pend_idx[i];
            if (pend_gather[i]) a.path_lp[pend_o[i]] = pend_x[i];
        }
    __syncwarp();
}

template <int LPU, int G, int NI>
__global__ void __launch_bounds__(BandShape<LPU, G>::WARPS * 32, 1) viterbi_band_kernel(BandArgs a) {
    using S = BandShape<LPU, G>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* smem_warp = smem_raw + (size_t)warp * a.smem_per_warp;
    {
        float* stage_buf = reinterpret_cast<float*>(smem_warp);
        const int nfl = BK_NST * S::UPW * a.seg_stride + BK_PAD;
        for (int i = lane; i < nfl; i += 32) stage_buf[i] = 0.0f;   // never-loaded slots must hold finite values
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(stage_buf + nfl);
        if (lane == 0)
            for (int i = 0; i < BK_NST * S::UPW; ++i) mbar_init(smem_u32(&bars[i]), 1);
        fence_mbar_init();   // also orders the generic-proxy zero fill before the first async copy
    }
    __syncwarp();
    const uint64_t pol = policy_evict_first();
    uint32_t phase = 0;
    const int gwarp = blockIdx.x * S::WARPS + warp;
    uint32_t* slab = a.bp_scratch + (size_t)gwarp * a.bp_slab_words;
    const int n_items = *a.n_items;
    const int n_tasks = (n_items + S::UPW - 1) / S::UPW;
    // static deal: task j -> CTA j % grid, warp (j / grid) % WARPS.  With one CTA per SM this spreads
    // ceil(n_tasks / SMs) tasks evenly over the SMs and over the four schedulers of each SM.
    for (int j = blockIdx.x + gridDim.x * warp; j < n_tasks; j += gridDim.x * S::WARPS)
        band_task<LPU, G, NI>(a, j * S::UPW, min(S::UPW, n_items - j * S::UPW), smem_warp, slab, phase, lane, pol);
}

}  // namespace bfa
