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
//    A group (p,m,b3) costs 6 FADD + 4 FMNMX and 4 decision bits (FADD sign + funnel shift) per frame instead
//    of 7 + 6 and 6 bits.
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
// 3. Fusion.  Target boost + log_softmax + floor (:121-129) are applied on the fly.  Rows are streamed exactly
//    once from HBM by 1-D bulk async copies (TMA engine, SASS UBLKCP; 8 rows per utterance per stage, double
//    buffered, one mbarrier per stage).  At the start of a chunk each lane computes the statistics of ONE staged
//    row (lane = (utterance, row): 66 exps, no shuffles, no idle lanes; per-utterance class weights come from a
//    small shared table) and publishes (log-sum-exp, blank emission) for the 8 frames that follow.
#pragma once
#include <type_traits>

#include "bfa_common.cuh"

namespace bfa {

constexpr int B3_LPU = 8;          // lanes per utterance
constexpr int B3_UPW = 4;          // utterances per warp
constexpr int B3_NST = 2;          // pipeline stages
constexpr int B3_ROWS = 8;         // rows per stage per utterance (8*C*4 bytes is always a multiple of 16)
constexpr int B3_WARPS = 8;        // warps per CTA, one CTA per SM
constexpr int B3_KK = 72;          // floats per utterance in the class-weight table (C <= 72)
constexpr int B3_STP = 9;          // float2 pitch of the per-utterance (lnS, eb) array (bank spreading)
constexpr int B3_MARGIN = 2;       // states of safety margin of the band-legality check

struct Band3Args {
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
    int seg_stride;            // floats per (stage, utterance) buffer = B3_ROWS * C
    int smem_per_warp;         // bytes
};

template <int G>
struct Band3Shape {
    static constexpr int W = B3_LPU * G;    // groups in the window
    static constexpr int ACC = 4 * G;       // decision accumulators per lane
    static constexpr int REC = ACC + 1;     // + window-slide word
    static constexpr int CELLS = 3 * W;     // back-trace cells per utterance
};

// Smallest window (in groups) that covers band(t-1) U band(t) for every frame of every 8-frame chunk when the
// window base is the group of the lower band edge at the frame before the chunk (scripts/check_window_fit.py
// verifies the closed form by brute force).
__host__ __device__ inline int band3_window_need(int N, int T, int L, int band) {
    if (band <= 0 || T < 2 || L < 2) return N + 1;
    const int adv = (int)(((long long)B3_ROWS * (L - 1) + (T - 2)) / (T - 1));   // ceil(8 * pace)
    const int w = (2 * band + adv + 5) / 4 + 1;
    return w < N + 1 ? w : N + 1;
}

__host__ __device__ inline size_t band3_smem_per_warp(int C, int G) {
    size_t b = (size_t)B3_NST * B3_UPW * B3_ROWS * C * 4;   // stage buffers
    b += (size_t)B3_UPW * B3_KK * 4;                        // class weights
    b += (size_t)B3_UPW * B3_STP * 8;                       // (lnS, eb) per staged row
    b += (size_t)B3_NST * 8;                                // mbarriers
    b = (b + 15) / 16 * 16;
    b += (size_t)B3_UPW * 3 * B3_LPU * G * 8;               // back-trace cells
    return (b + 127) / 128 * 128;
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// push (later > earlier) into acc: the sign bit of (earlier - later) is 1 iff later > earlier (ties -> 0)
__device__ __forceinline__ void b3_push(uint32_t& acc, float earlier, float later) {
    acc = __funnelshift_l(__float_as_uint(earlier - later), acc, 1);
}

template <int G, int CT>
__device__ void band3_task(const Band3Args& a, int first, int n_valid, unsigned char* smem_warp, uint32_t* slab, uint32_t& phase,
                           int lane, uint64_t pol) {
    using S = Band3Shape<G>;
    constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    const int seg = lane >> 3, l8 = lane & 7;
    const int C = CT ? CT : a.C;
    const float NEG = a.p.neg_inf;
    const int blank = a.p.blank_id;

    // ---- shared memory of this warp ----
    float* stage_buf = reinterpret_cast<float*>(smem_warp);
    const int seg_stride = B3_ROWS * C;
    const int stage_floats = B3_UPW * seg_stride;
    float* kk = stage_buf + B3_NST * stage_floats;                                         // [UPW][B3_KK]
    float2* stats = reinterpret_cast<float2*>(kk + B3_UPW * B3_KK);                        // [UPW][B3_STP]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(stats + B3_UPW * B3_STP);   // [NST]
    uint2* bt2 = reinterpret_cast<uint2*>(smem_warp + ((reinterpret_cast<unsigned char*>(bars + B3_NST) - smem_warp + 15) / 16) * 16);   // [UPW][CELLS]

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
    const bool warp_stats = __any_sync(FULL, use_stats);   // every item of a call shares the mode
    const float min_lp = (flags & ITEM_FLOOR) ? a.p.min_log_prob : -INFINITY;
    const int32_t* seq = a.tgt + it.seq_off;
    const int base_max = max(0, N + 1 - S::W);
    const float* my_src = a.logp + it.lp_off;
    const float boostv = a.p.boost_factor;
    int Tmax = T;
#pragma unroll
    for (int d = 8; d < 32; d <<= 1) Tmax = max(Tmax, __shfl_xor_sync(FULL, Tmax, d));
    const int n_chunks = (Tmax + B3_ROWS - 1) / B3_ROWS;

    // ---- class weights of the fused log-sum-exp: exp(x + boost*[c in targets] - boost) = 2^(x*log2e + kk[c]) ----
    __syncwarp();   // previous task's readers are done with kk / bt2
    if (warp_stats) {
        for (int c = l8; c < B3_KK; c += B3_LPU) {
            const bool ok = c < C;
            const bool tg = ok && seg_on && ((a.tmask[(size_t)utt * MAX_WORDS + (c >> 5)] >> (c & 31)) & 1u);
            kk[seg * B3_KK + c] = ok ? (tg ? 0.0f : -boostv * LOG2E) : -INFINITY;
        }
    }

    // lanes with l8 == 0 issue their own utterance's copy and arrive once per chunk on the stage barrier (count = UPW)
    auto issue = [&](int c) {
        if (l8 == 0) {
            const int rows = min(B3_ROWS, T - c * B3_ROWS);
            const uint32_t bar = smem_u32(&bars[c & 1]);
            if (rows > 0) {
                float* dst = stage_buf + (c & 1) * stage_floats + seg * seg_stride;
                const float* s = my_src + (size_t)c * B3_ROWS * C;
                const uint32_t bytes = (uint32_t)rows * C * 4, bulk = bytes & ~15u;
                for (uint32_t w = bulk >> 2; w < (bytes >> 2); ++w) dst[w] = s[w];   // < 4 tail floats of a partial last chunk
                mbar_expect_tx(bar, bulk);
                if (bulk) bulk_g2s_hint(smem_u32(dst), s, bulk, bar, pol);
            } else {
                mbar_arrive(bar);
            }
        }
    };

    // ---- per-lane window state: groups base + l8*G + g, g = 0..G-1; -inf = invalid ----
    int base = 0;
    float P[G], M[G], B3[G];
    int cls[G];
    uint32_t acc[S::ACC];
    uint32_t slide_acc = 0;
#pragma unroll
    for (int i = 0; i < S::ACC; ++i) acc[i] = 0;
    auto group_class = [&](int gi) { return (seg_on && gi >= 1 && gi <= N) ? seq[gi - 1] : blank; };
#pragma unroll
    for (int g = 0; g < G; ++g) {
        P[g] = M[g] = B3[g] = -INFINITY;
        cls[g] = group_class(l8 * G + g);
    }
    if (l8 == 0) B3[0] = 0.0f;        // virtual frame -1: only state 0 is alive, with score 0 (:594-596)
    int next_cls = group_class(S::W);   // class of the group that enters at the next slide (last lane of the segment)
    const bool seg_first = l8 == 0, seg_last = l8 == B3_LPU - 1;

    issue(0);

    bool bad = false;                // needs the exact path
    float fin_val = NEG;
    int fin_cell = 0, fin_base = 0;
    float emax = -INFINITY;          // raw mode: emissions must be <= 0 (log-probabilities)
    float lse_chk = 0.f;             // running sum of the rows' log-sum-exp (finite <=> all rows sane)

    for (int c = 0; c < n_chunks; ++c) {
        const int t0 = c * B3_ROWS;
        __syncwarp();                                  // every lane is done with the other stage (chunk c-1)
        if (c + 1 < n_chunks) issue(c + 1);

        // ---- slide the window to the group of the lower band edge at frame t0-1 (exact :651 arithmetic) ----
        int nslide = 0;
        if (use_band && c > 0) {
            const double center = (double)(t0 - 1) * pace;
            const int s_lo = __float2int_ru((float)(center - bandd));
            nslide = min(max((s_lo + 3) >> 2, 0), base_max) - base;
        }
        slide_acc = (slide_acc << 4) | (uint32_t)nslide;
        while (__any_sync(FULL, nslide > 0)) {
            const float nP = __shfl_down_sync(FULL, P[0], 1), nM = __shfl_down_sync(FULL, M[0], 1);
            const float n3 = __shfl_down_sync(FULL, B3[0], 1);
            const int nc = __shfl_down_sync(FULL, cls[0], 1);
            if (nslide > 0) {
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

        mbar_wait(smem_u32(&bars[c & 1]), (phase >> (c & 1)) & 1u);
        phase ^= 1u << (c & 1);
        __syncwarp();                                  // tail floats written by the issuing lane become visible

        const float* seg_rows = stage_buf + (c & 1) * stage_floats + seg * seg_stride;   // the 8 staged rows of this utterance

        // ---- row statistics: lane (seg, l8) owns row l8 of its utterance: log-sum-exp of the boosted row (:51-54)
        //      and the blank emission; no cross-lane traffic ----
        {
            const float* rowp = seg_rows + l8 * C;
            float lnS = 0.f;
            if (warp_stats) {
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                const float* kp = kk + seg * B3_KK;
                if (CT != 0 && (CT & 1) == 0) {
                    const float2* x2 = reinterpret_cast<const float2*>(rowp);
                    const float4* k4 = reinterpret_cast<const float4*>(kp);
#pragma unroll
                    for (int i = 0; i < CT / 4; ++i) {
                        const float4 k = k4[i];
                        const float2 xa = x2[2 * i], xb = x2[2 * i + 1];
                        s0 += b3_ex2(fmaf(xa.x, LOG2E, k.x));
                        s1 += b3_ex2(fmaf(xa.y, LOG2E, k.y));
                        s2 += b3_ex2(fmaf(xb.x, LOG2E, k.z));
                        s3 += b3_ex2(fmaf(xb.y, LOG2E, k.w));
                    }
                    if (CT % 4) {
                        const float2 xa = x2[CT / 2 - 1];
                        const float2 k = reinterpret_cast<const float2*>(kp)[CT / 2 - 1];
                        s0 += b3_ex2(fmaf(xa.x, LOG2E, k.x));
                        s1 += b3_ex2(fmaf(xa.y, LOG2E, k.y));
                    }
                } else {
                    int i = 0;
                    for (; i + 4 <= C; i += 4) {
                        s0 += b3_ex2(fmaf(rowp[i], LOG2E, kp[i]));
                        s1 += b3_ex2(fmaf(rowp[i + 1], LOG2E, kp[i + 1]));
                        s2 += b3_ex2(fmaf(rowp[i + 2], LOG2E, kp[i + 2]));
                        s3 += b3_ex2(fmaf(rowp[i + 3], LOG2E, kp[i + 3]));
                    }
                    for (; i < C; ++i) s0 += b3_ex2(fmaf(rowp[i], LOG2E, kp[i]));
                }
                lnS = b3_lg2((s0 + s1) + (s2 + s3)) * LN2;   // log sum exp(x + b - boost)
                if (use_stats && t0 + l8 < T) lse_chk += lnS;   // any zero / overflowing / NaN sum leaves a non-finite trace
            } else {
                float m = -INFINITY;
                for (int i = 0; i < C; ++i) m = fmaxf(m, rowp[i]);
                if (t0 + l8 < T) emax = fmaxf(emax, m);
            }
            // blank is never a target: x - lse with lse = lnS + boost; phoneme classes are boosted targets: x - lnS
            const float eb = rowp[blank] - (warp_stats ? lnS + boostv : 0.f);
            stats[seg * B3_STP + l8] = make_float2(lnS, eb);
        }
        __syncwarp();

        const int fin_r = T - 1 - t0;                                      // row of the last frame if it is in this chunk
        const bool fin_here = __any_sync(FULL, fin_r >= 0 && fin_r < B3_ROWS);
        const float* xg[G];                                                // row 0 address of each group's class
#pragma unroll
        for (int g = 0; g < G; ++g) xg[g] = seg_rows + cls[g];
        const float2* stp = stats + seg * B3_STP;
        float xp_n[G];                                                     // raw emissions are fetched one frame ahead
#pragma unroll
        for (int g = 0; g < G; ++g) xp_n[g] = xg[g][0];
        float2 st_n = stp[0];

        auto frame = [&](const int r, auto check_fin) {
            constexpr bool CHECK = decltype(check_fin)::value;
            const float lnS = st_n.x, eb = st_n.y;
            float ep[G];
#pragma unroll
            for (int g = 0; g < G; ++g) ep[g] = fmaxf(xp_n[g] - lnS, min_lp);
            if (CHECK || r + 1 < B3_ROWS) {
                const int rn = (r + 1 < B3_ROWS) ? r + 1 : r;
                st_n = stp[rn];
#pragma unroll
                for (int g = 0; g < G; ++g) xp_n[g] = xg[g][rn * C];
            }

            // ---- DP update, right-most group first so that left neighbours are still frame t-1 ----
            float lm = __shfl_up_sync(FULL, M[G - 1], 1), l3 = __shfl_up_sync(FULL, B3[G - 1], 1);
            if (seg_first) { lm = -INFINITY; l3 = -INFINITY; }   // nothing (or a dropped group, dead for every legal path) to the left
#pragma unroll
            for (int g = G - 1; g >= 0; --g) {
                const float Lm = (g > 0) ? M[(g > 0) ? g - 1 : 0] : lm;
                const float L3 = (g > 0) ? B3[(g > 0) ? g - 1 : 0] : l3;
                const float c0 = P[g] + ep[g], c1 = L3 + ep[g], c2 = Lm + ep[g];   // stay / advance from b3' / skip from b2' (= m')
                const float SP = P[g] + eb, SM = M[g] + eb, S3 = B3[g] + eb;
                uint32_t* A = &acc[4 * g];
                // p  : first max of (c0, c1, c2)                      (:645)
                const float m01 = fmaxf(c0, c1);
                b3_push(A[0], c0, c1);
                b3_push(A[1], m01, c2);
                P[g] = fmaxf(m01, c2);
                // m  : (stay SM, from p SP)  -- b1<-{b1,p} / b2<-{b2,b1,p} merged
                b3_push(A[2], SM, SP);
                M[g] = fmaxf(SM, SP);
                // b3 : (stay S3, advance SM)
                b3_push(A[3], S3, SM);
                B3[g] = fmaxf(S3, SM);
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
                    rec[S::ACC * 32] = slide_acc << (4 * (3 - (c & 3)));
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
        // ---- flush one full 32-frame record (utterances that end inside this chunk flushed at their last frame)
        if ((c & 3) == 3 && T - 1 > t0 + B3_ROWS - 1) {
            uint32_t* rec = slab + (size_t)(c >> 2) * S::REC * 32 + lane;
#pragma unroll
            for (int i = 0; i < S::ACC; ++i) rec[i * 32] = acc[i];
            rec[S::ACC * 32] = slide_acc;
        }
    }
    __syncwarp();
    if (emax > 0.f) bad = true;
    if (use_stats && !(fabsf(lse_chk) < 3.0e38f)) bad = true;
    {   // make `bad` uniform per segment
        const unsigned m = __ballot_sync(FULL, bad);
        bad = ((m >> (seg * B3_LPU)) & 0xffu) != 0;
    }

    // ---- back-trace (:686-703) ----
    // A cell is addressed by its window-relative index ci = 3*(group - base) + k.  Staging lays a record out as
    // bt2[seg][ci] = (first decision word, second decision word or 0), so one 64-bit shared load per frame yields
    // (b0, b1) and every transition is  ci -= b1 ? 2 : b0  (p: from m' = -2 / from b3' = -1, m: from p = -1,
    // b3: from m = -1); a window slide of n groups adds 3n.  cabs = 3*base + ci is slide-invariant.
    const bool walk = seg_on && !bad && T > 0;
    int ci = fin_cell;
    int base3 = 3 * fin_base;
    bool illegal = false;
    const float pace_f = (float)pace, lim_f = (float)(band - B3_MARGIN);
    long long pend_o[4] = {-1, -1, -1, -1};   // outputs of the previous 32-frame block, stored one block late
    int pend_cls[4] = {0, 0, 0, 0}, pend_idx[4] = {0, 0, 0, 0};
    float pend_x[4] = {0.f, 0.f, 0.f, 0.f};
    bool pend_gather[4] = {false, false, false, false};
    const int nblk = (Tmax + 31) >> 5;
    for (int b = nblk - 1; b >= 0; --b) {
        __syncwarp();
        uint32_t sfw;
        {
            const uint32_t* rec = slab + (size_t)b * S::REC * 32 + lane;
            uint32_t w[S::REC];
#pragma unroll
            for (int i = 0; i < S::REC; ++i) w[i] = rec[i * 32];
            uint2* dst = bt2 + seg * S::CELLS + l8 * G * 3;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                dst[3 * g + 0] = make_uint2(w[4 * g + 0], w[4 * g + 1]);   // p : (A0, A1)
                dst[3 * g + 1] = make_uint2(w[4 * g + 2], 0u);             // m : (A2, 0)
                dst[3 * g + 2] = make_uint2(w[4 * g + 3], 0u);             // b3: (A3, 0)
            }
            sfw = w[S::ACC];
        }
        __syncwarp();
        const uint2* cell = bt2 + seg * S::CELLS;
        const int qhi = walk ? min(31, T - 1 - b * 32) : -1;   // last frame of this utterance inside the block
        int keep4[4] = {0, 0, 0, 0};                            // cabs of frames 32b + 8i + l8
#pragma unroll
        for (int q = 31; q >= 0; --q) {
            if (q <= qhi) {
                if ((q & 7) == l8) keep4[q >> 3] = base3 + ci;
                if (q > 0 || b > 0) {                      // frame 0 has no predecessor
                    const uint2 wv = cell[ci];
                    const uint32_t bit = 1u << (31 - q);
                    const int d = (wv.y & bit) ? 2 : ((wv.x & bit) ? 1 : 0);
                    ci -= d;
                    if ((q & 7) == 0) {                    // the window slid before this chunk's first frame was computed
                        const int n3 = 3 * (int)((sfw >> (4 * (3 - (q >> 3)))) & 15u);
                        ci += n3;
                        base3 -= n3;
                    }
                }
            }
        }
        // ---- output of the block, software-pipelined by one block: the loads it needs (target ids, gathered
        //      log-probs for the confidences) are issued now, four independent chains per lane, and consumed
        //      after the next block's walk, so their latency never stalls the warp ----
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (pend_o[i] >= 0) {
                a.frame_ph[pend_o[i]] = pend_cls[i];
                a.frame_idx[pend_o[i]] = pend_idx[i];
                if (pend_gather[i]) a.path_lp[pend_o[i]] = pend_x[i];
            }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            pend_o[i] = -1;
            const int tf = b * 32 + 8 * i + l8;
            if (walk && tf < T) {
                const int gi = keep4[i] / 3, k = keep4[i] - 3 * gi;
                // band legality of the state at frame tf (:650-653), conservative: centre of the (merged) state
                // must be at least B3_MARGIN states inside the band
                if (use_band) {
                    const float s_c = (float)(4 * gi) - (k == 0 ? 3.0f : (k == 1 ? 1.5f : 0.0f));
                    if (fabsf(s_c - (float)tf * pace_f) > lim_f) illegal = true;
                }
                const int rel = tf - trim;
                const long long o = out_off + rel;
                if (rel >= 0 && rel < n_out && o < out_lim) {
                    const bool ph = k == 0;
                    pend_o[i] = o;
                    pend_cls[i] = ph ? seq[gi - 1] : blank;
                    pend_idx[i] = ph ? idx0 + gi - 1 : -1;
                    pend_gather[i] = a.path_lp && (ph || !a.p.ignore_noise);
                }
            }
        }
        // confidences (utils.py:89-103) read lp[f, phoneme of the stamp]: gather it here, once per frame
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (pend_o[i] >= 0 && pend_gather[i]) pend_x[i] = __ldg(my_src + (long long)(b * 32 + 8 * i + l8) * C + pend_cls[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (pend_o[i] >= 0) {
            a.frame_ph[pend_o[i]] = pend_cls[i];
            a.frame_idx[pend_o[i]] = pend_idx[i];
            if (pend_gather[i]) a.path_lp[pend_o[i]] = pend_x[i];
        }
    {   // a path that left (or came too close to) the band is not provably the reference's: exact path
        const unsigned m = __ballot_sync(FULL, illegal);
        if ((m >> (seg * B3_LPU)) & 0xffu) bad = true;
    }
    if (seg_on && l8 == 0) {
        if (bad) {
            const int slot = atomicAdd(a.n_retry, 1);
            a.retry_items[slot] = it;
        } else if ((flags & ITEM_FINAL) && a.dp_final) {
            a.dp_final[utt] = fin_val;
        }
    }
    __syncwarp();
}

template <int G, int CT>
__global__ void __launch_bounds__(B3_WARPS * 32, 1) viterbi_band3_kernel(Band3Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* smem_warp = smem_raw + (size_t)warp * a.smem_per_warp;
    {
        const int C = CT ? CT : a.C;
        float* stage_buf = reinterpret_cast<float*>(smem_warp);
        const int nfl = B3_NST * B3_UPW * B3_ROWS * C + B3_UPW * B3_KK + B3_UPW * B3_STP * 2;
        for (int i = lane; i < nfl; i += 32) stage_buf[i] = 0.0f;   // never-loaded slots must hold finite values
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(stage_buf + nfl);
        if (lane == 0)
            for (int i = 0; i < B3_NST; ++i) mbar_init(smem_u32(&bars[i]), B3_UPW);
        fence_mbar_init();   // also orders the generic-proxy zero fill before the first async copy
    }
    __syncwarp();
    const uint64_t pol = policy_evict_first();
    uint32_t phase = 0;
    const int gwarp = blockIdx.x * B3_WARPS + warp;
    uint32_t* slab = a.bp_scratch + (size_t)gwarp * a.bp_slab_words;
    const int n_items = *a.n_items;
    const int n_tasks = (n_items + B3_UPW - 1) / B3_UPW;
    // static deal: task j -> CTA j % grid, warp (j / grid) % WARPS.  With one CTA per SM this spreads
    // ceil(n_tasks / SMs) tasks evenly over the SMs and over the four schedulers of each SM.
    for (int j = blockIdx.x + gridDim.x * warp; j < n_tasks; j += gridDim.x * B3_WARPS)
        band3_task<G, CT>(a, j * B3_UPW, min(B3_UPW, n_items - j * B3_UPW), smem_warp, slab, phase, lane, pol);
}

}  // namespace bfa
