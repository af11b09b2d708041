// viterbi_generic.cuh -- exact warp-per-item Viterbi fill + back-trace for ANY path / stride / band.
//
// Restates ViterbiDecoder._viterbi_decode (forced_alignment.py:563-703) including the
// degenerate behaviour (masked candidates are the constant -1000, negative back-pointers wrap).
// One warp owns one DP problem.  State s lives in lane s / J, register s % J (blocked layout:
// only two shuffles per frame bring the left neighbour's last two states).  Log-posterior rows
// are streamed HBM -> shared memory as a flat byte range with 1-D bulk async copies
// (cp.async.bulk + mbarrier, the TMA engine's non-tensor path) into a per-warp ring; each row is
// transformed once (target boost / log-softmax / floor / sub-silence anchoring, the reference's
// :121-129 and :543-561) into a double-buffered row buffer, from which the lanes gather emissions.
// Back-pointers are packed 2 bits/state, one word per lane per frame, in a per-warp scratch slab.
#pragma once
#include "bfa_common.cuh"

namespace bfa {

constexpr int VG_WARPS = 8;            // warps per CTA
constexpr int RING = 2048;             // floats per warp ring (8 KB)
constexpr int CHUNK = 512;             // floats per bulk copy (2 KB)
constexpr int NSTAGE = RING / CHUNK;   // 4
constexpr int ROWBUF = BFA_MAX_C;      // floats per row buffer

struct VitArgs {
    BfaParams p;
    int C;
    const float* logp;
    const uint32_t* tmask;      // [B][MAX_WORDS] target-class bitmask; may be null (explicit items)
    const int32_t* tgt;         // flat targets (structured items)
    const int32_t* path;        // explicit path arrays (explicit items)
    const int32_t* true_idx;
    const uint32_t* anchors;    // 4-bit anchor counts, 8 frames per word
    const Item* items;
    const int* n_items;         // device scalar
    int* work_counter;          // device scalar, zeroed before launch
    const int* first;           // device scalar: index of the first item of this launch (null: 0)
    int warp_base;              // first slab of this launch (launches that run side by side own disjoint slab ranges)
    int32_t* frame_ph;
    int32_t* frame_idx;
    float* path_lp;             // [total_frames] raw log-prob of the assigned class per frame, or null
    float* dp_final;            // per utterance (FINAL items) -- or per item for explicit entry
    int32_t* status;            // per utterance: DEGENERATE flag or-ed in
    int32_t* final_state;       // per item, may be null
    uint32_t* bp_scratch;       // per resident warp slab
    long long bp_slab_words;    // words per slab
};

struct WarpSmem {
    float ring[RING];
    float rowbuf[2][ROWBUF];
    unsigned long long bar[NSTAGE];
};

// class id of state s (forced_alignment.py:181-187 path construction)
__device__ __forceinline__ int state_class(const VitArgs& a, const Item& it, int s) {
    if (it.stride == 0) return a.path[it.seq_off + s];
    if (s >= 1 && (s - 1) % it.stride == 0) return a.tgt[it.seq_off + (s - 1) / it.stride];
    return a.p.blank_id;
}
__device__ __forceinline__ int state_tidx(const VitArgs& a, const Item& it, int s) {
    if (it.stride == 0) return a.true_idx ? a.true_idx[it.seq_off + s] : -1;
    if (s >= 1 && (s - 1) % it.stride == 0) return it.idx0 + (s - 1) / it.stride;
    return -1;
}

struct Stream {          // per-warp streaming state for the current item
    const float* g16;    // 16-B aligned-down global address of the first element
    int shift;           // floats between g16 and the first element (0..3)
    int total;           // shift + T*C  (positions [shift, total) hold data)
    int n_chunks;
    int issued;          // chunks issued so far
    int ready;           // chunks waited for so far
    uint32_t phase;      // per-stage parity bits (persist across items)
};

// Issue chunk c (lane 0 only): bulk-copy the 16-B aligned interior, patch head/tail floats by hand.
__device__ __forceinline__ void issue_chunk(WarpSmem& sm, const Stream& st, int c, uint64_t pol) {
    const int stage = c % NSTAGE;
    int p0 = c * CHUNK, p1 = min(p0 + CHUNK, (st.total + 3) & ~3);
    // aligned interior of the valid range: [ceil4(shift), floor4(total))
    int lo = max(p0, (st.shift + 3) & ~3), hi = min(p1, st.total & ~3);
    const uint32_t bar = smem_u32(&sm.bar[stage]);
    uint32_t bytes = hi > lo ? (uint32_t)(hi - lo) * 4u : 0u;
    mbar_expect_tx(bar, bytes);
    if (bytes) bulk_g2s_hint(smem_u32(&sm.ring[lo & (RING - 1)]), st.g16 + lo, bytes, bar, pol);
    // head (positions shift..3 when shift>0) and tail (floor4(total)..total) are < 4 floats each
    if (c == 0 && st.shift)
        for (int q = st.shift; q < min(4, st.total); ++q) sm.ring[q] = st.g16[q];
    int tail0 = max(st.total & ~3, (st.shift + 3) & ~3);
    if (tail0 >= p0 && tail0 < p1)
        for (int q = tail0; q < st.total; ++q) sm.ring[q & (RING - 1)] = st.g16[q];
}

// Make positions [.., pe] readable and recycle ring slots that precede position pb.
__device__ __forceinline__ void stream_advance(WarpSmem& sm, Stream& st, int pb, int pe, int lane, uint64_t pol) {
    if (lane == 0) {
        int limit = min(st.n_chunks, pb / CHUNK + NSTAGE);
        while (st.issued < limit) issue_chunk(sm, st, st.issued++, pol);
    }
    st.issued = __shfl_sync(FULL, st.issued, 0);
    int ce = pe / CHUNK;
    while (st.ready <= ce) {
        int stage = st.ready % NSTAGE;
        mbar_wait(smem_u32(&sm.bar[stage]), (st.phase >> stage) & 1u);
        st.phase ^= 1u << stage;
        ++st.ready;
    }
}

// Transform row t of the item into rowbuf[t & 1]: target boost + log_softmax + floor (:121-129) and the sub-silence
// anchoring of silence-anchored segments (:543-561), from the warp's ring.
struct RowProducer {
    const VitArgs& a;
    const Item& it;
    WarpSmem& sm;
    Stream& st;
    const int lane;
    const uint64_t pol;
    const uint32_t tbits;
    const bool use_stats, do_floor, has_anchor;
    uint32_t anc_blk;                           // lane q holds anchor word q of the current 256-frame block
    __device__ __forceinline__ void operator()(int t) {
        const int C = a.C, T = it.T, blank = a.p.blank_id;

        int pos = st.shift + t * C;
        stream_advance(sm, st, pos, pos + C - 1, lane, pol);
        __syncwarp();
        float m = 0.f, ls = 0.f;
        if (use_stats) {   // boost + log_softmax statistics of this row (:51-54), same bits as rowstat_kernel
            float2 st2 = row_stats_warp([&](int c) { return sm.ring[(pos + c) & (RING - 1)]; }, C, lane, tbits, a.p.boost_factor);
            m = st2.x; ls = st2.y;
        }
        int cnt = 0;
        if (has_anchor) {
            if ((t & 255) == 0) {
                int w = (t >> 3) + lane;
                anc_blk = (w < (T + 7) / 8) ? a.anchors[it.anchor_off + w] : 0u;
            }
            uint32_t word = __shfl_sync(FULL, anc_blk, (t >> 3) & 31);
            cnt = (word >> ((t & 7) * 4)) & 15;
        }
        float* rb = sm.rowbuf[t & 1];
        if (cnt == 0) {
#pragma unroll
            for (int i = 0; i < MAX_WORDS; ++i) {
                int c = lane + 32 * i;
                if (c < C) {
                    float x = sm.ring[(pos + c) & (RING - 1)];
                    rb[c] = mod_value(x, (tbits >> i) & 1u, use_stats, do_floor, a.p.boost_factor, m, ls, a.p.min_log_prob);
                }
            }
        } else {
            // sub-silence anchoring (:543-561): cnt times { row[blank] += boost; row = log_softmax(row) }
            float v[MAX_WORDS];
#pragma unroll
            for (int i = 0; i < MAX_WORDS; ++i) {
                int c = lane + 32 * i;
                v[i] = -INFINITY;
                if (c < C) {
                    float x = sm.ring[(pos + c) & (RING - 1)];
                    v[i] = mod_value(x, (tbits >> i) & 1u, use_stats, do_floor, a.p.boost_factor, m, ls, a.p.min_log_prob);
                }
            }
            for (int r = 0; r < cnt; ++r) {
                float mx = -INFINITY;
#pragma unroll
                for (int i = 0; i < MAX_WORDS; ++i) {
                    int c = lane + 32 * i;
                    if (c == blank) v[i] += a.p.sub_boost;
                    if (c < C) mx = fmaxf(mx, v[i]);
                }
                mx = warp_max(mx);
                float sum = 0.f;
#pragma unroll
                for (int i = 0; i < MAX_WORDS; ++i)
                    if (lane + 32 * i < C) sum += expf(v[i] - mx);
                sum = warp_sum(sum);
                float lsum = logf(sum);
#pragma unroll
                for (int i = 0; i < MAX_WORDS; ++i) v[i] = (v[i] - mx) - lsum;
            }
#pragma unroll
            for (int i = 0; i < MAX_WORDS; ++i)
                if (lane + 32 * i < C) rb[lane + 32 * i] = v[i];
        }
    }
};

template <int J>
__device__ __forceinline__ float get_state(const float (&dp)[J], int s, int lane) {
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < J; ++j)
        if (j == (s % J)) v = dp[j];
    return __shfl_sync(FULL, v, s / J);
}

template <int J>
__device__ void run_item(const VitArgs& a, const Item& it, WarpSmem& sm, Stream& st, uint32_t* bp, int lane, uint64_t pol) {
    using bpw_t = typename std::conditional<(J > 16), unsigned long long, uint32_t>::type;
    constexpr int BPW = (J > 16) ? 2 : 1;
    const int T = it.T, L = it.L, C = a.C;
    const float NEG = a.p.neg_inf;
    const bool use_stats = (it.flags & ITEM_STATS) != 0;
    const bool do_floor = (it.flags & ITEM_FLOOR) != 0;
    const bool has_anchor = (it.flags & ITEM_ANCHOR) != 0;
    const int blank = a.p.blank_id;

    // ---- per-lane path description ----
    int pid[J];
    uint32_t skipmask = 0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        int s = lane * J + j;
        int cls = blank;
        if (s < L) {
            cls = state_class(a, it, s);
            if (s >= 2 && cls != state_class(a, it, s - 2)) skipmask |= 1u << j;  // can_skip (:603-605)
        }
        pid[j] = min(max(cls, 0), C - 1);
    }
    // target-class bits for the classes this lane transforms: class lane + 32*i
    uint32_t tbits = 0;
    if (a.tmask && (use_stats || do_floor)) {
#pragma unroll
        for (int i = 0; i < MAX_WORDS; ++i) tbits |= ((a.tmask[(size_t)it.utt * MAX_WORDS + i] >> lane) & 1u) << i;
    }

    // ---- stream set-up ----
    {
        const float* g = a.logp + it.lp_off;
        uintptr_t addr = (uintptr_t)g;
        st.g16 = (const float*)(addr & ~(uintptr_t)15);
        st.shift = (int)((addr & 15) >> 2);
        st.total = st.shift + T * C;
        st.n_chunks = (st.total + CHUNK - 1) / CHUNK;
        st.issued = 0;
        st.ready = 0;
    }

    RowProducer produce{a, it, sm, st, lane, pol, tbits, use_stats, do_floor, has_anchor, 0u};

    // ---- t = 0 (:594-596) ----
    produce(0);
    __syncwarp();
    float dp[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        int s = lane * J + j;
        dp[j] = NEG;
        if (s == 0) dp[j] = sm.rowbuf[0][blank];
        if (s == 1 && L > 1) dp[j] = sm.rowbuf[0][pid[j]];
    }
    if (T > 1) produce(1);
    __syncwarp();

    const bool use_band = it.band > 0 && T > 1 && L > 1;                      // :586
    const double pace = use_band ? (double)(L - 1) / (double)(T - 1) : 0.0;  // :587

    // ---- forward pass (:608-653) ----
    for (int t = 1; t < T; ++t) {
        if (t + 1 < T) produce(t + 1);
        const float* rb = sm.rowbuf[t & 1];
        float lo = -INFINITY, hi = INFINITY;
        if (use_band) {
            double center = (double)t * pace;                                  // :651
            lo = (float)(center - (double)it.band);
            hi = (float)(center + (double)it.band);
        }
        float l1 = __shfl_up_sync(FULL, dp[J - 1], 1);
        float l2 = (J >= 2) ? __shfl_up_sync(FULL, dp[(J >= 2) ? J - 2 : 0], 1) : __shfl_up_sync(FULL, dp[0], 2);
        bpw_t word = 0;
#pragma unroll
        for (int j = J - 1; j >= 0; --j) {
            const int s = lane * J + j;
            const float e = rb[pid[j]];
            const float p1 = (j >= 1) ? dp[(j >= 1) ? j - 1 : 0] : l1;
            const float p2 = (j >= 2) ? dp[(j >= 2) ? j - 2 : 0] : ((j == 1) ? l1 : l2);
            const float c0 = dp[j] + e;                                        // stay    (:613)
            const float c1 = (s >= 1) ? p1 + e : NEG;                          // advance (:616-617)
            const float c2 = ((skipmask >> j) & 1u) ? p2 + e : NEG;            // skip    (:620-625)
            int k = 0;
            float best = c0;                                                   // first max wins (:645)
            if (c1 > best) { best = c1; k = 1; }
            if (c2 > best) { best = c2; k = 2; }
            if ((float)s < lo || (float)s > hi) best = NEG;                    // band (:650-653)
            dp[j] = best;
            word |= (bpw_t)k << (2 * j);
        }
        if (BPW == 1) bp[(size_t)t * 32 + lane] = (uint32_t)word;
        else reinterpret_cast<unsigned long long*>(bp)[(size_t)t * 32 + lane] = (unsigned long long)word;
        __syncwarp();
    }

    // ---- final state (:656-682) ----
    int f;
    float fv;
    {
        int rmost = -1, bs = -1;
        float bv = -INFINITY;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            int s = lane * J + j;
            if (s < L && dp[j] > NEG) {
                rmost = s;
                if (bs < 0 || dp[j] > bv) { bv = dp[j]; bs = s; }
            }
        }
        if (!a.p.truly_forced) {
            bool any = __any_sync(FULL, bs >= 0);
            if (!any) {  // argmax over all states (first max)
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    int s = lane * J + j;
                    if (s < L && (bs < 0 || dp[j] > bv)) { bv = dp[j]; bs = s; }
                }
            }
            // warp arg-max, ties -> lower state index
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                float ov = __shfl_xor_sync(FULL, bv, d);
                int os = __shfl_xor_sync(FULL, bs, d);
                if (os >= 0 && (bs < 0 || ov > bv || (ov == bv && os < bs))) { bv = ov; bs = os; }
            }
            f = bs;
            fv = bv;
        } else {
            f = L - 1;
            fv = get_state<J>(dp, f, lane);
            if (fv <= NEG && L >= 2) { f = L - 2; fv = get_state<J>(dp, f, lane); }
            if (fv <= NEG) {
                int r = warp_max_i(rmost);
                f = (r >= 0) ? r : L - 1;
                fv = get_state<J>(dp, f, lane);
            }
        }
    }
    if (lane == 0) {
        if (a.final_state) a.final_state[(int)(&it - a.items)] = f;
        if (it.flags & ITEM_FINAL) {
            if (a.dp_final) a.dp_final[it.utt] = fv;
        }
        if (a.status && fv <= NEG) atomicOr(&a.status[it.utt], BFA_ST_DEGENERATE);
    }

    // ---- back-trace (:686-703), 16 frames of back-pointer words pre-loaded per step ----
    int ps = f;
    int keep = -1;  // lane (t & 31) keeps the state of frame t until a 32-frame group is complete
    auto flush = [&](int t_lo) {  // write frames [t_lo, t_lo+32) ∩ [0,T): lane q holds frame t_lo + q
        int t = t_lo + lane;
        if (t < T && keep >= 0) {
            int rel = t - it.trim;
            long long o = it.out_off + rel;
            if (rel >= 0 && rel < it.n_out && o < it.out_lim) {
                const int cls = state_class(a, it, keep);
                a.frame_ph[o] = cls;
                a.frame_idx[o] = state_tidx(a, it, keep);
                if (a.path_lp && cls >= 0 && cls < C) a.path_lp[o] = a.logp[it.lp_off + (long long)t * C + cls];
            }
        }
        keep = -1;
    };
    for (int tb = T - 1; tb >= 0; tb -= 16) {
        bpw_t w[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            int t = tb - q;
            w[q] = 0;
            if (t >= 1) {
                if (BPW == 1) w[q] = bp[(size_t)t * 32 + lane];
                else w[q] = reinterpret_cast<const unsigned long long*>(bp)[(size_t)t * 32 + lane];
            }
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            int t = tb - q;
            if (t >= 0) {
                if ((t & 31) == lane) keep = ps;
                if ((t & 31) == 0) flush(t);
                if (t >= 1) {
                    int owner = ps / J, jj = ps - owner * J;
                    bpw_t ww;
                    if (BPW == 1) ww = __shfl_sync(FULL, (uint32_t)w[q], owner);
                    else ww = (bpw_t)__shfl_sync(FULL, (unsigned long long)w[q], owner);
                    int k = (int)((ww >> (2 * jj)) & 3);
                    ps -= k;                   // back-pointer = s - k (:647)
                    if (ps < 0) ps += L;       // python negative-index wrap (:692)
                }
            }
        }
    }
    __syncwarp();
}

// Two register classes so that the common short-path case keeps 2 CTAs (16 warps) per SM:
// KCLASS 0 handles items with L <= 256 (J <= 8), KCLASS 1 the rest (J = 16 / 32).  Both kernels
// walk the same item list with their own work counter and skip items of the other class.
template <int KCLASS>
__global__ void __launch_bounds__(VG_WARPS * 32, KCLASS == 0 ? 2 : 1) viterbi_generic_kernel(VitArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_release();
    pdl_wait();                   // the item list of the planner / the banded kernel's retries
    const int first = a.first ? *a.first : 0;
    const int n_items = *a.n_items - first;
    if (n_items <= 0) return;     // nothing (left) for the exact kernel
    WarpSmem& sm = reinterpret_cast<WarpSmem*>(smem_raw)[warp];
    if (lane == 0) {
        for (int i = 0; i < NSTAGE; ++i) mbar_init(smem_u32(&sm.bar[i]), 1);
        fence_mbar_init();
    }
    __syncwarp();
    const uint64_t pol = policy_evict_first();
    Stream st;
    st.phase = 0;
    const int gwarp = a.warp_base + blockIdx.x * VG_WARPS + warp;
    uint32_t* bp = a.bp_scratch + (size_t)gwarp * a.bp_slab_words;
    for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(a.work_counter + KCLASS, 1);
        i = __shfl_sync(FULL, i, 0);
        if (i >= n_items) break;
        const Item& it = a.items[first + i];
        const int J = (it.L + 31) / 32;
        if (KCLASS == 0) {
            if (J > 8) continue;
            if (J <= 1) run_item<1>(a, it, sm, st, bp, lane, pol);
            else if (J <= 2) run_item<2>(a, it, sm, st, bp, lane, pol);
            else if (J <= 4) run_item<4>(a, it, sm, st, bp, lane, pol);
            else run_item<8>(a, it, sm, st, bp, lane, pol);
        } else {
            if (J <= 8 || J > 32) continue;        // J > 32: viterbi_wide_kernel
            if (J <= 16) run_item<16>(a, it, sm, st, bp, lane, pol);
            else run_item<32>(a, it, sm, st, bp, lane, pol);
        }
    }
}

}  // namespace bfa
