// viterbi_wide.cuh -- the exact Viterbi (viterbi_generic.cuh) for DP problems of more than 1024 path states: targets of more
// than 255 phonemes that did not split into silence-anchored segments (forced_alignment.py:298-308 keeps retrying for exactly
// those, and :153-190 then aligns them whole).  One CTA owns one problem: warp w holds states [1024 w, 1024 w + 1024) in the
// blocked layout of the generic kernel (J = 32 states per lane); the two states a warp's first lane needs from its left
// neighbour cross through shared memory, one CTA barrier per frame.  Every warp streams and transforms the rows for itself
// (the rows come from L2 after the first warp touched them).  Same arithmetic, candidate order and degenerate behaviour as
// run_item<32>; a rare path, built for results, not for speed.
#pragma once
#include "viterbi_generic.cuh"

namespace bfa {

constexpr int VW_WARPS = 8;                       // warps per CTA = 8192 states at most
constexpr int VW_J = 32;
constexpr int VW_SPAN = 32 * VW_J;                // states per warp
constexpr int VW_MAX_L = VW_WARPS * VW_SPAN;      // = BFA_MAX_L

struct WideShared {
    float edge[2][VW_WARPS][2];                   // [frame parity][warp] = (dp of the warp's last state, of the one before it)
    float fetch;                                  // one state's value, handed from its owner to everybody
    int rmost[VW_WARPS];
    float bv[VW_WARPS];
    int bs[VW_WARPS];
    int item;
};

__device__ __forceinline__ void vw_bar(int nthreads) {
    __syncwarp();
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// words per frame of a wide item's back-pointer slab: nw warps x 32 lanes x one 64-bit word (2 bits per state)
__host__ __device__ inline long long vw_slab_words(int max_T, int max_L) {
    const int nw = (max_L + VW_SPAN - 1) / VW_SPAN;
    return (long long)(max_T + 2) * nw * 32 * 2;
}

__global__ void __launch_bounds__(VW_WARPS * 32, 1) viterbi_wide_kernel(VitArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ WideShared ws;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_wait();
    const int n_items = *a.n_items;
    WarpSmem& sm = reinterpret_cast<WarpSmem*>(smem_raw)[warp];
    if (lane == 0) {
        for (int i = 0; i < NSTAGE; ++i) mbar_init(smem_u32(&sm.bar[i]), 1);
        fence_mbar_init();
    }
    __syncwarp();
    const uint64_t pol = policy_evict_first();
    Stream st;
    st.phase = 0;
    unsigned long long* bp = reinterpret_cast<unsigned long long*>(a.bp_scratch + (size_t)blockIdx.x * a.bp_slab_words);
    const float NEG = a.p.neg_inf;
    const int blank = a.p.blank_id, C = a.C;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {                   // next item with more than 1024 states
            int i;
            do { i = atomicAdd(a.work_counter, 1); } while (i < n_items && a.items[i].L <= VW_SPAN);
            ws.item = i;
        }
        __syncthreads();
        const int ii = ws.item;
        if (ii >= n_items) break;
        const Item& it = a.items[ii];
        const int T = it.T, L = it.L;
        const int nw = (L + VW_SPAN - 1) / VW_SPAN;
        const int nthr = nw * 32;
        if (warp >= nw) continue;                 // idle for this item; meets the others at the top of the loop
        const int base = warp * VW_SPAN;
        const bool use_stats = (it.flags & ITEM_STATS) != 0, do_floor = (it.flags & ITEM_FLOOR) != 0, has_anchor = (it.flags & ITEM_ANCHOR) != 0;

        int pid[VW_J];
        uint32_t skipmask = 0;
#pragma unroll
        for (int j = 0; j < VW_J; ++j) {
            const int s = base + lane * VW_J + j;
            int cls = blank;
            if (s < L) {
                cls = state_class(a, it, s);
                if (s >= 2 && cls != state_class(a, it, s - 2)) skipmask |= 1u << j;      // can_skip (:603-605)
            }
            pid[j] = min(max(cls, 0), C - 1);
        }
        uint32_t tbits = 0;
        if (a.tmask && (use_stats || do_floor)) {
#pragma unroll
            for (int i = 0; i < MAX_WORDS; ++i) tbits |= ((a.tmask[(size_t)it.utt * MAX_WORDS + i] >> lane) & 1u) << i;
        }
        {
            const float* g = a.logp + it.lp_off;
            const uintptr_t addr = (uintptr_t)g;
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
        float dp[VW_J];
#pragma unroll
        for (int j = 0; j < VW_J; ++j) {
            const int s = base + lane * VW_J + j;
            dp[j] = NEG;
            if (s == 0) dp[j] = sm.rowbuf[0][blank];
            if (s == 1 && L > 1) dp[j] = sm.rowbuf[0][pid[j]];
        }
        if (lane == 31) { ws.edge[0][warp][0] = dp[VW_J - 1]; ws.edge[0][warp][1] = dp[VW_J - 2]; }
        if (T > 1) produce(1);
        __syncwarp();
        vw_bar(nthr);

        const bool use_band = it.band > 0 && T > 1 && L > 1;                      // :586
        const double pace = use_band ? (double)(L - 1) / (double)(T - 1) : 0.0;  // :587

        // ---- forward pass (:608-653) ----
        for (int t = 1; t < T; ++t) {
            if (t + 1 < T) produce(t + 1);
            const float* rb = sm.rowbuf[t & 1];
            float lo = -INFINITY, hi = INFINITY;
            if (use_band) {
                const double center = (double)t * pace;                            // :651
                lo = (float)(center - (double)it.band);
                hi = (float)(center + (double)it.band);
            }
            float l1 = __shfl_up_sync(FULL, dp[VW_J - 1], 1);
            float l2 = __shfl_up_sync(FULL, dp[VW_J - 2], 1);
            if (lane == 0 && warp > 0) { l1 = ws.edge[(t - 1) & 1][warp - 1][0]; l2 = ws.edge[(t - 1) & 1][warp - 1][1]; }
            unsigned long long word = 0;
#pragma unroll
            for (int j = VW_J - 1; j >= 0; --j) {
                const int s = base + lane * VW_J + j;
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
                word |= (unsigned long long)k << (2 * j);
            }
            bp[((size_t)t * nw + warp) * 32 + lane] = word;
            if (lane == 31) { ws.edge[t & 1][warp][0] = dp[VW_J - 1]; ws.edge[t & 1][warp][1] = dp[VW_J - 2]; }
            __syncwarp();
            vw_bar(nthr);
        }

        // ---- final state (:656-682) ----
        auto fetch = [&](int s) -> float {        // dp of state s at the last frame, for every thread of the item
            if (warp == s / VW_SPAN) {
                const float v = get_state<VW_J>(dp, s - base, lane);
                if (lane == 0) ws.fetch = v;
            }
            vw_bar(nthr);
            const float v = ws.fetch;
            vw_bar(nthr);
            return v;
        };
        {
            int rmost = -1, bs = -1;
            float bv = -INFINITY;
#pragma unroll
            for (int j = 0; j < VW_J; ++j) {
                const int s = base + lane * VW_J + j;
                if (s < L && dp[j] > NEG) {
                    rmost = s;
                    if (bs < 0 || dp[j] > bv) { bv = dp[j]; bs = s; }
                }
            }
            rmost = warp_max_i(rmost);
            if (lane == 0) ws.rmost[warp] = rmost;
        }
        vw_bar(nthr);
        int f;
        float fv;
        if (!a.p.truly_forced) {
            // best reachable state (first max); when none is reachable, arg-max over all states (:659-665)
            bool any = false;
            for (int w = 0; w < nw; ++w) any = any || ws.rmost[w] >= 0;
            int bs = -1;
            float bv = -INFINITY;
#pragma unroll
            for (int j = 0; j < VW_J; ++j) {
                const int s = base + lane * VW_J + j;
                if (s < L && (!any || dp[j] > NEG) && (bs < 0 || dp[j] > bv)) { bv = dp[j]; bs = s; }
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const float ov = __shfl_xor_sync(FULL, bv, d);
                const int os = __shfl_xor_sync(FULL, bs, d);
                if (os >= 0 && (bs < 0 || ov > bv || (ov == bv && os < bs))) { bv = ov; bs = os; }
            }
            if (lane == 0) { ws.bv[warp] = bv; ws.bs[warp] = bs; }
            vw_bar(nthr);
            f = -1; fv = -INFINITY;
            for (int w = 0; w < nw; ++w)
                if (ws.bs[w] >= 0 && (f < 0 || ws.bv[w] > fv)) { fv = ws.bv[w]; f = ws.bs[w]; }   // ties -> lower state
            vw_bar(nthr);
        } else {
            f = L - 1;
            fv = fetch(f);
            if (fv <= NEG && L >= 2) { f = L - 2; fv = fetch(f); }
            if (fv <= NEG) {
                int r = -1;
                for (int w = 0; w < nw; ++w) r = max(r, ws.rmost[w]);
                f = (r >= 0) ? r : L - 1;
                fv = fetch(f);
            }
        }
        if (threadIdx.x == 0) {
            if (a.final_state) a.final_state[ii] = f;
            if ((it.flags & ITEM_FINAL) && a.dp_final) a.dp_final[it.utt] = fv;
            if (a.status && fv <= NEG) atomicOr(&a.status[it.utt], BFA_ST_DEGENERATE);
        }
        __threadfence_block();
        vw_bar(nthr);                              // every warp's back-pointer words are visible to warp 0
        if (warp != 0) continue;

        // ---- back-trace (:686-703) by warp 0: lane q pre-loads, for frame tb - q, the words of the lane that owns the current
        //      state and of the two lanes below it (the path moves down at most two states per frame: 64 states in 32 frames)
        int ps = f;
        int keep = -1;
        auto flush = [&](int t_lo) {
            const int t = t_lo + lane;
            if (t < T && keep >= 0) {
                const int rel = t - it.trim;
                const long long o = it.out_off + rel;
                if (rel >= 0 && rel < it.n_out && o < it.out_lim) {
                    const int cls = state_class(a, it, keep);
                    a.frame_ph[o] = cls;
                    a.frame_idx[o] = state_tidx(a, it, keep);
                    if (a.path_lp && cls >= 0 && cls < C) a.path_lp[o] = a.logp[it.lp_off + (long long)t * C + cls];
                }
            }
            keep = -1;
        };
        for (int tb = T - 1; tb >= 0; tb -= 32) {
            const int g0 = ps / VW_J;              // global lane (warp * 32 + lane) that owns the state at frame tb
            unsigned long long w3[3] = {0ull, 0ull, 0ull};
            {
                const int t = tb - lane;
                if (t >= 1) {
#pragma unroll
                    for (int d = 0; d < 3; ++d)
                        if (g0 - d >= 0) w3[d] = bp[(size_t)t * nw * 32 + (g0 - d)];
                }
            }
            for (int q = 0; q < 32; ++q) {
                const int t = tb - q;
                if (t < 0) break;
                if ((t & 31) == lane) keep = ps;
                if ((t & 31) == 0) flush(t);
                if (t >= 1) {
                    const int g = ps / VW_J, jj = ps - g * VW_J, d = g0 - g;
                    unsigned long long ww;
                    if (d >= 0 && d < 3) {
                        const unsigned long long mine = d == 0 ? w3[0] : (d == 1 ? w3[1] : w3[2]);
                        ww = __shfl_sync(FULL, mine, q);
                    } else {
                        ww = bp[(size_t)t * nw * 32 + g];          // after a negative-index wrap (:692): anywhere
                    }
                    const int k = (int)((ww >> (2 * jj)) & 3);
                    ps -= k;                       // back-pointer = s - k (:647)
                    if (ps < 0) ps += L;           // python negative-index wrap (:692)
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace bfa
