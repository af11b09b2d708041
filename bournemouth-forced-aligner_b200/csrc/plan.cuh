// plan.cuh -- per-utterance planning on the device: target masks, row statistics, silence
// scan + silence-anchored segmentation, and emission of DP work items.
//
// Restates the control logic of ViterbiDecoder.decode_with_forced_alignment
// (forced_alignment.py:87-199), _segmented_viterbi_decode (:268-469), _find_target_sil_groups
// (:203-224), _match_silences (:226-266), _detect_silence_segments (:471-541) and
// AlignmentUtils.decode_alignments_simple (:932-986).  One warp per utterance; list logic is
// executed redundantly by all lanes (uniform control flow), lane 0 writes, the frame-parallel
// parts (exp / prefix sums / fills) use all 32 lanes.  No host round trip: the work-item list and
// its length stay on the device.
#pragma once
#include "bfa_common.cuh"
#include "silscan.cuh"
#include "viterbi_band3.cuh"

namespace bfa {

struct PlanArgs {
    BfaParams p;
    int B, C, max_T, max_N;
    const float* logp;
    const long long* row_off;
    const int32_t* T;
    const int32_t* tgt;
    const long long* tgt_off;
    const long long* frame_off;
    double* D;                 // chunk-local running sums of the silence probabilities (silscan.cuh), or null (no segmentation)
    int sil_ready;             // silprob_kernel ran: D holds every utterance whose target has silence_id.  0: the caller hinted that
                               // no target does; an utterance that has one anyway gets its sums from the planner itself (slow, same values)
    uint32_t* tmask;           // [B][MAX_WORDS]
    int32_t* frame_ph;
    int32_t* frame_idx;
    float* dp_final;
    int32_t* status;
    // scratch
    int item_cap;              // local item slots per utterance
    int gmax, amax;            // list capacities per utterance
    int anchor_words;          // anchor words per utterance
    Item* items_local;         // [B][item_cap]
    Item* items;               // compacted list for the exact generic kernel
    int* n_items;              // device counter
    Item* fast_items[3];       // compacted lists for the banded kernel (window 24 / 40 / 64 groups)
    int* n_fast[3];
    int* frames_fast[3];       // frames of each list's items (the banded kernel sizes the classes' CTA shares with them)
    int fast_enable;
    int32_t* lists;            // [B][list_ints]
    int list_ints;
    double* cbase;             // [B][cb_pitch] chunk bases of the running sums
    int cb_pitch;
    uint32_t* anchors;         // [B][anchor_words]
    float* path_lp;            // [total_frames] raw log-prob of the class each frame was assigned to (for confidences), or null
    const int* deferred;       // when non-null: plan only the utterances deferred[0 .. *n_deferred) (what the direct kernel left)
    const int* n_deferred;
};

// Which kernel runs an item: 0 / 1 / 2 = banded kernel with a 24 / 40 / 64-group window, -1 = exact generic kernel.
__device__ __forceinline__ int fast_class(const Item& it, int C, const float* logp, bool tgt_ok, int fast_enable) {
    if (!fast_enable || !tgt_ok) return -1;
    if (it.stride != 4 || (it.flags & ITEM_ANCHOR) || C > B3_KK || it.T < 2 || it.L > it.T || it.n > B3_NMAX) return -1;
    // Bulk copies need 16-byte aligned sources: a misaligned item (e.g. a segment that starts on an odd row of a C=66 matrix)
    // is copied from the boundary before it, `lead` floats early.  Those floats must exist inside the caller's buffer, and the
    // C=66 instantiation reads the staged rows as 8-byte pairs.
    if (it.lead != 0 && (it.lp_off < it.lead || (C == 66 && (it.lead & 1)))) return -1;
    const int need = band3_window_need(it.n, it.T, it.L, it.band);
    if (need <= 24) return 0;
    if (need <= 40) return 1;
    if (need <= 64) return 2;
    return -1;
}

// ------------------------------------------------------------------------------------------
struct UttCtx {
    const PlanArgs* a;
    int u, lane, T, N;
    const float* lp;         // utterance rows
    const int32_t* seq;
    const double* D;         // running sums of the silence probabilities, local to chunks of SS_CHUNK rows (silprob_kernel), or null
    double* cbase;           // [T / SS_CHUNK + 1] sum of everything before each chunk (chunk_bases)
    const uint32_t* mw;      // the utterance's target-class mask words (shared memory)
    double* ds;              // this warp's staging buffer for a slab of running sums: [PLAN_SLAB + PLAN_KMAX + 4] doubles + 4 chunk bases
    uint32_t bar;            // its mbarrier
    uint32_t* phase;         // parity of the next completion of that barrier
};
constexpr int PLAN_SLAB = 512;     // windows of the silence scan per staged slab
constexpr int PLAN_KMAX = 64;      // longest moving average the staged scan handles (silence_anchors; 10 by default)
constexpr int PLAN_DS = PLAN_SLAB + PLAN_KMAX + 4;

// Sum of the silence probabilities of rows 0 .. t of the utterance (0 for t < 0): chunk-local running sum + chunk base.
__device__ __forceinline__ double sil_prefix(const UttCtx& c, int t) {
    return t < 0 ? 0.0 : c.D[t] + c.cbase[t >> SS_CHUNK_SHIFT];
}
// cbase[j] = sum of the probabilities of all rows before chunk j (an fp64 scan over the chunks' last values).
__device__ void chunk_bases(const UttCtx& c) {
    const int nch = (c.T + SS_CHUNK - 1) >> SS_CHUNK_SHIFT;
    double carry = 0.0;
    for (int b = 0; b < nch; b += 32) {
        const int j = b + c.lane;
        double v = 0.0;
        if (j < nch) v = c.D[min(j * SS_CHUNK + SS_CHUNK - 1, c.T - 1)];
        const double own = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double o = __shfl_up_sync(FULL, v, d);
            if (c.lane >= d) v += o;
        }
        v += carry;
        if (j < nch) c.cbase[j] = v - own;
        carry = __shfl_sync(FULL, v, 31);
    }
    __syncwarp();
}

// Run logic of _detect_silence_segments (:519-539) over the silence bits of windows [base, base + 32): carries (in_sil, start, n).
struct SilRuns {
    int n = 0, in_sil = 0, start = 0;
    __device__ __forceinline__ void feed(uint32_t bits, int base, int nwin, int k, int Tn, int lane, int32_t* out, int max_out) {
        const uint32_t vmask = (nwin - base >= 32) ? FULL : ((1u << (nwin - base)) - 1u);
        uint32_t trans = (bits ^ ((bits << 1) | (in_sil ? 1u : 0u))) & vmask;
        while (trans) {                                                                           // :524-533
            const int b = __ffs(trans) - 1;
            trans &= trans - 1;
            const int i2 = base + b;
            if (!in_sil) { in_sil = 1; start = i2; }
            else {
                in_sil = 0;
                const int e = min(i2 + k - 1, Tn);
                if (e - start >= k && n < max_out) {
                    if (lane == 0) { out[2 * n] = start; out[2 * n + 1] = e; }
                    ++n;
                }
            }
        }
    }
    __device__ __forceinline__ int finish(int k, int Tn, int lane, int32_t* out, int max_out) {
        if (in_sil && Tn - start >= k && n < max_out) {                                          // :536-539
            if (lane == 0) { out[2 * n] = start; out[2 * n + 1] = Tn; }
            ++n;
        }
        __syncwarp();
        return n;
    }
};

// Moving averages longer than the staged scan holds, or k == 1 (the probability itself): every prefix value straight from
// global memory.  Same arithmetic as detect_silence.
__device__ __noinline__ int detect_silence_direct(const double* D, const double* cbase, int lane, int r0, int Tn, float thr, int k,
                                                  int32_t* out, int max_out) {
    UttCtx c;
    c.D = D; c.cbase = const_cast<double*>(cbase);
    const int nwin = (k > 1) ? Tn - k + 1 : Tn;
    const double base0 = sil_prefix(c, r0 - 1);
    const float kf = (float)k;
    SilRuns runs;
    for (int base = 0; base < nwin; base += 32) {
        const int i = base + lane;
        float av = -INFINITY;
        if (i < nwin) {
            if (k > 1) {
                const float hi = (float)(sil_prefix(c, r0 + i + k - 1) - base0);
                const float lo = i == 0 ? 0.0f : (float)(sil_prefix(c, r0 + i - 1) - base0);
                av = (hi - lo) / kf;
            } else {
                av = (float)(sil_prefix(c, r0 + i) - sil_prefix(c, r0 + i - 1));
            }
        }
        runs.feed(__ballot_sync(FULL, av >= thr), base, nwin, k, Tn, lane, out, max_out);
    }
    return runs.finish(k, Tn, lane, out, max_out);
}

// _detect_silence_segments (:471-541) over rows [r0, r0+Tn).  Writes (start,end) pairs to out (lane 0) and returns the
// count (uniform).  The reference's k-frame moving average comes from a prefix sum of the range (torch.cumsum: accumulated
// in fp64, stored as fp32, :506-512); here every prefix value is a difference of two running sums of the utterance, rounded
// to fp32 the same way.  The running sums a slab of PLAN_SLAB windows needs come into shared memory with ONE bulk copy
// (one memory round trip per slab instead of one per group of windows).
__device__ int detect_silence(const UttCtx& c, int r0, int Tn, float thr, int k, int32_t* out, int max_out) {
    const BfaParams& p = c.a->p;
    if (p.silence_id >= c.a->C || c.D == nullptr) return 0;   // :497
    if (Tn < k || Tn <= 0) return 0;                            // :499
    if (k <= 1 || k > PLAN_KMAX) return detect_silence_direct(c.D, c.cbase, c.lane, r0, Tn, thr, k, out, max_out);
    const int lane = c.lane;
    const int nwin = Tn - k + 1;
    const double base0 = sil_prefix(c, r0 - 1);
    const float kf = (float)k;
    const double* cbs = c.ds + PLAN_DS;
    SilRuns runs;
    for (int w0 = 0; w0 < nwin; w0 += PLAN_SLAB) {
        // prefix value j of the range, padded[j] (:508-509), is fp32(sum of rows r0 .. r0+j-1) = fp32(S(r0+j-1) - S(r0-1));
        // the slab needs j in [w0, min(w0 + PLAN_SLAB + k, Tn)]
        const int t_lo = max(r0 + w0 - 1, 0);
        const int t_hi = min(r0 + w0 + PLAN_SLAB + k - 1, r0 + Tn - 1);
        const double* src = c.D + t_lo;
        const int skew = (int)(((unsigned long long)src >> 3) & 1ull);      // bulk copies start on 16-byte boundaries
        const int ch0 = t_lo >> SS_CHUNK_SHIFT;
        __syncwarp();                                                       // the previous slab's readers are done
        if (lane == 0) {
            const uint32_t bytes = ((uint32_t)(t_hi - t_lo + 1 + skew) * 8u + 15u) & ~15u;
            mbar_expect_tx(c.bar, bytes);
            bulk_g2s(smem_u32(c.ds), src - skew, bytes, c.bar);
        }
        if (lane < 4) c.ds[PLAN_DS + lane] = c.cbase[min(ch0 + lane, (c.T - 1) >> SS_CHUNK_SHIFT)];
        mbar_wait(c.bar, *c.phase & 1u);
        *c.phase ^= 1u;
        __syncwarp();
        const double* dsk = c.ds + (skew - t_lo);
        auto S = [&](int t) -> double {      // running sum through row t of the utterance (t >= 0), from the staged slab
            return dsk[t] + cbs[(t >> SS_CHUNK_SHIFT) - ch0];
        };
        const int wend = min(w0 + PLAN_SLAB, nwin);
#pragma unroll 2
        for (int base = w0; base < wend; base += 32) {
            const int i = base + lane;
            float av = -INFINITY;
            if (i < nwin) {
                const float hi = (float)(S(r0 + i + k - 1) - base0);
                const float lo = (r0 + i == 0) ? 0.0f : (float)(S(r0 + i - 1) - base0);
                av = (hi - lo) / kf;                                                               // :510-512
            }
            runs.feed(__ballot_sync(FULL, av >= thr), base, nwin, k, Tn, lane, out, max_out);       // :517
        }
    }
    return runs.finish(k, Tn, lane, out, max_out);
}

__device__ __forceinline__ void fill_frames(const UttCtx& c, long long o0, long long olim, int nf, int ph, int idx) {
    const long long ob = c.a->frame_off[c.u];
    for (int f = c.lane; f < nf; f += 32)
        if (o0 + f < olim) {
            c.a->frame_ph[o0 + f] = ph; c.a->frame_idx[o0 + f] = idx;
            if (c.a->path_lp && ph >= 0 && ph < c.a->C && o0 + f - ob < c.T) c.a->path_lp[o0 + f] = c.lp[(o0 + f - ob) * c.a->C + ph];
        }
}

__device__ __forceinline__ int band_rule(int L, int div, int floor_) { return (L > 60) ? max(L / div, floor_) : 0; }

// _segmented_viterbi_decode (:268-469).  Returns the number of local items on success, -1 when the
// reference returns ([], []) and the caller must fall back to a single Viterbi.
__device__ int plan_segmented(const UttCtx& c, Item* loc, int32_t* lists, uint32_t* anchors) {
    const PlanArgs& a = *c.a;
    const BfaParams& p = a.p;
    const int lane = c.lane, T = c.T, N = c.N, C = a.C;
    int32_t* grp = lists;                      // [2*gmax]
    int32_t* asil = grp + 2 * a.gmax;          // [2*amax]
    int32_t* mt = asil + 2 * a.amax;           // [2*gmax]
    int32_t* segs = mt + 2 * a.gmax;           // [5*(2*gmax+2)]
    int32_t* sub = segs + 5 * (2 * a.gmax + 2);// [2*amax]

    // Step 1: target SIL groups (:203-224): maximal runs of silence_id in the target, 32 targets per step
    int ng = 0;
    {
        int in_run = 0, start = 0;
        for (int base = 0; base < N; base += 32) {
            const int i = base + lane;
            const uint32_t bits = __ballot_sync(FULL, i < N && c.seq[i] == p.silence_id);
            const uint32_t vmask = (N - base >= 32) ? FULL : ((1u << (N - base)) - 1u);
            uint32_t trans = (bits ^ ((bits << 1) | (in_run ? 1u : 0u))) & vmask;
            while (trans) {
                const int pos = base + __ffs(trans) - 1;
                trans &= trans - 1;
                if (!in_run) { in_run = 1; start = pos; }
                else {
                    in_run = 0;
                    if (lane == 0) { grp[2 * ng] = start; grp[2 * ng + 1] = pos; }
                    ++ng;
                }
            }
        }
        if (in_run) {
            if (lane == 0) { grp[2 * ng] = start; grp[2 * ng + 1] = N; }
            ++ng;
        }
    }
    if (ng == 0) return -1;                                                       // :293-295
    if (c.D) {
        if (!a.sil_ready) {          // a wrong BFA_HINT_NO_SIL: nobody streamed this utterance's rows
            const bool sil_tgt = (c.mw[p.silence_id >> 5] >> (p.silence_id & 31)) & 1u;
            for (int r0 = 0; r0 < T; r0 += SS_CHUNK)
                ss_gather_rows(p, C, c.lp + (long long)r0 * C, min(SS_CHUNK, T - r0), a.D + a.frame_off[c.u] + r0, lane, c.mw, sil_tgt);
            __threadfence();
            asm volatile("fence.proxy.async.global;" ::: "memory");      // detect_silence reads these sums back with bulk copies
            __syncwarp();
        }
        chunk_bases(c);
    }
    int k = p.silence_anchors;                                                    // :296
    int na = detect_silence(c, 0, T, 0.9f, k, asil, a.amax);                      // :297
    if (na == 0 && N > 200) {                                                     // :298-304
        double nt = 1.0 - (0.09 * (double)k);
        if (nt < 0.05) nt = 0.05;
        na = detect_silence(c, 0, T, (float)nt, k, asil, a.amax);
    }
    if (na == 0 && N > 200 && k > 3) {                                            // :305-308
        k = 3;
        na = detect_silence(c, 0, T, 0.9f, k, asil, a.amax);
    }
    if (na == 0) return -1;                                                       // :315-320
    __syncwarp();

    // Step 2: _match_silences (:226-266)
    int nm = 0, audio_idx = 0;
    for (int g = 0; g < ng; ++g) {
        double tpos = (double)(grp[2 * g] + grp[2 * g + 1]) / 2.0 / (double)N;
        int best = -1;
        double bd = INFINITY;
        for (int ai = audio_idx; ai < na; ++ai) {
            double apos = (double)(asil[2 * ai] + asil[2 * ai + 1]) / 2.0 / (double)T;
            double d = fabs(tpos - apos);
            if (d < bd) { bd = d; best = ai; }
            else if (d > bd) break;
        }
        if (best >= 0 && bd < 0.3) {
            if (lane == 0) { mt[2 * nm] = g; mt[2 * nm + 1] = best; }
            ++nm;
            audio_idx = best + 1;
        }
    }
    if (nm == 0) return -1;                                                       // :324-325
    __syncwarp();

    // Step 3 + 3b: segment list with short-speech merge (:327-369); built in registers, lane 0 stores
    int ns = 0, pa = 0, pt = 0;
    auto push = [&](int a0, int a1, int t0, int t1, int sil) {
        // merge rule (:363-366): short non-silence segment with phonemes joins the previous entry
        if (!sil && (t1 - t0) > 0 && (a1 - a0) < p.min_speech_frames && ns > 0) {
            if (lane == 0) { segs[5 * (ns - 1) + 1] = a1; segs[5 * (ns - 1) + 3] = t1; segs[5 * (ns - 1) + 4] = 0; }
        } else {
            if (lane == 0) { int32_t* s = segs + 5 * ns; s[0] = a0; s[1] = a1; s[2] = t0; s[3] = t1; s[4] = sil; }
            ++ns;
        }
    };
    for (int m = 0; m < nm; ++m) {
        int tg0 = grp[2 * mt[2 * m]], tg1 = grp[2 * mt[2 * m] + 1];
        int as0 = asil[2 * mt[2 * m + 1]], as1 = asil[2 * mt[2 * m + 1] + 1];
        if (pa < as0 && pt < tg0) push(pa, as0, pt, tg0, 0);
        else if (pa < as0) push(pa, as0, pt, pt, 0);
        push(as0, as1, tg0, tg1, 1);
        pa = as1;
        pt = tg1;
    }
    if (pa < T && pt < N) push(pa, T, pt, N, 0);
    else if (pa < T) push(pa, T, pt, pt, 0);
    __syncwarp();

    // Step 4 (:377-451)
    const long long o_base = a.frame_off[c.u], o_lim = a.frame_off[c.u + 1];
    int w = 0, n_items = 0, anc_used = 0;
    for (int i = 0; i < ns; ++i) {
        const int a0 = segs[5 * i], a1 = segs[5 * i + 1], t0 = segs[5 * i + 2], t1 = segs[5 * i + 3], sil = segs[5 * i + 4];
        const int nf = a1 - a0;
        if (nf <= 0) continue;
        if (sil) {                                                                // :382-397
            const int nsil = t1 - t0;
            fill_frames(c, o_base + w, o_lim, nf, p.silence_id, -1);
            __syncwarp();
            if (nsil > 0) {
                double fps = (double)nf / (double)nsil;
                for (int q = 0; q < nsil; ++q) {
                    int f0 = (int)((double)q * fps), f1 = min((int)((double)(q + 1) * fps), nf);
                    for (int f = f0 + lane; f < f1; f += 32)
                        if (o_base + w + f < o_lim) a.frame_idx[o_base + w + f] = t0 + q;
                    __syncwarp();
                }
            }
            w += nf;
            continue;
        }
        const int n = t1 - t0;
        if (n == 0) {                                                             // :409-412
            fill_frames(c, o_base + w, o_lim, nf, p.blank_id, -1);
            w += nf;
            continue;
        }
        const int p0 = max(0, a0 - p.boundary_pad), p1 = min(T, a1 + p.boundary_pad);   // :401-403
        const int Ts = p1 - p0;
        // sub-silence anchoring (:415-419): counts per frame, 4 bits each
        const int nss = detect_silence(c, p0, Ts, 0.8f, k, sub, a.amax);
        const int words = (Ts + 7) / 8;
        int flags_anchor = 0;
        if (nss > 0 && anc_used + words <= a.anchor_words) {
            for (int wd = lane; wd < words; wd += 32) {
                uint32_t v = 0;
                for (int q = 0; q < 8; ++q) {
                    int f = wd * 8 + q, cnt = 0;
                    for (int s = 0; s < nss; ++s) cnt += (f >= sub[2 * s] && f < sub[2 * s + 1]);
                    v |= (uint32_t)min(cnt, 15) << (4 * q);
                }
                anchors[anc_used + wd] = v;
            }
            flags_anchor = ITEM_ANCHOR;
        }
        int stride = 4;                                                           // :423-426
        if ((double)(stride * n + 1) > (double)Ts * 0.9) stride = 3;
        if ((double)(stride * n + 1) > (double)Ts * 0.8) stride = 2;
        const int L = stride * n + 1;
        if ((double)L > (double)Ts * 1.2) return -1;                              // :427-429
        if (L > BFA_MAX_L) return -2;                                             // beyond the wide exact kernel's state capacity
        if (n_items >= a.item_cap) return -1;  // cannot happen (item_cap = gmax + 1); defensive
        if (lane == 0) {
            Item it;
            it.lp_off = a.row_off[c.u] + (long long)p0 * C;
            it.stat_off = a.frame_off[c.u] + p0;
            it.out_off = o_base + w;
            it.out_lim = o_lim;
            it.seq_off = a.tgt_off[c.u] + t0;
            it.T = Ts; it.L = L; it.band = band_rule(L, 3, 30);                   // :441
            it.stride = stride; it.n = n; it.idx0 = t0;
            it.trim = a0 - p0; it.n_out = nf;                                     // :447-448
            it.utt = c.u;
            it.anchor_off = (int)((size_t)c.u * a.anchor_words) + anc_used;
            it.flags = (p.boost_targets ? ITEM_STATS : 0) | (p.enforce_minimum ? ITEM_FLOOR : 0) | flags_anchor;
            it.lead = (int)(((unsigned long long)(a.logp + it.lp_off) & 15ull) >> 2);
            loc[n_items] = it;
        }
        if (flags_anchor) anc_used += words;
        ++n_items;
        w += nf;
    }
    if (w == 0) return -1;                                                        // :454-455
    // tail padding with blanks if short (:461-464); over-long output is clipped by out_lim
    if (w < T) fill_frames(c, o_base + w, o_lim, T - w, p.blank_id, -1);
    __syncwarp();
    return n_items;
}

__global__ void __launch_bounds__(256, 4) plan_kernel(const __grid_constant__ PlanArgs a) {
    int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    int n_items = 0;
    Item* loc = nullptr;
    Item single;                 // the common case, one item = the whole utterance, never leaves lane 0's registers
    bool have_single = false;
    bool tok = false;
    __shared__ uint32_t s_mask[8][MAX_WORDS];
    __shared__ __align__(16) double s_ds[8][PLAN_DS + 4];      // per warp: a slab of running sums + 4 chunk bases (detect_silence)
    __shared__ __align__(8) unsigned long long s_bar[8];
    uint32_t ds_phase = 0;
    if (lane == 0) {
        mbar_init(smem_u32(&s_bar[threadIdx.x >> 5]), 1);
        fence_mbar_init();
    }
    __syncwarp();
    pdl_release();               // the banded kernel's CTAs may take their SMs while this grid drains
    if (a.deferred) {            // the direct kernel went first: only what it handed back
        pdl_wait();
        const int nd = *a.n_deferred;
        if (blockIdx.x * (blockDim.x >> 5) >= nd) return;      // uniform over the CTA
        u = u < nd ? a.deferred[u] : a.B;
    }
    if (u < a.B) {
    const BfaParams& p = a.p;
    UttCtx c;
    c.a = &a; c.u = u; c.lane = lane;
    c.T = a.T[u];
    c.N = (int)(a.tgt_off[u + 1] - a.tgt_off[u]);
    c.lp = a.logp + a.row_off[u];
    c.seq = a.tgt + a.tgt_off[u];
    c.D = a.D ? a.D + a.frame_off[u] : nullptr;
    c.cbase = a.cbase + (size_t)u * a.cb_pitch;
    uint32_t* sw = s_mask[threadIdx.x >> 5];
    c.mw = sw;
    c.ds = s_ds[threadIdx.x >> 5];
    c.bar = smem_u32(&s_bar[threadIdx.x >> 5]);
    c.phase = &ds_phase;
    const TgtInfo ti = target_info(a.tgt, a.tgt_off[u], a.tgt_off[u + 1], a.C, p.blank_id, p.silence_id, lane, sw);
    if (lane < MAX_WORDS) a.tmask[(size_t)u * MAX_WORDS + lane] = sw[lane];      // consumed by the Viterbi kernels
    const int T = c.T, N = c.N;
    const long long o_base = a.frame_off[u], o_lim = a.frame_off[u + 1];
    loc = a.items_local + (size_t)u * a.item_cap;
    int st = BFA_ST_OK;
    if (lane == 0 && a.dp_final) a.dp_final[u] = 0.0f;

    if (N == 0) {                                                                 // :894-897 / :112-118
        fill_frames(c, o_base, o_lim, T, p.blank_id, -1);
        st = BFA_ST_EMPTY_TARGET;
    } else {
        bool done = false;
        // :133; without a SIL in the target _find_target_sil_groups is empty and the attempt returns [] at once (:293-295)
        if (p.mode == BFA_MODE_FULL && p.silence_anchors > 0 && p.silence_id >= 0 && T > 0 && ti.has_sil) {
            int r = plan_segmented(c, loc, a.lists + (size_t)u * a.list_ints, a.anchors + (size_t)u * a.anchor_words);
            if (r >= 0) { n_items = r; st = BFA_ST_SEGMENTED; done = true; }
            else if (r == -2) {          // a segment's path is longer than BFA_MAX_L states: this utterance is refused, the batch goes on
                st = BFA_ST_UNSUPPORTED;
                __syncwarp();
                fill_frames(c, o_base, o_lim, T, p.blank_id, -1);
                done = true;
            }
        }
        if (!done) {
            int stride = 4;
            bool ok = true;
            if (p.mode == BFA_MODE_SIMPLE) {                                      // :963-968
                if ((double)(stride * N + 1) > (double)T * 0.9) stride = 3;
                if ((double)(stride * N + 1) > (double)T * 0.8) stride = 2;
            } else {                                                              // :153-157
                if (stride * N + 1 > T) stride = 3;
                if (stride * N + 1 > T) stride = 2;
                if (stride * N + 1 > T) stride = 1;
                if (stride * N + 1 > T) {                                         // :159-176
                    ok = false;
                    if (T < N) {
                        st = BFA_ST_TOO_SHORT;
                        fill_frames(c, o_base, o_lim, T, p.blank_id, -1);
                    } else {
                        st = BFA_ST_PROPORTIONAL;
                        for (int t = lane; t < T; t += 32) {
                            int j = (int)(((long long)t * N) / T);
                            a.frame_ph[o_base + t] = c.seq[j];
                            a.frame_idx[o_base + t] = j;
                            if (a.path_lp && c.seq[j] >= 0 && c.seq[j] < a.C) a.path_lp[o_base + t] = c.lp[(long long)t * a.C + c.seq[j]];
                        }
                    }
                }
            }
            if (ok && T > 0 && stride * N + 1 > BFA_MAX_L) {   // more states than the wide exact kernel holds (N > 2047 at stride 4)
                ok = false;
                st = BFA_ST_UNSUPPORTED;
                fill_frames(c, o_base, o_lim, T, p.blank_id, -1);
            }
            if (ok && T > 0) {
                const int L = stride * N + 1;
                if (lane == 0) {
                    Item it;
                    it.lp_off = a.row_off[u];
                    it.stat_off = a.frame_off[u];
                    it.out_off = o_base; it.out_lim = o_lim;
                    it.seq_off = a.tgt_off[u];
                    it.T = T; it.L = L; it.band = band_rule(L, 4, 20);            // :190 / :976
                    it.stride = stride; it.n = N; it.idx0 = 0; it.trim = 0; it.n_out = T;
                    it.utt = u; it.anchor_off = 0;
                    it.flags = ITEM_FINAL;
                    if (p.mode == BFA_MODE_FULL)
                        it.flags |= (p.boost_targets ? ITEM_STATS : 0) | (p.enforce_minimum ? ITEM_FLOOR : 0);
                    it.lead = (int)(((unsigned long long)(a.logp + it.lp_off) & 15ull) >> 2);
                    single = it;
                }
                have_single = true;
                n_items = 1;
            }
        }
    }
    __syncwarp();
    if (lane == 0) a.status[u] = st;
    tok = ti.all_ok;
    }   // u < B

    // ---- publish the items into the compact global lists (banded kernels / exact kernel).  Counts are aggregated per
    //      CTA in shared memory first: one global atomic per list per CTA instead of one per utterance ----
    __shared__ int s_cnt[4], s_base[4], s_frm[4];
    if (threadIdx.x < 4) { s_cnt[threadIdx.x] = 0; s_frm[threadIdx.x] = 0; }
    __syncthreads();
    int my_cnt[4] = {0, 0, 0, 0}, my_off[4] = {0, 0, 0, 0};
    int cl0 = -2;                // class of item `lane` (the first, usually the only, pass over the items)
    for (int i0 = 0; i0 < n_items; i0 += 32) {
        const int i = i0 + lane;
        int cl = -2;
        if (i < n_items) cl = have_single ? fast_class(single, a.C, a.logp, tok, a.fast_enable) : fast_class(loc[i], a.C, a.logp, tok, a.fast_enable);
        if (i0 == 0) cl0 = cl;
        const int Ti = (i < n_items) ? (have_single ? single.T : loc[i].T) : 0;
#pragma unroll
        for (int c = -1; c <= 2; ++c) {
            my_cnt[c + 1] += __popc(__ballot_sync(FULL, cl == c));
            if (c >= 0) {
                const int fr = __reduce_add_sync(FULL, cl == c ? Ti : 0);
                if (fr && lane == 0) atomicAdd(&s_frm[c + 1], fr);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (lane == 0 && my_cnt[c]) my_off[c] = atomicAdd(&s_cnt[c], my_cnt[c]);
        my_off[c] = __shfl_sync(FULL, my_off[c], 0);
    }
    __syncthreads();
    if (threadIdx.x < 4 && s_cnt[threadIdx.x]) {
        int* cnt = (threadIdx.x == 0) ? a.n_items : a.n_fast[threadIdx.x - 1];
        s_base[threadIdx.x] = atomicAdd(cnt, s_cnt[threadIdx.x]);
        if (threadIdx.x > 0 && s_frm[threadIdx.x]) atomicAdd(a.frames_fast[threadIdx.x - 1], s_frm[threadIdx.x]);
    }
    __syncthreads();
    for (int i0 = 0; i0 < n_items; i0 += 32) {
        const int i = i0 + lane;
        int cl = -2;
        if (i0 == 0) cl = cl0;
        else if (i < n_items) cl = fast_class(loc[i], a.C, a.logp, tok, a.fast_enable);
#pragma unroll
        for (int c = -1; c <= 2; ++c) {
            const unsigned m = __ballot_sync(FULL, cl == c);
            Item* dst = (c < 0) ? a.items : a.fast_items[c];
            if (cl == c) {
                Item* d = dst + (s_base[c + 1] + my_off[c + 1] + __popc(m & ((1u << lane) - 1u)));
                if (have_single) *d = single;
                else *d = loc[i];
            }
            my_off[c + 1] += __popc(m);
        }
    }
}

// Build explicit-path items from parallel arrays (bfa_viterbi_paths entry).
__global__ void items_from_arrays_kernel(int n, int C, const long long* row_off, const int32_t* T, const long long* path_off,
                                         const int32_t* L, const int32_t* band, const long long* frame_off, Item* items,
                                         int* n_items) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *n_items = n;
    if (i >= n) return;
    Item it;
    it.lp_off = row_off[i]; it.stat_off = 0; it.out_off = frame_off[i]; it.out_lim = frame_off[i] + T[i];
    it.seq_off = path_off[i];
    it.T = T[i]; it.L = L[i]; it.band = band[i]; it.stride = 0; it.n = 0; it.idx0 = 0; it.trim = 0; it.n_out = T[i];
    it.utt = i; it.anchor_off = 0; it.flags = ITEM_FINAL; it.lead = 0;
    items[i] = it;
}

}  // namespace bfa
