// assort.cuh -- run-length timestamps + per-phoneme confidence, one warp per utterance.
//
// Restates ViterbiDecoder.assort_frames (forced_alignment.py:777-834) and
// utils._calculate_confidences (utils.py:70-113).
#pragma once
#include <type_traits>

#include "bfa_common.cuh"

namespace bfa {

// utils.py:84-111 for one stamp; sequential fp32 accumulation in frame order like the reference.
// NB utils.py:89: avg_confidence is a VIEW of probs[start, ph]; the in-place += and /= also
// rewrite that element, so the max of :107 runs over {avg, p[start+1 .. end-1]}.
// `lse` (optional): the rows hold un-normalised logits and lse[f] is row f's log-sum-exp (bfa_align_batch_logits' row_lse).
__device__ __forceinline__ float stamp_confidence(const float* lp, int T, int C, int ph, int start, int end, const float* lse = nullptr) {
    int s = max(0, start), e = min(T, end);                     // :86-87
    if (s >= T || ph < 0 || ph >= C) return __int_as_float(0x7fc00000);   // the reference raises IndexError here (probs[start, ph], :89); never read out of bounds
    float avg = expf(lp[(long long)s * C + ph] - (lse ? lse[s] : 0.0f));   // :89
    if (s < e && ph < C) {                                      // :93
        const float half = avg / 2.0f;                          // :95
        int good = 1;
        float mx = 0.f;
        for (int f = s + 1; f < e; ++f) {                       // :99-103
            float pr = expf(lp[(long long)f * C + ph] - (lse ? lse[f] : 0.0f));
            mx = fmaxf(mx, pr);
            if (pr > half || pr > 0.1f) { avg += pr; ++good; }
        }
        if (good > 1) {                                         // :104-109
            avg /= (float)good;
            mx = fmaxf(mx, avg);
            if (avg < mx / 2.0f) avg = mx;
        }
    }
    return avg;
}

// Same as stamp_confidence, but lp[f, ph] comes from the per-frame array the Viterbi kernels gathered
// (valid because every frame of a stamp was assigned the stamp's phoneme).
__device__ __forceinline__ float stamp_confidence_path(const float* plp, int T, int start, int end, const float* lse = nullptr) {
    int s = max(0, start), e = min(T, end);
    float avg = expf(plp[s] - (lse ? lse[s] : 0.0f));
    if (s < e) {
        const float half = avg / 2.0f;
        int good = 1;
        float mx = 0.f;
        for (int f = s + 1; f < e; ++f) {
            float pr = expf(plp[f] - (lse ? lse[f] : 0.0f));
            mx = fmaxf(mx, pr);
            if (pr > half || pr > 0.1f) { avg += pr; ++good; }
        }
        if (good > 1) {
            avg /= (float)good;
            mx = fmaxf(mx, avg);
            if (avg < mx / 2.0f) avg = mx;
        }
    }
    return avg;
}

// Same again on probabilities that are already exponentiated (the staged path below).
__device__ __forceinline__ float stamp_confidence_prob(const float* pr_s, int T, int start, int end) {
    int s = max(0, start), e = min(T, end);
    float avg = pr_s[s];
    if (s < e) {
        const float half = avg / 2.0f;
        int good = 1;
        float mx = 0.f;
#pragma unroll 4
        for (int f = s + 1; f < e; ++f) {
            const float pr = pr_s[f];
            mx = fmaxf(mx, pr);
            if (pr > half || pr > 0.1f) { avg += pr; ++good; }
        }
        if (good > 1) {
            avg /= (float)good;
            mx = fmaxf(mx, avg);
            if (avg < mx / 2.0f) avg = mx;
        }
    }
    return avg;
}

constexpr int ASSORT_WARPS = 4;
constexpr int ASSORT_R = 6;           // rounds of 32 x 32 frames the staged run-length sweep keeps in registers
constexpr int ASSORT_TS_MAX = ASSORT_R * 1024;   // longest utterance whose per-frame arrays are staged in shared memory (8 bytes per frame per warp)
constexpr int ASSORT_SS_MAX = 512;    // largest stamp pitch assembled in shared memory
// staged frames / stamps per warp for a batch shape, and the dynamic shared memory of one CTA
__host__ __device__ inline int assort_ts(int max_T) { return max_T <= ASSORT_TS_MAX ? (max_T + 31) / 32 * 32 : 0; }
__host__ __device__ inline int assort_ss(int max_stamps) { return max_stamps <= ASSORT_SS_MAX ? max_stamps : 0; }
__host__ __device__ inline size_t assort_smem(int ts, int ss) { return (size_t)ASSORT_WARPS * ((size_t)8 * ts + (size_t)16 * ss); }

struct AssortArgs {
    BfaParams p;
    int B, C, max_stamps;
    int ts, ss;        // shared-memory staging per warp: frames (0: read global memory) and stamps (0: assemble in place)
    const float* logp;
    const long long* row_off;
    const int32_t* T;
    const long long* frame_off;
    const int32_t* frame_ph;
    const int32_t* frame_idx;
    int32_t* status;
    BfaStamp* stamps;
    float* conf;       // may be null
    int32_t* n_stamps;
    const float* path_lp;   // per-frame gathered lp (or null: gather from logp)
    const int32_t* uflag;   // when non-null: utterances with uflag[u] != 0 were finished (stamps included) by the direct kernel
    const float* row_lse;   // logits in: [total_frames] log-sum-exp of the rows (silprob_kernel<true>), subtracted where a value is exponentiated; or null
};

__global__ void __launch_bounds__(ASSORT_WARPS * 32) assort_confidence_kernel(AssortArgs a) {
    const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    pdl_wait();                   // the frame labels of the Viterbi kernels
    if (u >= a.B) return;
    if (a.uflag && a.uflag[u]) return;
    const int st = a.status[u] & 7;
    if (st == BFA_ST_EMPTY_TARGET || st == BFA_ST_TOO_SHORT || st == BFA_ST_UNSUPPORTED) {   // :894-897 -> [] ; ValueError ; refused
        if (lane == 0) a.n_stamps[u] = 0;
        return;
    }
    const int T = a.T[u];
    const long long fo = a.frame_off[u];
    const int32_t* ph = a.frame_ph + fo;
    const int32_t* ix = a.frame_idx + fo;
    const float* plp = a.path_lp ? a.path_lp + fo : nullptr;
    const float* lse = a.row_lse ? a.row_lse + fo : nullptr;
    // Stage the utterance's per-frame arrays in shared memory with independent coalesced loads (all in flight at
    // once), exponentiating the confidence inputs on the way in (one frame per lane instead of one stamp per lane);
    // the run-length scan and the per-stamp confidence loops below then never wait on global memory.
    extern __shared__ __align__(16) int32_t assort_smem_raw[];
    int32_t* wsm = assort_smem_raw + (size_t)(threadIdx.x >> 5) * (2 * a.ts + 4 * a.ss);
    const float* pr_s = nullptr;                    // exp(plp[t]) when staged
    const int32_t* pk_s = nullptr;                  // (idx << 16) | phoneme when staged (C <= 256, idx < 32768)
    if (T <= a.ts) {
        int32_t* pk = wsm;
        float* lp_s = reinterpret_cast<float*>(pk + a.ts);
        constexpr int UNR = 10;                     // 3 x 10 independent loads per lane before the first use
        // (two copies of the loop, chosen once per utterance: a fourth, conditional load inside the first loop keeps the
        // compiler from putting all loads in flight before the first use -- measured 31 -> 85 us on config 3)
        auto stage = [&](auto with_lse) {
            constexpr bool LSE = decltype(with_lse)::value;
            for (int t0 = lane; t0 < T; t0 += 32 * UNR) {
                int32_t v_ph[UNR], v_ix[UNR];
                float v_lp[UNR], v_ls[LSE ? UNR : 1];
#pragma unroll
                for (int j = 0; j < UNR; ++j) {
                    const int t = t0 + 32 * j;
                    if (t < T) {
                        v_ph[j] = ph[t];
                        v_ix[j] = ix[t];
                        if (plp) v_lp[j] = plp[t];
                        if constexpr (LSE) v_ls[j] = lse[t];
                    }
                }
#pragma unroll
                for (int j = 0; j < UNR; ++j) {
                    const int t = t0 + 32 * j;
                    if (t < T) {
                        pk[t] = (v_ix[j] << 16) | (v_ph[j] & 0xffff);
                        if (plp) lp_s[t] = expf(LSE ? v_lp[j] - v_ls[j] : v_lp[j]);
                    }
                }
            }
        };
        if (lse) stage(std::true_type{});
        else stage(std::false_type{});
        __syncwarp();
        pk_s = pk;
        if (plp) pr_s = lp_s;
    }
    BfaStamp* gout = a.stamps + (size_t)u * a.max_stamps;
    BfaStamp* out = a.ss ? reinterpret_cast<BfaStamp*>(wsm + 2 * a.ts) : gout;   // stamps are assembled here
    const int blank = a.p.blank_id;
    // Pass 1: every run start emits a provisional stamp (end filled by the next run start).
    // A run is emitted if non-blank (:830-831), or blank, !ignore_noise and longer than max_blanks
    // (:819-827) -- the length test needs the end, so blank candidates are compacted in pass 2.
    int n = 0;            // provisional stamps so far (uniform)
    int open = -1;        // index of the provisional stamp whose end is still unknown
    if (pk_s) {
        // Staged utterances (T <= ASSORT_TS_MAX: ASSORT_R rounds of 32 blocks of 32 frames).  (A) one sweep turns the packed
        // (idx, phoneme) words into run-start / candidate bit masks, block j kept by lane j % 32 in round j / 32; (B) every lane
        // emits the stamps that start in its own blocks: slots from a prefix count over the blocks, ends from the next run
        // start (same mask, a later block of the same round found with one indexed shuffle, or the first run start of the
        // next non-empty round, which is uniform).
        uint32_t sb[ASSORT_R], cb[ASSORT_R];
#pragma unroll
        for (int r = 0; r < ASSORT_R; ++r) { sb[r] = 0u; cb[r] = 0u; }
        const int nb = (T + 31) >> 5;
        int32_t carry = 0;
#pragma unroll
        for (int r = 0; r < ASSORT_R; ++r) {         // the round index is static: the masks stay in registers
            const int b_end = min(32, nb - r * 32);
            for (int bb = 0; bb < b_end; ++bb) {
                const int t = (r * 32 + bb) * 32 + lane;
                const int32_t w = t < T ? pk_s[t] : 0;
                int32_t wp = __shfl_up_sync(FULL, w, 1);
                if (lane == 0) wp = carry;
                carry = __shfl_sync(FULL, w, 31);
                const bool startf = t < T && (t == 0 || w != wp);                                  // :798-801
                const bool cand = startf && ((w & 0xffff) != blank || !a.p.ignore_noise);
                const uint32_t sbits = __ballot_sync(FULL, startf), cbits = __ballot_sync(FULL, cand);
                if (lane == bb) { sb[r] = sbits; cb[r] = cbits; }
            }
        }
        // exclusive prefix of the candidate counts over the blocks, round by round
        const int nr = (nb + 31) >> 5;       // rounds in use (uniform)
        int ex[ASSORT_R], tot = 0;
#pragma unroll
        for (int r = 0; r < ASSORT_R; ++r) {
            ex[r] = tot;
            if (r >= nr) continue;
            const int c = __popc(cb[r]);
            int inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += o;
            }
            ex[r] = tot + inc - c;
            tot += __shfl_sync(FULL, inc, 31);
        }
        n = tot;
        // per round: which blocks hold a run start, and the round's first run start (uniform; -1: none)
        uint32_t nz[ASSORT_R];
        int first[ASSORT_R];
#pragma unroll
        for (int r = 0; r < ASSORT_R; ++r) {
            nz[r] = 0u; first[r] = -1;
            if (r >= nr) continue;
            nz[r] = __ballot_sync(FULL, sb[r] != 0u);
            const int src = nz[r] ? __ffs(nz[r]) - 1 : 0;
            const uint32_t m = __shfl_sync(FULL, sb[r], src);
            first[r] = nz[r] ? (r * 32 + src) * 32 + __ffs(m) - 1 : -1;
        }
        int after = T;               // first run start of the rounds after r (walking r downwards)
#pragma unroll
        for (int r = ASSORT_R - 1; r >= 0; --r) {
            if (r >= nr) continue;
            const uint32_t later_same = nz[r] & ~((2u << lane) - 1u);                    // blocks after mine in my round
            const int src = later_same ? __ffs(later_same) - 1 : 0;
            const uint32_t m = __shfl_sync(FULL, sb[r], src);
            const int next_start = later_same ? (r * 32 + src) * 32 + __ffs(m) - 1 : after;
            uint32_t cbl = cb[r];
            int slot = ex[r];
            const int blk = r * 32 + lane;
            while (cbl) {
                const int pos = __ffs(cbl) - 1;
                cbl &= cbl - 1u;
                if (slot < a.max_stamps) {
                    const int t = blk * 32 + pos;
                    const int32_t w = pk_s[t];
                    const uint32_t later = sb[r] & ~((2u << pos) - 1u);     // run starts after this one in the same block
                    BfaStamp sv;
                    sv.phoneme = w & 0xffff;
                    sv.start = t;
                    sv.end = later ? blk * 32 + __ffs(later) - 1 : next_start;
                    sv.target_idx = w >> 16;   // runs are constant in idx, the :812-816 search is a no-op
                    out[slot] = sv;
                }
                ++slot;
            }
            if (first[r] >= 0) after = first[r];
        }
        __syncwarp();
    } else
    for (int base = 0; base < T; base += 32) {
        int t = base + lane;
        bool startf = false, cand = false;
        int p_t = blank, i_t = -1;
        if (t < T) {
            p_t = ph[t]; i_t = ix[t];
            startf = (t == 0) || p_t != ph[t - 1] || i_t != ix[t - 1];          // :798-801
            cand = startf && (p_t != blank || !a.p.ignore_noise);
        }
        uint32_t sbits = __ballot_sync(FULL, startf), cbits = __ballot_sync(FULL, cand);
        // close the open stamp at the first run start of this block
        if (open >= 0 && sbits) {
            if (lane == 0) out[open].end = base + __ffs(sbits) - 1;
            open = -1;
        }
        if (cand) {
            int slot = n + __popc(cbits & ((1u << lane) - 1u));
            if (slot < a.max_stamps) {
                uint32_t later = sbits & ~((2u << lane) - 1u);       // run starts after this lane
                BfaStamp sv;
                sv.phoneme = p_t;
                sv.start = t;
                sv.end = later ? base + __ffs(later) - 1 : -1;
                sv.target_idx = i_t;   // runs are constant in idx, the :812-816 search is a no-op
                out[slot] = sv;
            }
        }
        if (cbits) {
            int last = 31 - __clz(cbits);
            uint32_t later = (last == 31) ? 0u : (sbits & ~((2u << last) - 1u));
            int cnt = __popc(cbits);
            if (!later) open = min(n + cnt - 1, a.max_stamps - 1);
            n += cnt;
        }
        __syncwarp();
    }
    if (n > a.max_stamps) {   // caller's stamp pitch too small: flag it, keep the first max_stamps
        if (lane == 0) atomicOr(&a.status[u], BFA_ST_STAMP_OVERFLOW);
        n = a.max_stamps;
    }
    __syncwarp();
    if (open >= 0 && lane == 0) out[open].end = T;
    __syncwarp();
    // Pass 2 (only when blanks may be kept): drop blank runs that are not longer than max_blanks
    if (!a.p.ignore_noise) {
        int m = 0;
        for (int base = 0; base < n; base += 32) {
            int i = base + lane;
            BfaStamp s = {0, 0, 0, 0};
            bool keep = false;
            if (i < n) { s = out[i]; keep = (s.phoneme != blank) || (s.end - s.start > a.p.max_blanks); }
            uint32_t kb = __ballot_sync(FULL, keep);
            __syncwarp();
            if (keep) out[m + __popc(kb & ((1u << lane) - 1u))] = s;   // m + rank <= i: never overtakes unread slots
            __syncwarp();
            m += __popc(kb);
        }
        n = m;
    }
    if (lane == 0) a.n_stamps[u] = n;
    const float* lp = a.logp ? a.logp + a.row_off[u] : nullptr;
    for (int i = lane; i < n; i += 32) {
        const BfaStamp s = out[i];
        if (a.ss) gout[i] = s;                       // one 16-byte store per stamp
        if (a.conf) {
            float cf;
            if (pr_s && s.phoneme < a.C) cf = stamp_confidence_prob(pr_s, T, s.start, s.end);
            else if (plp && s.phoneme < a.C) cf = stamp_confidence_path(plp, T, s.start, s.end, lse);
            else cf = stamp_confidence(lp, T, a.C, s.phoneme, s.start, s.end, lse);
            a.conf[(size_t)u * a.max_stamps + i] = cf;
        }
    }
}

// stand-alone confidence entry (bfa_confidence_batch)
__global__ void confidence_kernel(int B, int C, const float* logp, const long long* row_off, const int32_t* Tc,
                                  const BfaStamp* stamps, const int32_t* n_stamps, int max_stamps, float* conf,
                                  const float* row_lse, const long long* lse_off) {
    const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (u >= B) return;
    const float* lp = logp + row_off[u];
    const float* lse = row_lse ? row_lse + lse_off[u] : nullptr;
    const int n = min(n_stamps[u], max_stamps);
    for (int i = lane; i < n; i += 32) {
        BfaStamp s = stamps[(size_t)u * max_stamps + i];
        conf[(size_t)u * max_stamps + i] = stamp_confidence(lp, Tc[u], C, s.phoneme, s.start, s.end, lse);
    }
}

// PhonemeTimestampAligner.extend_soft_boundaries_func (core.py:682-809), the step between decode_alignments and
// _calculate_confidences in the reference pipeline (core.py:925-937): four passes that stretch each stamp's start / end
// over neighbouring frames while exp(lp[f, phoneme]) stays above a threshold.  Within a pass a stamp only reads what
// EARLIER passes wrote to its neighbours, so every pass is data-parallel over the stamps: one warp per utterance, the
// stamps dealt to the lanes, __syncwarp between passes.  Thresholds are doubles like the reference's Python floats.
constexpr int SOFT_WARPS = 4;
__global__ void __launch_bounds__(SOFT_WARPS * 32) soft_boundaries_kernel(int B, int C, const float* __restrict__ logp,
                                                                          const long long* row_off, const int32_t* T, BfaStamp* stamps,
                                                                          const int32_t* n_stamps, int max_stamps, double t1, double t2,
                                                                          const float* __restrict__ row_lse, const long long* lse_off) {
    extern __shared__ double soft_thr[];                         // [SOFT_WARPS][max_stamps]: min(mean * t1, t1) per stamp
    const int u = blockIdx.x * SOFT_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (u >= B) return;
    const float* lp = logp + row_off[u];
    const int Tu = T[u], n = min(n_stamps[u], max_stamps);
    BfaStamp* st = stamps + (size_t)u * max_stamps;
    double* thr = soft_thr + (size_t)(threadIdx.x >> 5) * max_stamps;
    const float* lse = row_lse ? row_lse + lse_off[u] : nullptr;   // logits in: lse[f] = log-sum-exp of row f (bfa_align_batch_logits)
    auto prob = [&](int f, int ph) { return (double)expf(lp[(size_t)f * C + ph] - (lse ? lse[f] : 0.0f)); };
    for (int i = lane; i < n; i += 32) {                         // mean probability over the ORIGINAL stamp (:710-716)
        const BfaStamp s = st[i];
        double m = 0.001;
        if (s.start < Tu && s.phoneme < C && s.start < s.end) {
            const int e = min(s.end, Tu);
            double acc = 0.0;
            for (int f = s.start; f < e; ++f) acc += prob(f, s.phoneme);
            m = (double)(float)(acc / (double)(e - s.start));
        }
        thr[i] = fmin(m * t1, t1);
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {                         // pass 1: starts, strict (:719-737)
        const BfaStamp s = st[i];
        if (s.start >= Tu || s.phoneme >= C) continue;
        int lo = max(0, (int)((double)s.start - (double)(s.end - s.start) * 10.0));
        if (i > 0) lo = max(lo, min(st[i - 1].end + 10, s.start));
        int ns = s.start;
        for (int f = s.start - 1; f >= lo; --f) { if (prob(f, s.phoneme) >= thr[i]) ns = f; else break; }
        st[i].start = ns;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {                         // pass 2: ends, strict (:740-757)
        const BfaStamp s = st[i];
        if (s.start >= Tu || s.phoneme >= C) continue;
        int hi = min(Tu, (int)((double)s.end + (double)(s.end - s.start) * 10.0));
        if (i + 1 < n) hi = min(hi, min(s.end, st[i + 1].start - 10));     // :749 as written
        int ne = s.end;
        for (int f = s.end; f < hi; ++f) { if (prob(f, s.phoneme) >= thr[i]) ne = f + 1; else break; }
        st[i].end = ne;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {                         // pass 3: starts, lenient, up to the previous end (:760-780)
        const BfaStamp s = st[i];
        if (s.start >= Tu || s.phoneme >= C) continue;
        const int lo = i > 0 ? st[i - 1].end : 0;
        if (s.start <= lo) continue;
        int ns = s.start;
        for (int f = s.start - 1; f >= lo; --f) { if (prob(f, s.phoneme) >= t2) ns = f; else break; }
        st[i].start = ns;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {                         // pass 4: ends, lenient, up to the next start (:784-805)
        const BfaStamp s = st[i];
        if (s.start >= Tu || s.phoneme >= C) continue;
        int hi = min(Tu, (int)((double)s.end + (double)(s.end - s.start) * 10.0));
        if (i + 1 < n) hi = min(hi, st[i + 1].start);
        int ne = s.end;
        for (int f = s.end; f < hi; ++f) { if (prob(f, s.phoneme) >= t2) ne = f + 1; else break; }
        st[i].end = ne;
    }
}

}  // namespace bfa
