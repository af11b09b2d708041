// assort.cuh -- run-length timestamps + per-phoneme confidence, one warp per utterance.
//
// Restates ViterbiDecoder.assort_frames (forced_alignment.py:777-834) and
// utils._calculate_confidences (utils.py:70-113).
#pragma once
#include "bfa_common.cuh"

namespace bfa {

// utils.py:84-111 for one stamp; sequential fp32 accumulation in frame order like the reference.
// NB utils.py:89: avg_confidence is a VIEW of probs[start, ph]; the in-place += and /= also
// rewrite that element, so the max of :107 runs over {avg, p[start+1 .. end-1]}.
__device__ __forceinline__ float stamp_confidence(const float* lp, int T, int C, int ph, int start, int end) {
    int s = max(0, start), e = min(T, end);                     // :86-87
    float avg = expf(lp[(long long)s * C + ph]);                // :89
    if (s < e && ph < C) {                                      // :93
        const float half = avg / 2.0f;                          // :95
        int good = 1;
        float mx = 0.f;
        for (int f = s + 1; f < e; ++f) {                       // :99-103
            float pr = expf(lp[(long long)f * C + ph]);
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
__device__ __forceinline__ float stamp_confidence_path(const float* plp, int T, int start, int end) {
    int s = max(0, start), e = min(T, end);
    float avg = expf(plp[s]);
    if (s < e) {
        const float half = avg / 2.0f;
        int good = 1;
        float mx = 0.f;
        for (int f = s + 1; f < e; ++f) {
            float pr = expf(plp[f]);
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

constexpr int ASSORT_TS = 1024;    // frames per utterance staged in shared memory (longer utterances read global memory)
constexpr int ASSORT_WARPS = 4;
constexpr int ASSORT_SMEM = ASSORT_WARPS * 3 * ASSORT_TS * 4;

struct AssortArgs {
    BfaParams p;
    int B, C, max_stamps;
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
};

__global__ void assort_confidence_kernel(AssortArgs a) {
    const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (u >= a.B) return;
    const int st = a.status[u] & 7;
    if (st == BFA_ST_EMPTY_TARGET || st == BFA_ST_TOO_SHORT) {   // :894-897 -> [] ; ValueError
        if (lane == 0) a.n_stamps[u] = 0;
        return;
    }
    const int T = a.T[u];
    const int32_t* ph = a.frame_ph + a.frame_off[u];
    const int32_t* ix = a.frame_idx + a.frame_off[u];
    const float* plp = a.path_lp ? a.path_lp + a.frame_off[u] : nullptr;
    // Stage the utterance's per-frame arrays in shared memory with independent coalesced loads (all in flight at
    // once); the run-length scan and the per-stamp confidence loops below then never wait on global memory.
    extern __shared__ int32_t assort_smem[];
    if (T <= ASSORT_TS) {
        int32_t* ph_s = assort_smem + (threadIdx.x >> 5) * 3 * ASSORT_TS;
        int32_t* ix_s = ph_s + ASSORT_TS;
        float* lp_s = reinterpret_cast<float*>(ix_s + ASSORT_TS);
#pragma unroll 4
        for (int t = lane; t < T; t += 32) {
            ph_s[t] = ph[t];
            ix_s[t] = ix[t];
            if (plp) lp_s[t] = plp[t];
        }
        __syncwarp();
        ph = ph_s; ix = ix_s;
        if (plp) plp = lp_s;
    }
    BfaStamp* out = a.stamps + (size_t)u * a.max_stamps;
    const int blank = a.p.blank_id;
    // Pass 1: every run start emits a provisional stamp (end filled by the next run start).
    // A run is emitted if non-blank (:830-831), or blank, !ignore_noise and longer than max_blanks
    // (:819-827) -- the length test needs the end, so blank candidates are compacted in pass 2.
    int n = 0;            // provisional stamps so far (uniform)
    int open = -1;        // index of the provisional stamp whose end is still unknown
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
                out[slot].phoneme = p_t;
                out[slot].start = t;
                out[slot].target_idx = i_t;   // runs are constant in idx, the :812-816 search is a no-op
                out[slot].end = later ? base + __ffs(later) - 1 : -1;
            }
        }
        if (cbits) {
            int last = 31 - __clz(cbits);
            uint32_t later = (last == 31) ? 0u : (sbits & ~((2u << last) - 1u));
            int cnt = __popc(cbits);
            if (!later) open = min(n + cnt - 1, a.max_stamps - 1);
            n += cnt;
        }
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
    if (a.conf) {
        const float* lp = a.logp + a.row_off[u];
        for (int i = lane; i < n; i += 32) {
            BfaStamp s = out[i];
            a.conf[(size_t)u * a.max_stamps + i] = (plp && s.phoneme < a.C)
                                                       ? stamp_confidence_path(plp, T, s.start, s.end)
                                                       : stamp_confidence(lp, T, a.C, s.phoneme, s.start, s.end);
        }
    }
}

// stand-alone confidence entry (bfa_confidence_batch)
__global__ void confidence_kernel(int B, int C, const float* logp, const long long* row_off, const int32_t* Tc,
                                  const BfaStamp* stamps, const int32_t* n_stamps, int max_stamps, float* conf) {
    const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (u >= B) return;
    const float* lp = logp + row_off[u];
    const int n = min(n_stamps[u], max_stamps);
    for (int i = lane; i < n; i += 32) {
        BfaStamp s = stamps[(size_t)u * max_stamps + i];
        conf[(size_t)u * max_stamps + i] = stamp_confidence(lp, Tc[u], C, s.phoneme, s.start, s.end);
    }
}

}  // namespace bfa
