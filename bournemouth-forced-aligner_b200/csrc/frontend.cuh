// frontend.cuh -- what sits directly in front of and behind the alignment in the reference's batch path:
//   * stich_window_predictions (cupe2i/windowing.py:103-173) + F.log_softmax (core.py:898-899): the acoustic model's
//     per-window logits are cross-faded into one frame sequence and normalised; here ONE pass over the window logits writes
//     the log-posteriors the aligner consumes (the reference runs a stitch pass and a softmax pass over [B, T, C]);
//   * ViterbiDecoder._calculate_alignment_score (forced_alignment.py:767-773): sum over frames of lp[t, frame_phoneme[t]].
#pragma once
#include "bfa_common.cuh"

namespace bfa {

struct StitchArgs {
    int B, W, fpw, C, total_frames, sf;      // sf = fpw / 2 (stride in frames); fpw == 0: no stitching, `in` holds [B, total_frames, C] logits
    const float* in;                          // [B, W, fpw, C] window logits
    const float* weights;                     // [fpw] the cross-fade window (the caller's torch.cos(torch.linspace(...)), :130)
    float* out;                               // [B, total_frames, C] log-posteriors
    long long in_pitch_b, out_pitch_b;        // elements between consecutive utterances
};

// One warp per output frame, classes dealt to the lanes (lane + 32 i).  The cross-fade follows the reference's arithmetic
// operation by operation (product, then sum in ascending window order, :147-166; weight_sum + 1e-8, :169); the
// log-softmax is max-subtracted like torch's.
constexpr int STITCH_WARPS = 8;
// Plain log-softmax of [B, total_frames, C] logits (no windows): the same arithmetic as the stitched path below (max-subtracted,
// expf / logf, classes dealt to the lanes), but one row per warp and iteration keeps only C * 4 bytes per warp in flight --
// far too little to cover HBM latency.  Here every warp takes LS_ROWS rows per iteration, all their loads issued before the
// first reduction.
// NW = 32-class words per row that are compiled in (C <= 32 NW): lanes do not issue predicated-off work for classes that do not exist.
constexpr int LS_ROWS = 4;
template <int NW>
__global__ void __launch_bounds__(STITCH_WARPS * 32) log_softmax_rows_kernel(const __grid_constant__ StitchArgs a) {
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)a.B * a.total_frames;
    const long long nw = (long long)gridDim.x * STITCH_WARPS;
    for (long long r0 = ((long long)blockIdx.x * STITCH_WARPS + (threadIdx.x >> 5)) * LS_ROWS; r0 < rows; r0 += nw * LS_ROWS) {
        float v[LS_ROWS][NW];
#pragma unroll
        for (int r = 0; r < LS_ROWS; ++r) {
            const unsigned row = (unsigned)min(r0 + r, rows - 1);            // rows < 2^31 (checked by the caller): 32-bit division
            const unsigned b = row / (unsigned)a.total_frames, f = row - b * (unsigned)a.total_frames;
            const float* src = a.in + (long long)b * a.in_pitch_b + (long long)f * a.C;
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                const int c = lane + 32 * i;
                v[r][i] = c < a.C ? __ldcs(src + c) : -INFINITY;
            }
        }
#pragma unroll
        for (int r = 0; r < LS_ROWS; ++r) {
            if (r0 + r >= rows) break;
            const unsigned row = (unsigned)(r0 + r);
            const unsigned b = row / (unsigned)a.total_frames, f = row - b * (unsigned)a.total_frames;
            float m = -INFINITY;
#pragma unroll
            for (int i = 0; i < NW; ++i) m = fmaxf(m, v[r][i]);
            m = warp_max(m);
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < NW; ++i)
                if (lane + 32 * i < a.C) s += expf(v[r][i] - m);
            s = warp_sum(s);
            const float ls = logf(s);
            float* dst = a.out + (long long)b * a.out_pitch_b + (long long)f * a.C;
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                const int c = lane + 32 * i;
                if (c < a.C) dst[c] = (v[r][i] - m) - ls;
            }
        }
    }
}

// NW: 32-class words compiled in (C <= 32 NW), as in log_softmax_rows_kernel.
template <int NW>
__global__ void __launch_bounds__(STITCH_WARPS * 32) stitch_log_softmax_kernel(const __grid_constant__ StitchArgs a) {
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)a.B * a.total_frames;
    const long long nw = (long long)gridDim.x * STITCH_WARPS;
    for (long long row = (long long)blockIdx.x * STITCH_WARPS + (threadIdx.x >> 5); row < rows; row += nw) {
        const int b = (int)(row / a.total_frames), f = (int)(row % a.total_frames);
        float v[NW];
        if (a.fpw == 0) {
            const float* src = a.in + (long long)b * a.in_pitch_b + (long long)f * a.C;
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                const int c = lane + 32 * i;
                v[i] = c < a.C ? src[c] : -INFINITY;
            }
        } else {
            // windows i with i*sf <= f < i*sf + fpw, ascending (at most 3: fpw = 2 sf or 2 sf + 1)
            const int i1 = f / a.sf;
            float acc[NW];
#pragma unroll
            for (int i = 0; i < NW; ++i) acc[i] = 0.0f;
            float ws = 0.0f;
#pragma unroll
            for (int d = 2; d >= 0; --d) {
                const int wi = i1 - d;
                const int off = f - wi * a.sf;
                if (wi < 0 || wi >= a.W || off >= a.fpw) continue;
                const float w = a.weights[off];
                const float* src = a.in + (long long)b * a.in_pitch_b + ((long long)wi * a.fpw + off) * a.C;
#pragma unroll
                for (int i = 0; i < NW; ++i) {
                    const int c = lane + 32 * i;
                    if (c < a.C) acc[i] = __fadd_rn(acc[i], __fmul_rn(src[c], w));      // combined += logits * w  (:152 / :166)
                }
                ws = __fadd_rn(ws, w);                                                   // weight_sum += w  (:153 / :167)
            }
            const float den = __fadd_rn(ws, 1e-8f);                                      // :169
#pragma unroll
            for (int i = 0; i < NW; ++i) v[i] = (lane + 32 * i < a.C) ? __fdiv_rn(acc[i], den) : -INFINITY;
        }
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < NW; ++i) m = fmaxf(m, v[i]);
        m = warp_max(m);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NW; ++i)
            if (lane + 32 * i < a.C) s += expf(v[i] - m);
        s = warp_sum(s);
        const float ls = logf(s);
        float* dst = a.out + (long long)b * a.out_pitch_b + (long long)f * a.C;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            const int c = lane + 32 * i;
            if (c < a.C) dst[c] = (v[i] - m) - ls;
        }
    }
}

// _calculate_alignment_score (:767-773): one warp per utterance, frames dealt to the lanes, the reference's Python-float
// (double) accumulation; labels >= C (never produced by the decoder) are skipped like the reference's `if`.
__global__ void alignment_score_kernel(int B, int C, const float* __restrict__ logp, const long long* row_off, const int32_t* T,
                                       const int32_t* __restrict__ frame_ph, const long long* frame_off, double* score) {
    const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (u >= B) return;
    const float* lp = logp + row_off[u];
    const int32_t* ph = frame_ph + frame_off[u];
    const int Tu = T[u];
    double s = 0.0;
    for (int t = lane; t < Tu; t += 32) {
        const int c = ph[t];
        if (c >= 0 && c < C) s += (double)lp[(long long)t * C + c];
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(FULL, s, d);
    if (lane == 0) score[u] = s;
}

}  // namespace bfa
