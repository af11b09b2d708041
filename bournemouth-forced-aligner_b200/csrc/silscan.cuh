// silscan.cuh -- the input of the planner's silence scan, streamed at HBM speed.
//
// _detect_silence_segments (forced_alignment.py:471-541) thresholds a k-frame moving average of
// exp(modified_log_probs[t, silence_id]), where modified_log_probs is the boosted, re-normalised and floored copy of the
// posteriors (:121-129).  The probability of one frame needs the log-sum-exp of its whole boosted row, i.e. one full read
// of every utterance whose target holds silence_id before its segments -- and with them the DP problems -- are known.
// This kernel is that read: rows come in by 1-D bulk copies (TMA engine), 32 rows of one utterance per copy, each lane
// reduces ONE row (sum of 2^(x*log2e + w[class]) over the classes, no shuffles), and the warp turns the 32 silence
// probabilities into running sums (fp64, like torch.cumsum on CPU which accumulates float in double, :508).
//
// Work unit = (utterance, chunk of SS_CHUNK rows); D[frame] = inclusive sum of the silence probabilities from the first
// row of the frame's chunk; the planner adds the chunk bases (plan.cuh: chunk_bases) and forms every window sum the
// reference forms as a difference of two prefix values.  silunits_kernel lists the units (only utterances whose target
// holds silence_id have any); they are dealt round-robin to the resident warps.
#pragma once
#include "bfa_common.cuh"
#include "viterbi_band3.cuh"

namespace bfa {

constexpr int SS_WARPS = 8;
constexpr int SS_ROWS = 32;        // rows per bulk copy = one row per lane
constexpr int SS_CHUNK = 256;      // rows per work unit
constexpr int SS_CHUNK_SHIFT = 8;
constexpr int SS_MAXST = 6;        // bulk copies in flight per warp, at most
constexpr int SS_PAD = 8;          // floats of slack per stage (lead-in of a misaligned utterance)

struct SilArgs {
    BfaParams p;
    int B, C, nst;
    int keep_in_l2;            // the batch fits in L2: leave the rows there for the banded kernel's read (no evict-first hint)
    const float* logp;
    const long long* row_off;
    const int32_t* T;
    const int32_t* tgt;
    const long long* tgt_off;
    const long long* frame_off;
    double* D;                 // [total_frames]
    uint32_t* tmask;           // [B][MAX_WORDS] target-class masks of the listed utterances (silunits_kernel -> silprob_kernel)
    int2* units;               // (utterance, chunk) work units of the utterances whose target holds silence_id
    int* n_units;
    const int* deferred;       // when non-null: only the utterances deferred[0 .. *n_deferred) (what the direct kernel left)
    const int* n_deferred;
    // logits in (bfa_align_batch_logits, planner chain): this pass reads every row of every utterance the chain works on anyway
    // (list_all: also those without silence_id in the target) and leaves the rows' plain log-sum-exp behind for the stamp kernel
    float* row_lse;            // [total_frames] or null
    int list_all;
};

__host__ __device__ inline size_t ss_stage_bytes(int C) { return C <= B3_KK ? ((size_t)SS_ROWS * C + SS_PAD) * 4 : 0; }
__host__ __device__ inline size_t ss_smem_per_warp(int C, int nst) {
    return ((size_t)nst * ss_stage_bytes(C) + (size_t)B3_KK * 4 + MAX_WORDS * 4 + SS_MAXST * 8 + 127) / 128 * 128;
}
// stages per warp that fit (0: the staged path is not available, every unit takes the gather path)
inline int ss_stages(int C, size_t smem_max) {
    if (C > B3_KK) return 0;
    int n = SS_MAXST;
    while (n >= 2 && ss_smem_per_warp(C, n) * SS_WARPS > smem_max) --n;
    return n >= 2 ? n : 0;
}

// exp(modified_lp[row, silence_id]) of one row the slow, exact way (max-subtracted, expf / logf): rows whose fast sum is
// not a positive finite number.  `row(c)` returns the raw value of class c, `w(c)` the class weight (0 for target classes).
template <typename Row, typename Wt>
__device__ __noinline__ float ss_exact_prob(Row row, Wt is_target, int C, int sil, float boost) {
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, row(c) + (is_target(c) ? boost : 0.0f));
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(row(c) + (is_target(c) ? boost : 0.0f) - m);
    return expf(((row(sil) + (is_target(sil) ? boost : 0.0f)) - m) - logf(s));
}

// Gather path of one chunk (rows_u <= SS_CHUNK rows at `base`): no boost (the probability is the raw one: one strided read per
// row), class counts beyond the staged path's table, an utterance whose lead-in would start before the caller's buffer, or
// the planner itself when it was told (wrongly) that no target holds silence_id.  Same results, not HBM speed.
// mw = the utterance's target-class mask words.
__device__ __noinline__ void ss_gather_rows(const BfaParams& p, int C, const float* base, int rows_u, double* Dout, int lane,
                                            const uint32_t* mw, bool sil_tgt, float* lse_out = nullptr) {
    const bool boost = p.boost_targets != 0;
    const int sil = p.silence_id;
    uint32_t tbits = 0;
#pragma unroll
    for (int i = 0; i < MAX_WORDS; ++i) tbits |= ((mw[i] >> lane) & 1u) << i;
    const int nblk = (rows_u + SS_ROWS - 1) / SS_ROWS;
    double carry = 0.0;
    for (int b = 0; b < nblk; ++b) {
        const int t = b * SS_ROWS + lane;
        float m = 0.f, ls = 0.f;
        if (boost) {
            const int rows = min(SS_ROWS, rows_u - b * SS_ROWS);
            for (int r = 0; r < rows; ++r) {
                const float* row = base + (long long)(b * SS_ROWS + r) * C;
                const float2 s = row_stats_warp([&](int c) { return row[c]; }, C, lane, tbits, p.boost_factor);
                if (lane == r) { m = s.x; ls = s.y; }
                if (lse_out) {                                                // logits in: the plain log-sum-exp of the row
                    const float2 s0 = row_stats_warp([&](int c) { return row[c]; }, C, lane, 0u, 0.0f);
                    if (lane == r) lse_out[b * SS_ROWS + r] = s0.x + s0.y;
                }
            }
        }
        float pr = 0.f;
        if (t < rows_u) {
            const float x = base[(long long)t * C + sil];
            pr = expf(mod_value(x, sil_tgt, boost, p.enforce_minimum != 0, p.boost_factor, m, ls, p.min_log_prob));
        }
        double v = (double)pr;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double o = __shfl_up_sync(FULL, v, d);
            if (lane >= d) v += o;
        }
        v += carry;
        if (t < rows_u) Dout[t] = v;
        carry = __shfl_sync(FULL, v, 31);
    }
}

// Which utterances need the pass: one warp per utterance looks for silence_id in the target (a segmentation attempt is made
// exactly then, :293-295), leaves the target-class mask behind and appends the utterance's chunks to the unit list.
__global__ void __launch_bounds__(256) silunits_kernel(const __grid_constant__ SilArgs a) {
    __shared__ uint32_t s_mask[8][MAX_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_wait();
    const int nU = a.deferred ? *a.n_deferred : a.B;
    const int ui = blockIdx.x * 8 + warp;
    if (ui >= nU) return;
    const int u = a.deferred ? a.deferred[ui] : ui;
    const int Tu = a.T[u];
    const TgtInfo ti = target_info(a.tgt, a.tgt_off[u], a.tgt_off[u + 1], a.C, a.p.blank_id, a.p.silence_id, lane, s_mask[warp]);
    if ((!ti.has_sil && !a.list_all) || Tu <= 0) return;
    if (lane < MAX_WORDS) a.tmask[(size_t)u * MAX_WORDS + lane] = s_mask[warp][lane];
    const int n = (Tu + SS_CHUNK - 1) >> SS_CHUNK_SHIFT;
    int base = 0;
    if (lane == 0) base = atomicAdd(a.n_units, n);
    base = __shfl_sync(FULL, base, 0);
    for (int i = lane; i < n; i += 32) a.units[base + i] = make_int2(u, i);
}

// LG: logits in -- the reduction also carries the plain sum (2^(-kk) = 1 + kk * MSC, see viterbi_band3.cuh) and row_lse is written.
template <bool LG>
__global__ void __launch_bounds__(SS_WARPS * 32, 1) silprob_kernel(const __grid_constant__ SilArgs a) {
    constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    extern __shared__ __align__(128) unsigned char ss_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = a.C, nst = a.nst;
    const BfaParams& p = a.p;
    const int sil = p.silence_id;
    const bool boost = p.boost_targets != 0;
    const size_t stage_bytes = ss_stage_bytes(C);
    const int stage_floats = (int)(stage_bytes >> 2);
    unsigned char* my = ss_smem + (size_t)warp * ss_smem_per_warp(C, nst);
    float* stage = reinterpret_cast<float*>(my);
    float* kk = reinterpret_cast<float*>(my + (size_t)nst * stage_bytes);     // class weights of the fused log-sum-exp
    uint32_t* mw = reinterpret_cast<uint32_t*>(kk + B3_KK);                   // target-class mask words
    const uint32_t bar0 = smem_u32(mw + MAX_WORDS);
    if (lane == 0) {
        for (int s = 0; s < nst; ++s) mbar_init(bar0 + 8u * s, 1);
        fence_mbar_init();
    }
    __syncwarp();
    pdl_wait();                // the deferred list (and its length) come from the kernel before this one
    if (sil < 0 || sil >= C) return;                                          // :497: nothing is ever detected
    const int total = *a.n_units;
    const int gw = blockIdx.x * SS_WARPS + warp, nw = gridDim.x * SS_WARPS;
    const uint64_t pol = policy_evict_first();
    const float min_p = expf(p.min_log_prob);
    const float MSC = (LG && p.boost_factor != 0.0f) ? (1.0f - expf(p.boost_factor)) / (p.boost_factor * LOG2E) : 0.0f;
    uint32_t phase = 0;

    for (int unit = gw; unit < total; unit += nw) {
        const int2 uc = a.units[unit];
        const int u = uc.x, r0 = uc.y * SS_CHUNK;
        const int Tu = a.T[u];
        __syncwarp();
        if (lane < MAX_WORDS) mw[lane] = a.tmask[(size_t)u * MAX_WORDS + lane];
        __syncwarp();
        const bool sil_tgt = (mw[sil >> 5] >> (sil & 31)) & 1u;
        const int rows_u = min(SS_CHUNK, Tu - r0);
        const int nblk = (rows_u + SS_ROWS - 1) / SS_ROWS;
        const long long ro = a.row_off[u];
        const float* base = a.logp + ro + (long long)r0 * C;
        double* Dout = a.D + a.frame_off[u] + r0;
        float* Lout = LG ? a.row_lse + a.frame_off[u] + r0 : nullptr;
        const int lead = (int)(((unsigned long long)(a.logp + ro) & 15ull) >> 2);     // 128*C bytes per copy: every copy of the utterance has this lead-in
        const bool staged = nst > 0 && boost && (lead == 0 || ro >= lead);
        double carry = 0.0;
        if (staged) {
            for (int c = lane; c < B3_KK; c += 32)
                kk[c] = c < C ? (((mw[c >> 5] >> (c & 31)) & 1u) ? 0.0f : -p.boost_factor * LOG2E) : -INFINITY;
            __syncwarp();
            const uint32_t lead_b = 4u * (uint32_t)lead;
            auto issue = [&](int b) {
                if (lane == 0) {
                    const int st = b % nst;
                    const int rows = min(SS_ROWS, rows_u - b * SS_ROWS);
                    const uint32_t bar = bar0 + 8u * st;
                    float* d = stage + (size_t)st * stage_floats;
                    const float* s = base + (long long)b * SS_ROWS * C - lead;
                    const uint32_t bytes = (uint32_t)rows * C * 4 + lead_b;
                    const bool last = r0 + b * SS_ROWS + rows >= Tu;          // never read past the utterance's last row
                    const uint32_t bulk = last ? (bytes & ~15u) : ((bytes + 15u) & ~15u);
                    for (uint32_t w = bulk >> 2; w < (bytes >> 2); ++w) d[w] = s[w];   // < 4 tail floats of the last copy
                    mbar_expect_tx(bar, bulk);
                    if (bulk) {
                        if (a.keep_in_l2) bulk_g2s(smem_u32(d), s, bulk, bar);
                        else bulk_g2s_hint(smem_u32(d), s, bulk, bar, pol);
                    }
                }
            };
            for (int b = 0; b < nst && b < nblk; ++b) issue(b);
            const bool vec2 = ((C | lead) & 1) == 0;                          // rows start on 8-byte boundaries
            for (int b = 0; b < nblk; ++b) {
                const int st = b % nst;
                mbar_wait(bar0 + 8u * st, (phase >> st) & 1u);
                phase ^= 1u << st;
                __syncwarp();
                const int t = b * SS_ROWS + lane;                             // row of the chunk
                float pr = 0.f;
                if (t < rows_u) {
                    const float* rowp = stage + (size_t)st * stage_floats + lead + lane * C;
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                    float z0 = 0.f, z1 = 0.f, z2 = 0.f, z3 = 0.f;         // LG: the plain sums
                    auto term = [&](float x, float kc, float& acc, float& zacc) {
                        const float e = b3_ex2(fmaf(x, LOG2E, kc));
                        acc += e;
                        if constexpr (LG) zacc = fmaf(e, kc, zacc);
                    };
                    if (vec2) {
                        const float2* x2 = reinterpret_cast<const float2*>(rowp);
                        const float4* k4 = reinterpret_cast<const float4*>(kk);
                        const int n4 = C >> 2;
#pragma unroll 4
                        for (int i = 0; i < n4; ++i) {
                            const float2 xa = x2[2 * i], xb = x2[2 * i + 1];
                            const float4 kv = k4[i];
                            term(xa.x, kv.x, s0, z0);
                            term(xa.y, kv.y, s1, z1);
                            term(xb.x, kv.z, s2, z2);
                            term(xb.y, kv.w, s3, z3);
                        }
                        if (C & 2) {
                            const float2 xa = x2[2 * n4];
                            term(xa.x, kk[4 * n4], s0, z0);
                            term(xa.y, kk[4 * n4 + 1], s1, z1);
                        }
                    } else {
                        int i = 0;
#pragma unroll 2
                        for (; i + 4 <= C; i += 4) {
                            term(rowp[i], kk[i], s0, z0);
                            term(rowp[i + 1], kk[i + 1], s1, z1);
                            term(rowp[i + 2], kk[i + 2], s2, z2);
                            term(rowp[i + 3], kk[i + 3], s3, z3);
                        }
                        for (; i < C; ++i) term(rowp[i], kk[i], s0, z0);
                    }
                    const float S = (s0 + s1) + (s2 + s3);
                    if (S > 0.f && S < 3.0e38f) pr = b3_ex2(fmaf(rowp[sil], LOG2E, kk[sil])) / S;
                    else pr = ss_exact_prob([&](int c) { return rowp[c]; }, [&](int c) { return kk[c] == 0.0f; }, C, sil, p.boost_factor);
                    if (p.enforce_minimum && sil_tgt) pr = fmaxf(pr, min_p);   // :75-81, exp is monotone
                    if constexpr (LG) {
                        const float Z = fmaf(MSC, (z0 + z1) + (z2 + z3), S);     // sum(e (1 + kk MSC)) = S + MSC sum(e kk)
                        float l0;
                        if (Z > 0.f && Z < 3.0e38f) l0 = b3_lg2(Z) * LN2;
                        else {                                               // the slow, exact way (max-subtracted)
                            float m = -INFINITY;
                            for (int c = 0; c < C; ++c) m = fmaxf(m, rowp[c]);
                            float sz = 0.f;
                            for (int c = 0; c < C; ++c) sz += expf(rowp[c] - m);
                            l0 = m + logf(sz);
                        }
                        Lout[t] = l0;
                    }
                }
                __syncwarp();                                                 // every lane has read its row: the stage may be refilled
                if (b + nst < nblk) issue(b + nst);
                double v = (double)pr;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const double o = __shfl_up_sync(FULL, v, d);
                    if (lane >= d) v += o;
                }
                v += carry;
                if (t < rows_u) Dout[t] = v;
                carry = __shfl_sync(FULL, v, 31);
            }
        } else {
            ss_gather_rows(p, C, base, rows_u, Dout, lane, mw, sil_tgt, Lout);
        }
    }
}

}  // namespace bfa
