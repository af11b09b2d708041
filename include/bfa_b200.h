/*
 * bfa_b200.h -- C ABI of libbfa_b200.so, the B200 (sm_100a) forced-alignment engine.
 *
 * Drop-in boundary for ONE path of tabahi/bournemouth-forced-aligner: the batched
 * monotonic Viterbi aligner + per-phoneme confidence pass
 *   bournemouth_aligner/forced_alignment.py   AlignmentUtils / ViterbiDecoder
 *   bournemouth_aligner/utils.py:70-113       _calculate_confidences
 * called from core.py:902-922 (decode_alignments), :1028 (decode_alignments_simple) and
 * :936-937 (_calculate_confidences).  The reference has no FFI layer for this path (it is
 * plain Python over torch); these entry points are what a ctypes binding in core.py would
 * call instead (INTEGRATION.md shows that binding).
 *
 * Conventions
 *  - plain pointers and sizes, no torch types.  Pointers marked [dev] are device pointers,
 *    [host] host pointers.  All buffers are caller-owned; inputs are never modified (the
 *    reference clones before mutating, forced_alignment.py:121).
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no
 *    hidden synchronisation and allocates nothing, except the *_host convenience entry.
 *    The device entries can therefore be captured into a CUDA graph (stream capture) and replayed on
 *    new data in the same buffers: the internal side streams fork and join with events only, and the
 *    kernels that use programmatic dependent launch are captured as such
 *    (tests/test_cuda_parity.py::test_align_batch_replays_from_a_cuda_graph).
 *  - return value: 0 on success, negative BFA_E_* on argument/launch errors.  Data conditions
 *    are reported per utterance in status[] (BFA_ST_*), mirroring the reference's exceptions
 *    (ValueError "Audio too short to align", forced_alignment.py:161-165).
 *  - log-posteriors are fp32, row-major [T_u, C] per utterance at logp + row_off[u]
 *    (elements).  Dense [B, T_max, C] is the special case row_off[u] = u*T_max*C.
 */
#ifndef BFA_B200_H
#define BFA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFA_VERSION 100 /* 0.1.0 */

/* return codes */
#define BFA_OK 0
#define BFA_E_INVALID (-1)     /* null pointer / bad size / C <= blank_id */
#define BFA_E_UNSUPPORTED (-2) /* shape outside compiled limits (C > 256); paths longer than BFA_MAX_L are a per-utterance status */
#define BFA_E_WORKSPACE (-3)   /* workspace too small: call bfa_workspace_bytes */
#define BFA_E_CUDA (-4)        /* CUDA runtime error, see bfa_last_cuda_error */

/* per-utterance status (low 3 bits = branch taken, bit 3 = degenerate score) */
#define BFA_ST_OK 0            /* single Viterbi over the utterance (forced_alignment.py:153-193) */
#define BFA_ST_EMPTY_TARGET 1  /* N == 0 (:894-897 / :112-118) */
#define BFA_ST_TOO_SHORT 2     /* T < N: the reference raises ValueError (:161-165) */
#define BFA_ST_PROPORTIONAL 3  /* T == N proportional assignment (:170-176) */
#define BFA_ST_SEGMENTED 4     /* silence-anchored segmentation accepted (:133-145) */
#define BFA_ST_DEFERRED 5      /* only with BFA_FLAG_DIRECT_ONLY: the utterance needs the full planner chain and was NOT aligned;
                                  run it again without that flag (the Python facade does) */
#define BFA_ST_UNSUPPORTED 6   /* the utterance's CTC path has more than BFA_MAX_L states (more than 2047 phonemes in one DP problem):
                                  it was NOT aligned (frames = blank, no stamps); every other utterance of the batch is unaffected */
#define BFA_ST_DEGENERATE 8    /* flag: winning DP score <= -1000 (the reference's "-inf") */
#define BFA_ST_STAMP_OVERFLOW 16 /* flag: more than max_stamps runs; stamps truncated */

#define BFA_MODE_FULL 0   /* AlignmentUtils.decode_alignments        (:856-910) */
#define BFA_MODE_SIMPLE 1 /* AlignmentUtils.decode_alignments_simple (:932-986) */

#define BFA_MAX_C 256
#define BFA_MAX_L 8192  /* CTC-path states per DP problem (N <= 2047 at stride 4); more than 1024 states: one CTA per problem */

/* Constants of ViterbiDecoder/AlignmentUtils (forced_alignment.py:16-23, :29, :226, :269, :418,
 * :777, :841).  Replaces the reference's constructor arguments + hard-coded literals. */
typedef struct BfaParams {
    int32_t blank_id;
    int32_t silence_id;       /* < 0 : None */
    int32_t silence_anchors;  /* 0 disables silence anchoring */
    int32_t ignore_noise;
    int32_t truly_forced;
    int32_t boost_targets;
    int32_t enforce_minimum;
    int32_t max_blanks;       /* 10 */
    float boost_factor;       /* 5.0 */
    float min_log_prob;       /* float32 log(1e-8) = -18.420681 */
    float neg_inf;            /* -1000.0 */
    float sub_boost;          /* 5.0 */
    int32_t boundary_pad;     /* 3 */
    int32_t min_speech_frames;/* 20 */
    int32_t mode;             /* BFA_MODE_* */
    int32_t reserved;         /* BFA_HINT_* / BFA_FLAG_* bits, 0 by default */
} BfaParams;

/* bits of BfaParams.reserved */
#define BFA_FLAG_EXACT_ONLY 1 /* run every item through the exact generic kernel (disables the banded fast path) */
#define BFA_HINT_NO_SIL 2     /* caller asserts that no target contains silence_id: skips the row-statistics pass that
                                 only the silence scan needs.  A pure performance hint: if it is wrong the planner
                                 recomputes what it needs (slower), results are unchanged. */
#define BFA_FLAG_UNFUSED_CONF 4  /* confidences gather lp[f, phoneme] from logp inside the stamp kernel instead of taking the
                                 * per-frame values the Viterbi back-trace collects (measurement / A-B switch) */
#define BFA_FLAG_FILL_ONLY 16     /* MEASUREMENT ONLY: the banded kernel stops after the DP fill (rows streamed, log-sum-exp, forward
                                 * recursion, decision records written) and skips the back-trace: outputs are NOT produced.  bench.py
                                 * uses it to time the fill phase by itself; never set it in production. */
#define BFA_FLAG_NO_SPEC 8       /* the banded kernel fetches every confidence input during its back-trace instead of keeping the
                                 * frame-wise best class's value while the row is on chip (measurement / A-B switch) */

#define BFA_FLAG_NO_DIRECT 32    /* never use the direct kernel, every utterance goes through planner + item lists: worth setting when
                                    (nearly) every target holds silence_id, i.e. the direct kernel would hand everything back; results are the same */
#define BFA_FLAG_DIRECT_ONLY 64  /* Launch ONLY the direct kernel (one kernel per call: in-kernel planning, banded stride-4 Viterbi,
                                  * frame labels, timestamps, confidences).  Utterances it cannot finish -- silence_id in the target
                                  * while silence anchoring is on, more phonemes than 4N+1 <= T allows, T == N, empty targets, paths
                                  * that touch the band edge, ... -- are reported as status BFA_ST_DEFERRED with n_stamps = 0 and must
                                  * be run again without this flag.  Nothing is ever silently different: an utterance is either
                                  * finished exactly as without the flag or flagged.  Ignored (full chain) when the direct kernel
                                  * cannot be used at all (ignore_noise == 0, C > 72, very long utterances). */
#define BFA_FLAG_PIPELINED 128   /* With BFA_FLAG_DIRECT_ONLY: the caller asserts that every INPUT of this call (logp, offsets,
                                  * lengths, targets) was complete before the operation that precedes this call in the stream was
                                  * enqueued (e.g. the posteriors are resident, or their producer is older than the previous call).
                                  * The kernel is then launched with programmatic stream serialization and starts reading its inputs
                                  * while its predecessor in the stream is still draining, SM by SM; it waits for the predecessor's
                                  * completion before writing any output.  Back-to-back calls overlap their ramp-up and tail. */

/* framestamp tuple (phoneme_id, start_frame, end_frame_exclusive, target_seq_idx)
 * = the 4-tuples returned by ViterbiDecoder.assort_frames (forced_alignment.py:777-834). */
typedef struct BfaStamp {
    int32_t phoneme;
    int32_t start;
    int32_t end;
    int32_t target_idx;
} BfaStamp;

/* Shape summary the host must provide (it owns pred_lens / true_seqs_lens anyway). */
typedef struct BfaShape {
    int32_t B;            /* utterances */
    int32_t C;            /* classes (66 / 67 / 17 ...) */
    int32_t max_T;        /* max frames of any utterance */
    int32_t max_N;        /* max target length of any utterance */
    int64_t total_frames; /* sum of T_u */
    int32_t max_stamps;   /* row pitch of stamps/conf (>= max_N, or >= max_T if !ignore_noise) */
    int32_t reserved;     /* hint, 0 = none: how many utterances the caller expects to need the exact kernel (too dense for the
                             stride-4 path, more than 128 phonemes, ...).  When > 0 the exact kernel's first pass runs next to
                             the banded kernel on a few SMs of its own.  Results never depend on it. */
} BfaShape;

int bfa_version(void);
const char *bfa_strerror(int code);
const char *bfa_last_cuda_error(void);
int bfa_sizeof_params(void);

/* == ViterbiDecoder.__init__/AlignmentUtils.__init__ defaults (forced_alignment.py:16-23, :841-853) */
void bfa_default_params(BfaParams *p, int32_t blank_id, int32_t silence_id);

/* Bytes of device scratch bfa_align_batch needs for this shape on the current device. */
size_t bfa_workspace_bytes(const BfaParams *p, const BfaShape *shape);

/*
 * == AlignmentUtils.decode_alignments(forced_alignment=True) (forced_alignment.py:856-910)
 *    or decode_alignments_simple (:932-986) when p->mode == BFA_MODE_SIMPLE,
 *    followed (if conf != NULL) by utils._calculate_confidences (utils.py:70-113) on the ORIGINAL
 *    log-posteriors, exactly as core.py:935-937 does.
 *
 *  logp      [dev] fp32 log-posteriors, utterance u at logp + row_off[u], T[u] rows of C
 *  row_off   [dev] int64[B]      T [dev] int32[B]
 *  tgt       [dev] int32 flat target ids; utterance u = tgt[tgt_off[u] .. tgt_off[u+1])
 *  tgt_off   [dev] int64[B+1]
 *  frame_ph/frame_idx [dev] int32[total_frames]: per-frame phoneme / target index (-1 on blanks),
 *            utterance u at frame_off[u]   (the tensors decode_with_forced_alignment returns)
 *  frame_off [dev] int64[B+1]
 *  dp_final  [dev] float[B] or NULL: winning DP score (unsegmented utterances; 0 otherwise)
 *  status    [dev] int32[B]  BFA_ST_*
 *  stamps    [dev] BfaStamp[B * max_stamps], n_stamps [dev] int32[B]
 *  conf      [dev] float[B * max_stamps] or NULL
 */
int bfa_align_batch(const BfaParams *p, const BfaShape *shape,
                    const float *logp, const int64_t *row_off, const int32_t *T,
                    const int32_t *tgt, const int64_t *tgt_off,
                    int32_t *frame_ph, int32_t *frame_idx, const int64_t *frame_off,
                    float *dp_final, int32_t *status,
                    BfaStamp *stamps, float *conf, int32_t *n_stamps,
                    void *workspace, size_t workspace_bytes, void *stream);

/*
 * == ViterbiDecoder._viterbi_decode (forced_alignment.py:563-703) on explicit CTC paths, one DP
 *    problem per item (white-box entry used by parity tests and by callers that build their own
 *    paths).  Item i: rows logp + row_off[i] (T[i] x C), path/true_idx at path_off[i] (L[i] states),
 *    band[i]; outputs at frame_off[i].  final_state [dev] int32[n] or NULL.
 */
int bfa_viterbi_paths(const BfaParams *p, int32_t n_items, int32_t C, int32_t max_T, int32_t max_L,
                      const float *logp, const int64_t *row_off, const int32_t *T,
                      const int32_t *path, const int32_t *true_idx, const int64_t *path_off,
                      const int32_t *L, const int32_t *band,
                      int32_t *frame_ph, int32_t *frame_idx, const int64_t *frame_off,
                      float *dp_final, int32_t *final_state,
                      void *workspace, size_t workspace_bytes, void *stream);
size_t bfa_viterbi_paths_workspace_bytes(int32_t n_items, int32_t max_T, int32_t max_L);

/*
 * == utils._calculate_confidences (utils.py:70-113) for a batch: utterance u has n_stamps[u]
 *    stamps at stamps + u*max_stamps; T_conf[u] is the row count used for clamping (core.py:936
 *    passes the un-sliced [T_max, C] matrix).
 */
int bfa_confidence_batch(int32_t B, int32_t C, const float *logp, const int64_t *row_off,
                         const int32_t *T_conf, const BfaStamp *stamps, const int32_t *n_stamps,
                         int32_t max_stamps, float *conf, void *stream);

/*
 * == ViterbiDecoder._calculate_alignment_score (forced_alignment.py:767-773) for a batch:
 *    score[u] = sum over t < T[u] of logp[row_off[u] + t*C + frame_ph[frame_off[u] + t]] (labels >= C skipped),
 *    accumulated in double like the reference's Python float.
 */
int bfa_alignment_score_batch(int32_t B, int32_t C, const float *logp, const int64_t *row_off,
                              const int32_t *T, const int32_t *frame_ph, const int64_t *frame_off,
                              double *score, void *stream);

/*
 * == the step in front of the aligner: stich_window_predictions (cupe2i/windowing.py:103-173) followed by
 *    F.log_softmax(.., dim=2) (core.py:898-899), in one pass.  window_logits is [B, n_windows, frames_per_window, C]
 *    (utterance b at window_logits + b*in_pitch), window_weights the reference's cross-fade window
 *    cos(linspace(-pi/2, pi/2, frames_per_window)) (:130, passed in so that it carries the caller's bits), logp_out
 *    [B, total_frames, C] (utterance b at logp_out + b*out_pitch).  The cross-fade follows the reference's arithmetic
 *    operation by operation; the result agrees with torch's stitch + log_softmax to 1e-5.
 *    frames_per_window == 0: no stitching, window_logits holds [B, total_frames, C] logits (plain log-softmax).
 */
int bfa_stitch_log_softmax(int32_t B, int32_t n_windows, int32_t frames_per_window, int32_t C,
                           int32_t total_frames, const float *window_logits, int64_t in_pitch,
                           const float *window_weights, float *logp_out, int64_t out_pitch, void *stream);

/*
 * == F.log_softmax (core.py:898-899) + decode_alignments + _calculate_confidences without the log-softmax pass: the same call
 *    as bfa_align_batch on rows that hold the acoustic model's UN-NORMALISED logits.
 *    Boosting re-normalises every row (forced_alignment.py:51-54), so emissions, path, timestamps and DP score do not
 *    depend on a per-row shift; the confidences (utils.py:81: exp of the ORIGINAL log-probabilities) do, and the
 *    kernel's row reduction carries the row's own log-sum-exp along for them.
 *    Requires p->mode == BFA_MODE_FULL and p->boost_targets (BFA_E_UNSUPPORTED otherwise).
 *  row_lse   [dev] float[total_frames] out: log(sum(exp(row))) of every frame of every utterance finished by this call
 *            (utterance u at frame_off[u]); log-probabilities downstream are logits[f, c] - row_lse[f].
 *    The planner chain follows as in bfa_align_batch whenever its silence pass runs (silence anchoring on, silence_id < C, no
 *    BFA_HINT_NO_SIL): that pass reads every row of the chain's utterances anyway and leaves their row_lse behind, the stamp
 *    kernel subtracts it where it exponentiates.  Otherwise (and under BFA_FLAG_DIRECT_ONLY) only the one-kernel pass runs and
 *    utterances it cannot finish come back with status BFA_ST_DEFERRED: normalise the rows (bfa_stitch_log_softmax with
 *    frames_per_window = 0) and call bfa_align_batch -- the Python facade does.
 */
int bfa_align_batch_logits(const BfaParams *p, const BfaShape *shape,
                           const float *logits, const int64_t *row_off, const int32_t *T,
                           const int32_t *tgt, const int64_t *tgt_off,
                           int32_t *frame_ph, int32_t *frame_idx, const int64_t *frame_off,
                           float *dp_final, int32_t *status,
                           BfaStamp *stamps, float *conf, int32_t *n_stamps, float *row_lse,
                           void *workspace, size_t workspace_bytes, void *stream);

/*
 * Host-buffer convenience entry (what a CPU-side caller binds): same semantics as bfa_align_batch
 * with every pointer a HOST pointer.  Copies inputs host->device in chunks of utterances on two
 * streams (copy of chunk i+1 overlaps the kernels of chunk i), runs the device pipeline, copies
 * results back, and synchronises before returning.  Device memory is taken from an internal
 * grow-only arena (released by bfa_host_release).  device = CUDA device ordinal.
 */
int bfa_align_batch_host(const BfaParams *p, const BfaShape *shape,
                         const float *logp, const int64_t *row_off, const int32_t *T,
                         const int32_t *tgt, const int64_t *tgt_off,
                         int32_t *frame_ph, int32_t *frame_idx, const int64_t *frame_off,
                         float *dp_final, int32_t *status,
                         BfaStamp *stamps, float *conf, int32_t *n_stamps,
                         int32_t device, int32_t chunk_utts);
void bfa_host_release(void);
const char *bfa_host_last_error(void);

/*
 * == ViterbiDecoder.assort_frames (forced_alignment.py:777-834) for a batch of per-frame
 *    (phoneme, target index) arrays.  status [dev] int32[B] is in/out: utterances whose low bits
 *    are EMPTY_TARGET / TOO_SHORT produce no stamps; BFA_ST_STAMP_OVERFLOW is or-ed in when a
 *    row needs more than max_stamps entries.
 */
int bfa_assort_batch(const BfaParams *p, int32_t B, const int32_t *T, const int64_t *frame_off,
                     const int32_t *frame_ph, const int32_t *frame_idx, int32_t *status,
                     BfaStamp *stamps, int32_t *n_stamps, int32_t max_stamps, void *stream);

/* == PhonemeTimestampAligner.extend_soft_boundaries_func (core.py:682-809), the step the reference runs between
 * decode_alignments and _calculate_confidences (core.py:925-937): stretches the start / end of every stamp in place over
 * neighbouring frames whose probability of the stamp's phoneme stays above the reference's thresholds (10^-3 scaled by the
 * stamp's mean probability, then 10^-boundary_softness).  stamps [B, max_stamps] with n_stamps[u] valid rows each (the
 * output layout of bfa_align_batch), T[u] frames per utterance.  max_stamps <= 6400. */
int bfa_soft_boundaries_batch(int32_t B, int32_t C, const float *logp, const int64_t *row_off, const int32_t *T, BfaStamp *stamps,
                              const int32_t *n_stamps, int32_t max_stamps, int32_t boundary_softness, void *stream);

/* The two steps after bfa_align_batch_logits on the same un-normalised rows: row_lse is that call's output and lse_off[u] the
 * offset of utterance u inside it (its frame_off); log-probabilities are formed on the fly as logits[f, c] - row_lse[f].
 * row_lse == NULL (and lse_off == NULL): exactly bfa_soft_boundaries_batch / bfa_confidence_batch. */
int bfa_soft_boundaries_batch_lse(int32_t B, int32_t C, const float *logits, const int64_t *row_off, const int32_t *T, BfaStamp *stamps,
                                  const int32_t *n_stamps, int32_t max_stamps, int32_t boundary_softness,
                                  const float *row_lse, const int64_t *lse_off, void *stream);
int bfa_confidence_batch_lse(int32_t B, int32_t C, const float *logits, const int64_t *row_off, const int32_t *T_conf,
                             const BfaStamp *stamps, const int32_t *n_stamps, int32_t max_stamps, float *conf,
                             const float *row_lse, const int64_t *lse_off, void *stream);

/* Measurement hook: when enabled, bfa_align_batch / bfa_viterbi_paths bracket the dominant kernel
 * (the Viterbi fill+back-trace) with CUDA events on the launch stream; bfa_profile_read waits for
 * them and returns the summed device time and the number of launches since the previous read.
 * on: 0 off, 1 on; (n << 8) | 1 brackets only every n-th bfa_align_batch call (n <= 255), starting
 * with the next one: the event records cost device time and stand between kernels that otherwise use
 * programmatic dependent launch, so a sampled measurement disturbs the timed loop less. */
void bfa_profile_enable(int on);
int bfa_profile_read(float *dominant_ms, int32_t *n_launches);
int bfa_profile_read_aux(float *out2);   /* development: planner / stamp kernel means after bfa_profile_enable(2) */

/* Development aid: sums of warp-clock cycles per phase of the banded Viterbi kernel (16 counters); all zero unless
 * the library was built with -DBFA_PHASE_PROF (scripts/phase_prof.sh).  reset != 0 clears them after the read. */
int bfa_debug_phases(unsigned long long *out32, int reset);
int bfa_debug_warps(unsigned long long *out32, int reset);
int bfa_debug_item_counts(int32_t *out4);
int bfa_debug_ctas(unsigned long long *out320, int reset);
int bfa_debug_fin(unsigned long long *out32, int reset);

/* Number of kernels this library has launched since load (bench.py's gpu_launches claim). */
int64_t bfa_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BFA_B200_H */
