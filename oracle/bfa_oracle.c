/*
 * bfa_oracle.c -- CPU restatement of the reference forced-alignment path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product (bournemouth-forced-aligner_b200/)
 * never links, imports or calls anything in oracle/.
 *
 * It restates, in plain scalar C, the algorithm of the reference's
 *   bournemouth_aligner/forced_alignment.py  (ViterbiDecoder, AlignmentUtils)
 *   bournemouth_aligner/utils.py:70-113      (_calculate_confidences)
 * Every function cites the reference lines it follows.  Parity is pinned by
 * tests/golden/ (vectors generated here by importing the unmodified reference
 * module, see tests/golden/make_golden.py) -- the reference's own test-suite
 * holds no vectors for this path.
 *
 * Numerics: the DP is single fp32 adds/compares (bit-reproducible).  The only
 * library numerics are expf/logf inside log-softmax / exp, where torch's
 * vectorised CPU kernels may differ by an ulp; tests state the tolerance.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_EMPTY_TARGET 1   /* N == 0 */
#define ORC_TOO_SHORT 2      /* T < N: reference raises ValueError (forced_alignment.py:161-165) */
#define ORC_PROPORTIONAL 3   /* T == N proportional assignment (:170-176) */
#define ORC_SEGMENTED 4      /* silence-anchored segmentation accepted (:133-145) */
#define ORC_DEGENERATE 8     /* flag bit: final dp <= -1000 */

typedef struct {
    int32_t blank_id;
    int32_t silence_id;       /* < 0 : None */
    int32_t silence_anchors;  /* AlignmentUtils default 10 */
    int32_t ignore_noise;
    int32_t truly_forced;
    int32_t boost_targets;
    int32_t enforce_minimum;
    int32_t max_blanks;       /* assort_frames default 10 */
    float boost_factor;       /* 5.0  (:29) */
    float min_log_prob;       /* float32 log(1e-8) (:70) */
    float neg_inf;            /* -1000.0 (:23) */
    float sub_boost;          /* 5.0 blank boost on sub-silences (:418) */
    int32_t boundary_pad;     /* 3  (:269) */
    int32_t min_speech_frames;/* 20 (:269) */
    int32_t mode;             /* 0 full (decode_alignments), 1 simple (decode_alignments_simple) */
    int32_t reserved;
} OrcParams;

typedef struct {
    int32_t phoneme, start, end, target_idx;
} OrcStamp;

/* ------------------------------------------------------------------------- */
/* forced_alignment.py:54, :560 -- F.log_softmax(x, dim=-1), one row.          */
static void row_log_softmax(float *x, int C)
{
    float m = x[0];
    for (int c = 1; c < C; ++c) if (x[c] > m) m = x[c];
    float s = 0.0f;
    for (int c = 0; c < C; ++c) s += expf(x[c] - m);
    float ls = logf(s);
    for (int c = 0; c < C; ++c) x[c] = (x[c] - m) - ls;
}

/* unique_targets = set(seq) - {blank, -100}, restricted to p < C (:44-49).  */
static void target_mask(const int32_t *seq, int N, int C, int blank_id, uint8_t *mask)
{
    memset(mask, 0, (size_t)C);
    for (int j = 0; j < N; ++j) {
        int p = seq[j];
        if (p == blank_id || p == -100) continue;
        if (p >= 0 && p < C) mask[p] = 1;
        /* negative ids other than -100 would wrap in torch; not a supported input */
    }
}

/* _boost_target_phonemes (:29-56) then _enforce_minimum_probabilities (:58-83) */
void orc_prep(float *lp, int T, int C, const int32_t *seq, int N, const OrcParams *p)
{
    uint8_t *mask = (uint8_t *)malloc((size_t)C);
    target_mask(seq, N, C, p->blank_id, mask);
    if (p->boost_targets) {
        for (int t = 0; t < T; ++t) {
            float *row = lp + (size_t)t * C;
            for (int c = 0; c < C; ++c) if (mask[c]) row[c] += p->boost_factor;
            row_log_softmax(row, C);
        }
    }
    if (p->enforce_minimum) {
        for (int t = 0; t < T; ++t) {
            float *row = lp + (size_t)t * C;
            for (int c = 0; c < C; ++c)
                if (mask[c] && row[c] < p->min_log_prob) row[c] = p->min_log_prob;
        }
    }
    free(mask);
}

/* ------------------------------------------------------------------------- */
/* _viterbi_decode (:563-703).  path/true_idx are explicit so that white-box
 * tests can feed arbitrary CTC paths.  dp_dump/bp_dump (may be NULL) receive
 * the full tables [T,L] for white-box comparison.                             */
int orc_viterbi(const float *lp, int T, int C, const int32_t *path, const int32_t *true_idx,
                int L, int band, int blank_id, int truly_forced, float neg_inf,
                int32_t *out_ph, int32_t *out_idx, float *dp_final, int32_t *final_state_out,
                float *dp_dump, int32_t *bp_dump)
{
    if (T <= 0 || L <= 0) return -1;
    const float NEG = neg_inf;
    float *prev = (float *)malloc(sizeof(float) * (size_t)L);
    float *cur = (float *)malloc(sizeof(float) * (size_t)L);
    int8_t *bp = (int8_t *)calloc((size_t)T * (size_t)L, 1);      /* k in {0,1,2}; row 0 = 0 */
    uint8_t *can_skip = (uint8_t *)calloc((size_t)L, 1);

    int use_band = (band > 0 && T > 1 && L > 1);                  /* :586 */
    double pace = use_band ? (double)(L - 1) / (double)(T - 1) : 0.0; /* :587 */

    for (int s = 0; s < L; ++s) prev[s] = NEG;                    /* :582 */
    prev[0] = lp[blank_id];                                       /* :594 */
    if (L > 1) prev[1] = lp[path[1]];                             /* :596 */
    for (int s = 2; s < L; ++s) can_skip[s] = (path[s] != path[s - 2]); /* :603-605 */
    if (dp_dump) memcpy(dp_dump, prev, sizeof(float) * (size_t)L);
    if (bp_dump) for (int s = 0; s < L; ++s) bp_dump[s] = 0;

    for (int t = 1; t < T; ++t) {                                 /* :608 */
        const float *row = lp + (size_t)t * C;
        int8_t *bpt = bp + (size_t)t * L;
        for (int s = 0; s < L; ++s) {
            float e = row[path[s]];
            float c0 = prev[s] + e;                               /* :613 */
            float c1 = (s >= 1) ? prev[s - 1] + e : NEG;          /* :616-617, :642 */
            float c2 = (s >= 2 && can_skip[s]) ? prev[s - 2] + e : NEG; /* :620-625, :642 */
            int k = 0; float best = c0;                           /* argmax: first max (:645) */
            if (c1 > best) { best = c1; k = 1; }
            if (c2 > best) { best = c2; k = 2; }
            cur[s] = best;
            bpt[s] = (int8_t)k;                                   /* backpointer = s - k (:647) */
        }
        if (bp_dump) for (int s = 0; s < L; ++s) bp_dump[(size_t)t * L + s] = s - bpt[s];
        if (use_band) {                                           /* :650-653 */
            double center = (double)t * pace;
            float lo = (float)(center - (double)band);
            float hi = (float)(center + (double)band);
            for (int s = 0; s < L; ++s)
                if ((float)s < lo || (float)s > hi) cur[s] = NEG;
        }
        if (dp_dump) memcpy(dp_dump + (size_t)t * L, cur, sizeof(float) * (size_t)L);
        float *tmp = prev; prev = cur; cur = tmp;
    }
    /* prev == dp[T-1] */
    int f;
    if (!truly_forced) {                                          /* :656-666 */
        f = -1; float best = 0.f;
        for (int s = 0; s < L; ++s)
            if (prev[s] > NEG && (f < 0 || prev[s] > best)) { f = s; best = prev[s]; }
        if (f < 0) {
            f = 0; best = prev[0];
            for (int s = 1; s < L; ++s) if (prev[s] > best) { best = prev[s]; f = s; }
        }
    } else {                                                      /* :668-682 */
        f = L - 1;
        if (prev[f] <= NEG && L >= 2) f = L - 2;
        if (prev[f] <= NEG) {
            int r = -1;
            for (int s = 0; s < L; ++s) if (prev[s] > NEG) r = s;
            f = (r >= 0) ? r : L - 1;
        }
    }
    if (dp_final) *dp_final = prev[f];
    if (final_state_out) *final_state_out = f;

    int ps = f;                                                   /* :686-692 */
    for (int t = T - 1; t >= 0; --t) {
        int idx = ps < 0 ? ps + L : ps;                           /* python negative-index wrap */
        out_ph[t] = path[idx];                                    /* :695 */
        out_idx[t] = true_idx ? true_idx[idx] : -1;               /* :700 */
        if (t > 0) ps = idx - bp[(size_t)t * L + idx];            /* bp[t][ps] = s - k, s = wrapped idx */
    }
    free(prev); free(cur); free(bp); free(can_skip);
    return 0;
}

/* Path construction shared by :181-187, :433-438, :970-973.                   */
static void build_path(const int32_t *seq, int N, int idx0, int stride, int blank_id,
                       int32_t *path, int32_t *true_idx)
{
    int L = stride * N + 1;
    for (int s = 0; s < L; ++s) { path[s] = blank_id; true_idx[s] = -1; }
    for (int j = 0; j < N; ++j) { path[1 + j * stride] = seq[j]; true_idx[1 + j * stride] = idx0 + j; }
}

/* ------------------------------------------------------------------------- */
/* _detect_silence_segments (:471-541).  segs: [2*max] ints; returns count.    */
int orc_detect_silence(const float *lp, int T, int C, int silence_id, double thr_d, int k,
                       int32_t *segs, int max_segs)
{
    if (silence_id >= C) return 0;                                /* :497 */
    if (T < k) return 0;                                          /* :499 */
    if (T <= 0) return 0;
    int nwin = (k > 1) ? T - k + 1 : T;
    float thr = (float)thr_d;                                     /* f32 tensor >= python scalar */
    float *padded = (float *)malloc(sizeof(float) * (size_t)(T + 1));
    float *p = (float *)malloc(sizeof(float) * (size_t)T);
    for (int t = 0; t < T; ++t) p[t] = expf(lp[(size_t)t * C + silence_id]); /* :503-504 */
    if (k > 1) {                                                  /* :507-510 */
        double acc = 0.0; padded[0] = 0.0f;                       /* CPU cumsum: double accumulate */
        for (int t = 0; t < T; ++t) { acc += (double)p[t]; padded[t + 1] = (float)acc; }
    }
    int n = 0, in_sil = 0, start = 0;
    for (int i = 0; i < nwin; ++i) {                              /* :524-533 */
        float avg = (k > 1) ? (padded[i + k] - padded[i]) / (float)k : p[i];
        int sil = avg >= thr;
        if (sil && !in_sil) { in_sil = 1; start = i; }
        else if (!sil && in_sil) {
            in_sil = 0;
            int e = i + k - 1; if (e > T) e = T;
            if (e - start >= k && n < max_segs) { segs[2 * n] = start; segs[2 * n + 1] = e; ++n; }
        }
    }
    if (in_sil) {                                                 /* :536-539 */
        if (T - start >= k && n < max_segs) { segs[2 * n] = start; segs[2 * n + 1] = T; ++n; }
    }
    free(padded); free(p);
    return n;
}

/* _anchor_silence_in_log_probs (:543-561) on a private copy of the segment.   */
static void anchor_silence(float *lp, int C, int blank_id, const int32_t *segs, int nseg, float boost)
{
    for (int i = 0; i < nseg; ++i)
        for (int t = segs[2 * i]; t < segs[2 * i + 1]; ++t) {
            float *row = lp + (size_t)t * C;
            row[blank_id] += boost;
            row_log_softmax(row, C);
        }
}

typedef struct { int a0, a1, t0, t1, sil; } Seg;

/* _segmented_viterbi_decode (:268-469).  Returns 1 when a full [T] result was
 * produced, 0 when the reference returns ([], []) (caller falls back).        */
static int segmented_decode(const float *lp, int T, int C, const int32_t *seq, int N,
                            const OrcParams *P, int32_t *out_ph, int32_t *out_idx, int *degenerate)
{
    const int sil_id = P->silence_id;
    /* Step 1: _find_target_sil_groups (:203-224) */
    int32_t *grp = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(N + 1));
    int ng = 0;
    for (int i = 0; i < N;) {
        if (seq[i] == sil_id) { int s = i; while (i < N && seq[i] == sil_id) ++i; grp[2 * ng] = s; grp[2 * ng + 1] = i; ++ng; }
        else ++i;
    }
    if (ng == 0) { free(grp); return 0; }                         /* :293-295 */
    int k = P->silence_anchors;                                   /* :296 */
    int max_s = T + 2;
    int32_t *asil = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)max_s);
    int na = orc_detect_silence(lp, T, C, sil_id, 0.9, k, asil, max_s);      /* :297 */
    if (na == 0 && N > 200) {                                     /* :298-304 */
        double nt = 1.0 - (0.09 * k); if (nt < 0.05) nt = 0.05;
        na = orc_detect_silence(lp, T, C, sil_id, nt, k, asil, max_s);
    }
    if (na == 0 && N > 200 && k > 3) {                            /* :305-308 */
        k = 3;
        na = orc_detect_silence(lp, T, C, sil_id, 0.9, k, asil, max_s);
    }
    if (na == 0) { free(grp); free(asil); return 0; }             /* :315-320 */

    /* Step 2: _match_silences (:226-266) */
    int32_t *mt = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)ng);
    int nm = 0, audio_idx = 0;
    for (int g = 0; g < ng; ++g) {
        double tpos = (double)(grp[2 * g] + grp[2 * g + 1]) / 2.0 / (double)N;
        int best = -1; double bd = INFINITY;
        for (int ai = audio_idx; ai < na; ++ai) {
            double apos = (double)(asil[2 * ai] + asil[2 * ai + 1]) / 2.0 / (double)T;
            double d = fabs(tpos - apos);
            if (d < bd) { bd = d; best = ai; }
            else if (d > bd) break;
        }
        if (best >= 0 && bd < 0.3) { mt[2 * nm] = g; mt[2 * nm + 1] = best; ++nm; audio_idx = best + 1; }
    }
    if (nm == 0) { free(grp); free(asil); free(mt); return 0; }   /* :324-325 */

    /* Step 3: segment list (:327-354) */
    Seg *segs = (Seg *)malloc(sizeof(Seg) * (size_t)(2 * nm + 2));
    int ns = 0, pa = 0, pt = 0;
    for (int m = 0; m < nm; ++m) {
        int tg0 = grp[2 * mt[2 * m]], tg1 = grp[2 * mt[2 * m] + 1];
        int as0 = asil[2 * mt[2 * m + 1]], as1 = asil[2 * mt[2 * m + 1] + 1];
        if (pa < as0 && pt < tg0) segs[ns++] = (Seg){pa, as0, pt, tg0, 0};
        else if (pa < as0) segs[ns++] = (Seg){pa, as0, pt, pt, 0};
        segs[ns++] = (Seg){as0, as1, tg0, tg1, 1};
        pa = as1; pt = tg1;
    }
    if (pa < T && pt < N) segs[ns++] = (Seg){pa, T, pt, N, 0};
    else if (pa < T) segs[ns++] = (Seg){pa, T, pt, pt, 0};
    /* Step 3b: merge short speech segments into the previous one (:356-369) */
    int nmg = 0;
    for (int i = 0; i < ns; ++i) {
        Seg s = segs[i];
        if (!s.sil && (s.t1 - s.t0) > 0 && (s.a1 - s.a0) < P->min_speech_frames && nmg > 0) {
            segs[nmg - 1].a1 = s.a1; segs[nmg - 1].t1 = s.t1; segs[nmg - 1].sil = 0;
        } else segs[nmg++] = s;
    }
    ns = nmg;

    /* Step 4 (:377-451).  Frames are appended in order; total may differ from T. */
    int cap = T + 8;
    for (int i = 0; i < ns; ++i) { int n = segs[i].a1 - segs[i].a0; if (n > 0) cap += n; }
    int32_t *ph = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
    int32_t *ix = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
    int w = 0, ok = 1;
    for (int i = 0; i < ns && ok; ++i) {
        Seg s = segs[i];
        int nf = s.a1 - s.a0;
        if (nf <= 0) continue;
        if (s.sil) {                                              /* :382-397 */
            int nsil = s.t1 - s.t0;
            for (int f = 0; f < nf; ++f) { ph[w + f] = sil_id; ix[w + f] = -1; }
            if (nsil > 0) {
                double fps = (double)nf / (double)nsil;
                for (int q = 0; q < nsil; ++q) {
                    int f0 = (int)((double)q * fps), f1 = (int)((double)(q + 1) * fps);
                    if (f1 > nf) f1 = nf;
                    for (int f = f0; f < f1; ++f) ix[w + f] = s.t0 + q;
                }
            }
            w += nf;
        } else {
            int p0 = s.a0 - P->boundary_pad; if (p0 < 0) p0 = 0;  /* :401-403 */
            int p1 = s.a1 + P->boundary_pad; if (p1 > T) p1 = T;
            int pad_left = s.a0 - p0;
            int n = s.t1 - s.t0, Ts = p1 - p0;
            if (n == 0) {                                         /* :409-412 */
                for (int f = 0; f < nf; ++f) { ph[w + f] = P->blank_id; ix[w + f] = -1; }
                w += nf;
                continue;
            }
            float *slp = (float *)malloc(sizeof(float) * (size_t)Ts * C);
            memcpy(slp, lp + (size_t)p0 * C, sizeof(float) * (size_t)Ts * C);
            int32_t *ss = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(Ts + 2));
            int nss = orc_detect_silence(slp, Ts, C, sil_id, 0.8, k, ss, Ts + 2);  /* :417 */
            anchor_silence(slp, C, P->blank_id, ss, nss, P->sub_boost);             /* :415-419 */
            free(ss);
            int stride = 4;                                       /* :423-426 */
            if ((double)(stride * n + 1) > (double)Ts * 0.9) stride = 3;
            if ((double)(stride * n + 1) > (double)Ts * 0.8) stride = 2;
            int L = stride * n + 1;
            if ((double)L > (double)Ts * 1.2) { free(slp); ok = 0; break; }         /* :427-429 */
            int32_t *path = (int32_t *)malloc(sizeof(int32_t) * (size_t)L);
            int32_t *tidx = (int32_t *)malloc(sizeof(int32_t) * (size_t)L);
            build_path(seq + s.t0, n, s.t0, stride, P->blank_id, path, tidx);
            int band = (L > 60) ? ((L / 3 > 30) ? L / 3 : 30) : 0;                  /* :441 */
            int32_t *sp = (int32_t *)malloc(sizeof(int32_t) * (size_t)Ts);
            int32_t *si = (int32_t *)malloc(sizeof(int32_t) * (size_t)Ts);
            float dpf;
            orc_viterbi(slp, Ts, C, path, tidx, L, band, P->blank_id, P->truly_forced, P->neg_inf,
                        sp, si, &dpf, NULL, NULL, NULL);
            if (dpf <= P->neg_inf && degenerate) *degenerate = 1;
            /* :447-448 python slicing clamps at the array end */
            int e = pad_left + nf; if (e > Ts) e = Ts;
            for (int f = pad_left; f < e; ++f) { ph[w] = sp[f]; ix[w] = si[f]; ++w; }
            free(path); free(tidx); free(sp); free(si); free(slp);
        }
    }
    if (ok && w == 0) ok = 0;                                     /* :454-455 */
    if (ok) {                                                     /* :457-467 */
        int n = w < T ? w : T;
        memcpy(out_ph, ph, sizeof(int32_t) * (size_t)n);
        memcpy(out_idx, ix, sizeof(int32_t) * (size_t)n);
        for (int t = n; t < T; ++t) { out_ph[t] = P->blank_id; out_idx[t] = -1; }
    }
    free(grp); free(asil); free(mt); free(segs); free(ph); free(ix);
    return ok;
}

/* decode_with_forced_alignment (:87-199) for one utterance.  lp is NOT
 * modified (the reference clones, :121).  Returns status (ORC_*), with
 * ORC_DEGENERATE or-ed in when the winning dp value is <= -1000.              */
int orc_decode_forced(const float *lp, int T, int C, const int32_t *seq, int N, const OrcParams *P,
                      int32_t *out_ph, int32_t *out_idx, float *dp_final)
{
    if (dp_final) *dp_final = 0.0f;
    if (N == 0) {                                                 /* :112-118 */
        for (int t = 0; t < T; ++t) { out_ph[t] = P->blank_id; out_idx[t] = -1; }
        return ORC_EMPTY_TARGET;
    }
    float *m = (float *)malloc(sizeof(float) * (size_t)T * C);
    memcpy(m, lp, sizeof(float) * (size_t)T * C);
    orc_prep(m, T, C, seq, N, P);                                 /* :123-129 */
    int status = ORC_OK, degenerate = 0;
    if (P->silence_anchors > 0 && P->silence_id >= 0) {           /* :133 ; anchor_pauses (:902) */
        if (T > 0 && segmented_decode(m, T, C, seq, N, P, out_ph, out_idx, &degenerate)) {
            free(m);
            return ORC_SEGMENTED | (degenerate ? ORC_DEGENERATE : 0);
        }
    }
    int stride = 4;                                               /* :153-157 */
    if (stride * N + 1 > T) stride = 3;
    if (stride * N + 1 > T) stride = 2;
    if (stride * N + 1 > T) stride = 1;
    int L = stride * N + 1;
    if (L > T) {                                                  /* :159-176 */
        free(m);
        if (T < N) return ORC_TOO_SHORT;
        for (int t = 0; t < T; ++t) {
            int j = (int)(((int64_t)t * N) / T);
            out_ph[t] = seq[j]; out_idx[t] = j;
        }
        return ORC_PROPORTIONAL;
    }
    int32_t *path = (int32_t *)malloc(sizeof(int32_t) * (size_t)L);
    int32_t *tidx = (int32_t *)malloc(sizeof(int32_t) * (size_t)L);
    build_path(seq, N, 0, stride, P->blank_id, path, tidx);       /* :181-187 */
    int band = (L > 60) ? ((L / 4 > 20) ? L / 4 : 20) : 0;        /* :190 */
    float dpf;
    orc_viterbi(m, T, C, path, tidx, L, band, P->blank_id, P->truly_forced, P->neg_inf,
                out_ph, out_idx, &dpf, NULL, NULL, NULL);
    if (dp_final) *dp_final = dpf;
    if (dpf <= P->neg_inf) status |= ORC_DEGENERATE;
    free(path); free(tidx); free(m);
    return status;
}

/* decode_alignments_simple body (:952-981) for one utterance.                 */
int orc_decode_simple(const float *lp, int T, int C, const int32_t *seq, int N, const OrcParams *P,
                      int32_t *out_ph, int32_t *out_idx, float *dp_final)
{
    int stride = 4;                                               /* :963-968 */
    if ((double)(stride * N + 1) > (double)T * 0.9) stride = 3;
    if ((double)(stride * N + 1) > (double)T * 0.8) stride = 2;
    int L = stride * N + 1;
    int32_t *path = (int32_t *)malloc(sizeof(int32_t) * (size_t)L);
    int32_t *tidx = (int32_t *)malloc(sizeof(int32_t) * (size_t)L);
    build_path(seq, N, 0, stride, P->blank_id, path, tidx);
    int band = (L > 60) ? ((L / 4 > 20) ? L / 4 : 20) : 0;        /* :976 */
    float dpf;
    orc_viterbi(lp, T, C, path, tidx, L, band, P->blank_id, P->truly_forced, P->neg_inf,
                out_ph, out_idx, &dpf, NULL, NULL, NULL);
    if (dp_final) *dp_final = dpf;
    free(path); free(tidx);
    return (dpf <= P->neg_inf) ? ORC_DEGENERATE : ORC_OK;
}

/* assort_frames (:777-834).  Returns number of stamps written (<= max).       */
int orc_assort(const int32_t *ph, const int32_t *idx, int T, int blank_id, int ignore_noise,
               int max_blanks, OrcStamp *out, int max_out)
{
    int n = 0;
    for (int s = 0; s < T;) {
        int e = s + 1;
        while (e < T && ph[e] == ph[e - 1] && idx[e] == idx[e - 1]) ++e;   /* :798-801 */
        int p = ph[s], ti = idx[s];
        if (ti == -1) for (int q = s; q < e; ++q) if (idx[q] != -1) { ti = idx[q]; break; } /* :812-816 */
        int emit = 0;
        if (p == blank_id) { if (!ignore_noise && (e - s) > max_blanks) emit = 1; }  /* :819-827 */
        else emit = 1;                                                               /* :830-831 */
        if (emit && n < max_out) { out[n].phoneme = p; out[n].start = s; out[n].end = e; out[n].target_idx = ti; ++n; }
        s = e;
    }
    return n;
}

/* utils.py:70-113 _calculate_confidences.  lp is the ORIGINAL [T,C] matrix.   */
void orc_confidence(const float *lp, int T, int C, const OrcStamp *st, int n, float *conf)
{
    for (int i = 0; i < n; ++i) {
        int ph = st[i].phoneme;
        int s = st[i].start < 0 ? 0 : st[i].start;                /* :86 */
        int e = st[i].end > T ? T : st[i].end;                    /* :87 */
        float avg = expf(lp[(size_t)s * C + ph]);                 /* :89 */
        if (s < e && ph < C) {                                    /* :93 */
            float half = avg / 2.0f; int good = 1;                /* :95-96 */
            for (int f = s + 1; f < e; ++f) {                     /* :99-103 */
                float p = expf(lp[(size_t)f * C + ph]);
                if (p > half || p > 0.1f) { avg += p; ++good; }
            }
            if (good > 1) {                                       /* :104-109 */
                avg /= (float)good;
                /* avg_confidence is a VIEW of probs[start, ph] (:89) and `+=` / `/=` are in-place, so
                 * by :107 probs[start, ph] already holds the average: the max runs over
                 * {avg, p[s+1..e-1]}, not over the original first-frame probability. */
                float mx = avg;
                for (int f = s + 1; f < e; ++f) { float p = expf(lp[(size_t)f * C + ph]); if (p > mx) mx = p; }
                if (avg < mx / 2.0f) avg = mx;
            }
        }
        conf[i] = avg;
    }
}

/* ------------------------------------------------------------------------- */
/* Batch driver == AlignmentUtils.decode_alignments (:856-910) + confidences   */
/* (core.py:935-937 minus coverage repair / boundary extension), threaded over */
/* utterances for the CPU baseline.                                            */
/* ------------------------------------------------------------------------------------------
 * PhonemeTimestampAligner.extend_soft_boundaries_func (core.py:682-809): four passes that stretch each stamp's start /
 * end over neighbouring frames while exp(lp[f, phoneme]) stays above a threshold.  Within a pass every stamp reads only
 * what earlier passes wrote (pass 1: previous ORIGINAL end, pass 2: next start after pass 1, pass 3: previous end after
 * pass 2, pass 4: next start after pass 3), so the passes are data-parallel over the stamps.
 * Thresholds are Python doubles (core.py:700-703); probabilities are float32 exp compared after promotion to double
 * (.item()).  mean_probs (:712-716) is a float32 tensor mean in the reference; it is restated as a double sum rounded to
 * float (within an ulp or two of torch's vectorised sum; it only scales a threshold). */
void orc_soft_boundaries(const float *lp, int T, int C, OrcStamp *st, int n, int boundary_softness)
{
    const double max_ext = 10.0;                                            /* :698 */
    const double t1 = pow(10.0, -1.0 * 3.0);                                /* :699-700: max(2, 7 - 4) = 3 */
    const double t2 = pow(10.0, -1.0 * (double)boundary_softness);          /* :701 */
    if (n <= 0) return;
    double *mean = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; ++i) {                                           /* :710-716 */
        const int ph = st[i].phoneme, s = st[i].start, e = st[i].end;
        if (s < T && ph < C && s < e) {
            const int ee = e < T ? e : T;
            double acc = 0.0;
            for (int f = s; f < ee; ++f) acc += (double)expf(lp[(size_t)f * C + ph]);
            mean[i] = (double)(float)(acc / (double)(ee - s));
        } else mean[i] = 0.001;
    }
    for (int i = 0; i < n; ++i) {                                           /* pass 1, :719-737 */
        const int ph = st[i].phoneme, s = st[i].start, e = st[i].end;
        if (s >= T || ph >= C) continue;
        const int dur = e - s;
        int lo = (int)((double)s - (double)dur * max_ext);
        if (lo < 0) lo = 0;
        if (i > 0) { int b = st[i - 1].end + 10; if (b > s) b = s; if (b > lo) lo = b; }
        const double thr = mean[i] * t1 < t1 ? mean[i] * t1 : t1;
        int ns = s;
        for (int f = s - 1; f >= lo; --f) { if ((double)expf(lp[(size_t)f * C + ph]) >= thr) ns = f; else break; }
        st[i].start = ns;
    }
    for (int i = 0; i < n; ++i) {                                           /* pass 2, :740-757 */
        const int ph = st[i].phoneme, s = st[i].start, e = st[i].end;
        if (s >= T || ph >= C) continue;
        const int dur = e - s;
        int hi = (int)((double)e + (double)dur * max_ext);
        if (hi > T) hi = T;
        if (i + 1 < n) { int b = st[i + 1].start - 10; if (b > e) b = e; if (b < hi) hi = b; }   /* :749: min(end, next - 10) */
        const double thr = mean[i] * t1 < t1 ? mean[i] * t1 : t1;
        int ne = e;
        for (int f = e; f < hi; ++f) { if ((double)expf(lp[(size_t)f * C + ph]) >= thr) ne = f + 1; else break; }
        st[i].end = ne;
    }
    for (int i = 0; i < n; ++i) {                                           /* pass 3, :760-780 */
        const int ph = st[i].phoneme, s = st[i].start;
        if (s >= T || ph >= C) continue;
        const int lo = i > 0 ? st[i - 1].end : 0;
        if (s <= lo) continue;
        int ns = s;
        for (int f = s - 1; f >= lo; --f) { if ((double)expf(lp[(size_t)f * C + ph]) >= t2) ns = f; else break; }
        st[i].start = ns;
    }
    for (int i = 0; i < n; ++i) {                                           /* pass 4, :784-805 */
        const int ph = st[i].phoneme, s = st[i].start, e = st[i].end;
        if (s >= T || ph >= C) continue;
        const int dur = e - s;
        int hi = (int)((double)e + (double)dur * max_ext);
        if (hi > T) hi = T;
        if (i + 1 < n && st[i + 1].start < hi) hi = st[i + 1].start;
        int ne = e;
        for (int f = e; f < hi; ++f) { if ((double)expf(lp[(size_t)f * C + ph]) >= t2) ne = f + 1; else break; }
        st[i].end = ne;
    }
    free(mean);
}

typedef struct {
    const float *lp; const int64_t *row_off; const int32_t *T; int C;
    const int32_t *tgt; const int64_t *tgt_off; const OrcParams *P;
    int32_t *frame_ph, *frame_idx; const int64_t *frame_off;
    float *dp_final; int32_t *status; OrcStamp *stamps; float *conf; int32_t *n_stamps; int max_stamps;
    int B, next; pthread_mutex_t mu;
} Batch;

static void *batch_worker(void *arg)
{
    Batch *b = (Batch *)arg;
    for (;;) {
        pthread_mutex_lock(&b->mu); int u = b->next++; pthread_mutex_unlock(&b->mu);
        if (u >= b->B) break;
        const float *lp = b->lp + b->row_off[u];
        int T = b->T[u]; int N = (int)(b->tgt_off[u + 1] - b->tgt_off[u]);
        const int32_t *seq = b->tgt + b->tgt_off[u];
        int32_t *ph = b->frame_ph + b->frame_off[u], *ix = b->frame_idx + b->frame_off[u];
        float dpf = 0.f; int st;
        if (N == 0) {                 /* decode_alignments: empty target -> empty result (:894-897) */
            st = ORC_EMPTY_TARGET; b->n_stamps[u] = 0;
            for (int t = 0; t < T; ++t) { ph[t] = b->P->blank_id; ix[t] = -1; }
        } else {
            st = (b->P->mode == 1) ? orc_decode_simple(lp, T, b->C, seq, N, b->P, ph, ix, &dpf)
                                   : orc_decode_forced(lp, T, b->C, seq, N, b->P, ph, ix, &dpf);
            if ((st & 7) == ORC_TOO_SHORT) b->n_stamps[u] = 0;
            else {
                OrcStamp *so = b->stamps + (size_t)u * b->max_stamps;
                int n = orc_assort(ph, ix, T, b->P->blank_id, b->P->ignore_noise, b->P->max_blanks, so, b->max_stamps);
                b->n_stamps[u] = n;
                if (b->conf) orc_confidence(lp, T, b->C, so, n, b->conf + (size_t)u * b->max_stamps);
            }
        }
        b->status[u] = st; if (b->dp_final) b->dp_final[u] = dpf;
    }
    return NULL;
}

int orc_align_batch(const OrcParams *P, const float *lp, const int64_t *row_off, const int32_t *T, int C,
                    const int32_t *tgt, const int64_t *tgt_off, int B,
                    int32_t *frame_ph, int32_t *frame_idx, const int64_t *frame_off,
                    float *dp_final, int32_t *status, OrcStamp *stamps, float *conf, int32_t *n_stamps,
                    int max_stamps, int n_threads)
{
    Batch b = {lp, row_off, T, C, tgt, tgt_off, P, frame_ph, frame_idx, frame_off,
               dp_final, status, stamps, conf, n_stamps, max_stamps, B, 0, PTHREAD_MUTEX_INITIALIZER};
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, batch_worker, &b);
    for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
    return 0;
}

void orc_default_params(OrcParams *P, int blank_id, int silence_id)
{
    memset(P, 0, sizeof(*P));
    P->blank_id = blank_id; P->silence_id = silence_id; P->silence_anchors = 10;
    P->ignore_noise = 1; P->truly_forced = 1; P->boost_targets = 1; P->enforce_minimum = 1;
    P->max_blanks = 10; P->boost_factor = 5.0f; P->min_log_prob = logf(1e-8f);
    P->neg_inf = -1000.0f; P->sub_boost = 5.0f; P->boundary_pad = 3; P->min_speech_frames = 20; P->mode = 0;
}

int orc_sizeof_params(void) { return (int)sizeof(OrcParams); }
