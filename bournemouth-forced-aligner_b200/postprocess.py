"""Host epilogue after the alignment path: frames -> milliseconds.

`convert_to_ms` mirrors utils.convert_to_ms (reference utils.py:115-149): same arguments, same 8-tuples, the same
double-precision operations in the same order, so the millisecond values are bit-identical.  `stamps_to_ms` does the
same for the fixed-pitch arrays of a BatchResult (numpy, whole batch at once)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def _seconds_per_frame(spectral_length, wav_len, sample_rate) -> float:
    total = wav_len / sample_rate                                   # utils.py:127
    return total / spectral_length if spectral_length > 0 else 0    # :128


def convert_to_ms(framestamps: Sequence[tuple], spectral_length, start_offset_time, wav_len, sample_rate) -> List[tuple]:
    """(phoneme, start, end, target_idx, is_estimated, confidence) -> the same + (start_ms, end_ms).  Shorter tuples get the
    reference's defaults (target_idx -1, is_estimated False, confidence 0.0; utils.py:134-139)."""
    spf = _seconds_per_frame(spectral_length, wav_len, sample_rate)
    out = []
    for t in framestamps:
        ph, s, e = t[0], t[1], t[2]
        idx = t[3] if len(t) > 3 else -1
        est = t[4] if len(t) > 4 else False
        conf = t[5] if len(t) > 5 else 0.0
        s_ms = (start_offset_time + (s * spf)) * 1000               # :142-146
        e_ms = (start_offset_time + (e * spf)) * 1000
        out.append((ph, s, e, idx, est, conf, s_ms, e_ms))
    return out


def stamps_to_ms(stamps: np.ndarray, n_stamps: np.ndarray, spectral_lengths, start_offsets, wav_lens, sample_rate) -> Tuple[np.ndarray, np.ndarray]:
    """Vectorised over a batch: stamps int32 [B, P, 4] (BfaStamp layout), n_stamps [B]; per-utterance spectral length, start
    offset (s) and audio length (samples).  Returns (start_ms, end_ms) float64 [B, P] (entries past n_stamps are 0)."""
    stamps = np.asarray(stamps)
    B, P = stamps.shape[0], stamps.shape[1]
    sl = np.asarray(spectral_lengths, np.float64).reshape(B)
    total = np.asarray(wav_lens, np.float64).reshape(B) / float(sample_rate)
    spf = np.where(sl > 0, total / np.where(sl > 0, sl, 1.0), 0.0)[:, None]
    off = np.asarray(start_offsets, np.float64).reshape(B, 1)
    live = np.arange(P)[None, :] < np.asarray(n_stamps).reshape(B, 1)
    s_ms = (off + stamps[:, :, 1].astype(np.float64) * spf) * 1000.0
    e_ms = (off + stamps[:, :, 2].astype(np.float64) * spf) * 1000.0
    return np.where(live, s_ms, 0.0), np.where(live, e_ms, 0.0)


# ------------------------------------------------------------------------------------------------------------------------
# Coverage repair: PhonemeTimestampAligner.ensure_target_coverage, core.py:462-679.
# Host-side list surgery between decode_alignments and the boundary extension / confidence pass (core.py:925-937): every target
# phoneme ends up with exactly one stamp -- stamps with indices outside the target are dropped, repeated indices are merged or
# reduced to their longest piece, missing indices get estimated one-frame-or-more stamps placed between their aligned
# neighbours.  Returns 5-tuples (phoneme, start, end, target_idx, is_estimated).
# ------------------------------------------------------------------------------------------------------------------------
class CoverageError(Exception):
    """The reference raises a bare Exception when the repaired list does not cover the target exactly (core.py:676-677)."""


def _dedupe(stamps, repeated):
    """Stamps whose target index occurs more than once: pieces that touch or overlap are fused, then the longest piece
    (the earliest among equals) stands for the index (core.py:514-537).  Other stamps keep their order and come first."""
    single = [s for s in stamps if int(s[3]) not in repeated]
    pieces = {}
    for s in stamps:
        if int(s[3]) in repeated:
            pieces.setdefault(int(s[3]), []).append(s)
    for idx in sorted(pieces):
        fused = []
        for s in sorted(pieces[idx], key=lambda x: x[1]):
            if fused and s[1] <= fused[-1][2]:
                p = fused[-1]
                fused[-1] = (p[0], p[1], max(p[2], s[2]), p[3])
            else:
                fused.append((s[0], s[1], s[2], s[3]))
        single.append(max(fused, key=lambda x: x[2] - x[1]))
    return single


def _make_room_at_the_end(stamps, need, silence_class):
    """Trailing targets without an aligned successor (core.py:590-626): free `need` frames at the end of the utterance by
    shortening stamps from the back -- silence stamps first, then any stamp longer than one frame -- and pulling everything
    behind a shortened stamp forward by the same amount."""
    seq = sorted(stamps, key=lambda x: x[1])
    freed = 0
    for only_silence in (True, False):
        for k in range(len(seq) - 1, -1, -1):
            if freed >= need:
                break
            s = seq[k]
            span = s[2] - s[1]
            if span > 1 and (s[0] == silence_class or not only_silence):
                give = min(span - 1, need - freed)
                seq[k] = (*s[:2], s[2] - give, *s[3:])
                for j in range(k + 1, len(seq)):
                    t = seq[j]
                    seq[j] = (t[0], t[1] - give, t[2] - give, *t[3:])
                freed += give
        if freed >= need:
            break
    return seq


def ensure_target_coverage(phoneme_sequences, aligned_frames, seq_lens=None, _silence_class=0, debug=False, ensure_completeness=True,
                           stats=None):
    """core.py:462-679 with the same arguments (the reference reads `ensure_completeness` from the aligner object; `stats`, if a
    dict, receives the counters it keeps there).  `aligned_frames` is updated in place and returned, like the reference does."""
    for b in range(len(aligned_frames)):
        stamps = aligned_frames[b]
        frames_end = max((s[2] for s in stamps), default=0)                                   # "ctc_len", :479
        seq = phoneme_sequences[b]
        seq = seq.tolist() if hasattr(seq, "tolist") else list(seq)
        targets = seq[: (int(seq_lens[b]) if seq_lens is not None else len(seq))]
        n_t = len(targets)
        seen = [0] * n_t
        outside = set()
        for s in stamps:                                                                     # :487-493 (a -1 never counts)
            i = int(s[3])
            if i < n_t and i != -1:
                seen[i] += 1
            else:
                outside.add(i)
        repeated = {i for i, c in enumerate(seen) if c > 1}
        missing = [i for i, c in enumerate(seen) if c == 0]
        if stats is not None:
            stats["aligned"] = stats.get("aligned", 0) + len(stamps)
            stats["target"] = stats.get("target", 0) + n_t
            stats["extra"] = stats.get("extra", 0) + sum(seen[i] - 1 for i in repeated)
            stats["missed"] = stats.get("missed", 0) + len(missing)
        if outside:                                                                          # :509-513
            stamps = [s for s in stamps if int(s[3]) not in outside]
        if repeated and ensure_completeness:
            stamps = _dedupe(stamps, repeated)
        skipped_silence = set()
        if missing and ensure_completeness:                                                  # :540-651
            by_idx = {int(s[3]): s for s in stamps}
            runs = [[missing[0]]]
            for i in missing[1:]:
                if i == runs[-1][-1] + 1:
                    runs[-1].append(i)
                else:
                    runs.append([i])
            for run in runs:
                before = next((by_idx[i] for i in range(run[0] - 1, -1, -1) if i in by_idx), None)
                after = next((by_idx[i] for i in range(run[-1] + 1, n_t) if i in by_idx), None)
                if before and not after:          # the tail of the target: silence needs no stamp, the rest gets one frame each
                    wanted = [i for i in run if targets[i] != _silence_class]
                    skipped_silence.update(i for i in run if targets[i] == _silence_class)
                    if not wanted:
                        continue
                    stamps = _make_room_at_the_end(stamps, len(wanted), _silence_class)
                    by_idx = {int(s[3]): s for s in stamps}
                    tail = max(s[2] for s in stamps)
                    for k, i in enumerate(wanted):
                        a = min(tail + k, frames_end - 1)
                        new = (targets[i], a, min(a + 1, frames_end), i, True)
                        stamps.append(new)
                        by_idx[i] = new
                    continue
                if before and after:
                    lo, hi = before[2], after[1]
                elif after:
                    hi = after[1]
                    lo = max(0, hi - len(run))
                else:
                    lo, hi = 0, len(run)
                step = max(hi - lo, len(run)) / len(run)                                     # :640-641, Python float
                for k, i in enumerate(run):
                    a = min(int(lo + k * step), frames_end - 1)
                    z = min(max(int(lo + (k + 1) * step), a + 1), frames_end)
                    new = (targets[i], a, z, i, True)
                    stamps.append(new)
                    by_idx[i] = new
        stamps.sort(key=lambda x: x[1])                                                      # :654 (stable)
        easy = 0
        for k, s in enumerate(stamps):
            if len(s) == 4:
                stamps[k] = (*s, False)
                easy += 1
        if stats is not None:
            stats["aligned_easily"] = stats.get("aligned_easily", 0) + easy
        aligned_frames[b] = stamps
        if ensure_completeness:                                                              # :663-677
            want = n_t - len(skipped_silence)
            cover = [0] * n_t
            for s in stamps:
                i = int(s[3])
                if i < n_t and i != -1:
                    cover[i] += 1
            if sum(cover) != want or len(stamps) != want:
                raise CoverageError(f"Post-processing error: target coverage mismatch for segment {b}. Expected {want}, got {sum(cover)} "
                                    f"covered, {len(stamps)} aligned. Skipped SIL: {skipped_silence}")
    return aligned_frames


# ------------------------------------------------------------------------------------------------------------------------
# Word timestamps: PhonemeTimestampAligner._align_words, core.py:1062-1120.
# ------------------------------------------------------------------------------------------------------------------------
def align_words(phoneme_ts, word_num, words_list):
    """Groups consecutive phoneme timestamps (dicts with phoneme_id, ipa_label, start_ms, end_ms, confidence) by their word
    index into word entries {word, start_ms, end_ms, confidence, ph66, ipa}.  Same walk as the reference, including what it
    does at the last position: the phoneme at index len(word_num) - 1 closes the current word (joining it only if it carries
    the same word index), and nothing is emitted after it."""
    if not phoneme_ts or not word_num:
        return []
    n = min(len(word_num), len(phoneme_ts))
    last = len(word_num) - 1
    out = []
    cur, start, members = word_num[0], phoneme_ts[0]["start_ms"], []
    for i in range(n):
        ph = phoneme_ts[i]
        if word_num[i] == cur and i != last:
            members.append(ph)
            continue
        if i == last and word_num[i] == cur:
            members.append(ph)
        out.append({"word": words_list[cur] if cur < len(words_list) else f"UNK_WORD_{cur}",
                    "start_ms": start, "end_ms": members[-1]["end_ms"],
                    "confidence": sum(m["confidence"] for m in members) / len(members),
                    "ph66": [m["phoneme_id"] for m in members], "ipa": [m["ipa_label"] for m in members]})
        if i < last:
            cur, start, members = word_num[i], ph["start_ms"], [ph]
    return out


# ------------------------------------------------------------------------------------------------------------------------
# The per-segment output record: analyze_alignment_coverage (core.py:1701-1732) and
# post_process_segment (core.py:1140-1210) -- the dict that becomes the `.vs.json` file.  The label tables come from the
# phonemizer in the reference (index_to_plabel / index_to_glabel); here they are arguments.
# ------------------------------------------------------------------------------------------------------------------------
def analyze_alignment_coverage(target_sequence, aligned_timestamps, index_to_label):
    wanted = set(target_sequence.tolist() if hasattr(target_sequence, "tolist") else target_sequence)
    found = set([t[0] for t in aligned_timestamps])
    missing, extra = wanted - found, found - wanted
    ratio = len(wanted - missing) / len(wanted) if wanted else 1.0
    return {"target_count": len(wanted), "aligned_count": len(found), "missing_count": len(missing), "extra_count": len(extra),
            "coverage_ratio": ratio,
            "missing_phonemes": [index_to_label.get(p, f"UNK_{p}") for p in missing],
            "extra_phonemes": [index_to_label.get(p, f"UNK_{p}") for p in extra],
            "bad_alignment": ratio < 0.8}


def post_process_segment(segment, ts, phoneme_sequence, phoneme_timestamps, group_timestamps=None, *, index_to_plabel, index_to_glabel=None):
    """segment: the caller's dict (copied); ts: the phonemizer's record (eipa / word_num / words); phoneme_timestamps /
    group_timestamps: 8-tuples (id, start_frame, end_frame, target_idx, is_estimated, confidence, start_ms, end_ms) as produced by
    convert_to_ms.  Returns the reference's output dict: coverage_analysis, ipa, word_num, words, phoneme_ts[, group_ts], words_ts."""
    out = segment.copy()
    out["coverage_analysis"] = analyze_alignment_coverage(phoneme_sequence, phoneme_timestamps, index_to_plabel)
    out["ipa"] = ts.get("eipa", "")
    out["word_num"] = ts.get("word_num", "")
    out["words"] = ts.get("words", "")
    ipa = out["ipa"]
    out["phoneme_ts"] = [{"phoneme_id": int(p[0]), "phoneme_label": index_to_plabel.get(p[0], f"UNK_{p[0]}"),
                          "ipa_label": ipa[p[3]] if 0 <= p[3] < len(ipa) else "overflow",
                          "start_ms": float(p[6]), "end_ms": float(p[7]), "confidence": float(p[5]), "is_estimated": bool(p[4]),
                          "target_seq_idx": int(p[3]), "index": k} for k, p in enumerate(phoneme_timestamps)]
    if group_timestamps is not None:
        labels = index_to_glabel if index_to_glabel is not None else {}
        out["group_ts"] = [{"group_id": int(q[0]), "group_label": labels.get(q[0], f"UNK_{q[0]}"), "start_ms": float(q[6]),
                            "end_ms": float(q[7]), "confidence": float(q[5]), "is_estimated": bool(q[4]), "target_seq_idx": int(q[3]),
                            "index": k} for k, q in enumerate(group_timestamps)]
    out["words_ts"] = align_words(out["phoneme_ts"], ts.get("word_num", []), ts.get("words", []))
    return out
