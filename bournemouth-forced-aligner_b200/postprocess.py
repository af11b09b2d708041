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
