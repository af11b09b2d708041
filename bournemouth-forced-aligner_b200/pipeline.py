"""The post-acoustic part of PhonemeTimestampAligner ("the shim that accepts injected posteriors").

The reference computes phoneme / group log-posteriors with its acoustic model and then runs, per batch
(core.py:896-957): decode_alignments on both heads -> ensure_target_coverage -> extend_soft_boundaries_func ->
_calculate_confidences -> convert_to_ms -> sort by start time.  This class is that second half with the posteriors as an
argument: the alignment, the boundary extension and the confidences are one kernel launch sequence per batch on the GPU,
the list surgery stays on the host like in the reference."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .aligner import AlignmentUtils, _calculate_confidences_batch, extend_soft_boundaries_func
from .postprocess import convert_to_ms, ensure_target_coverage


class PhonemeTimestampAligner:
    """Constructor arguments carry the reference's names and defaults where it has them (core.py:41-125, :256-257)."""

    def __init__(self, blank_class=66, silence_class=0, blank_group=16, silence_group=0, silence_anchors=10, ignore_noise=True,
                 enforce_all_targets=True, boost_targets=True, enforce_minimum=True, extend_soft_boundaries=True, boundary_softness=3,
                 ensure_completeness=True, resampler_sample_rate=16000):
        self.blank_class, self.silence_class = blank_class, silence_class
        self.blank_group, self.silence_group = blank_group, silence_group
        self.boost_targets, self.enforce_minimum = boost_targets, enforce_minimum
        self.extend_soft_boundaries, self.boundary_softness = extend_soft_boundaries, boundary_softness
        self.ensure_completeness = ensure_completeness
        self.resampler_sample_rate = resampler_sample_rate
        self.alignment_utils_p = AlignmentUtils(blank_id=blank_class, silence_id=silence_class, silence_anchors=silence_anchors,   # :256-257
                                                ignore_noise=ignore_noise, truly_forced=enforce_all_targets)
        self.alignment_utils_g = AlignmentUtils(blank_id=blank_group, silence_id=silence_group, silence_anchors=silence_anchors,
                                                ignore_noise=ignore_noise, truly_forced=enforce_all_targets)
        self.stats = {}

    def _head(self, utils, log_probs, seqs, seq_lens, spectral_lens, wav_lens, offsets, silence):
        frames = utils.decode_alignments(log_probs, true_seqs=seqs, pred_lens=spectral_lens, true_seqs_lens=seq_lens,          # :902-922
                                         forced_alignment=True, boost_targets=self.boost_targets, enforce_minimum=self.enforce_minimum)
        frames = ensure_target_coverage(seqs, frames, seq_lens=seq_lens, _silence_class=silence,                               # :925-926
                                        ensure_completeness=self.ensure_completeness, stats=self.stats)
        if self.extend_soft_boundaries:                                                                                        # :928-931
            frames = extend_soft_boundaries_func(log_probs, frames, boundary_softness=self.boundary_softness)
        frames = _calculate_confidences_batch(log_probs, frames)                                                               # :936-937 (padded rows, like log_probs[b])
        out = []
        for b, fs in enumerate(frames):
            off = offsets[b] if isinstance(offsets, (list, tuple)) else offsets
            fs = convert_to_ms(fs, int(spectral_lens[b]), off, int(wav_lens[b]), self.resampler_sample_rate)                   # :939-952
            out.append(sorted(fs, key=lambda x: x[6]))                                                                         # :955-956
        return out

    def timestamps_from_posteriors(self, log_probs_p: torch.Tensor, ph_seqs: torch.Tensor, ph_seq_lens, spectral_lens, wav_lens,
                                   start_offset_times=0.0, log_probs_g: Optional[torch.Tensor] = None,
                                   grp_seqs: Optional[torch.Tensor] = None) -> List[dict]:
        """log_probs_p [B, T, C_p] (CUDA, log-softmaxed like core.py:898), ph_seqs [B, S] padded targets, ph_seq_lens / spectral_lens /
        wav_lens per utterance.  Returns the reference's `timestamp_dicts` (core.py:958-964): 8-tuples
        (id, start_frame, end_frame, target_idx, is_estimated, confidence, start_ms, end_ms) per head."""
        ph = self._head(self.alignment_utils_p, log_probs_p, ph_seqs, ph_seq_lens, spectral_lens, wav_lens, start_offset_times,
                        self.silence_class)
        if log_probs_g is not None and grp_seqs is not None:
            gr = self._head(self.alignment_utils_g, log_probs_g, grp_seqs, ph_seq_lens, spectral_lens, wav_lens, start_offset_times,
                            self.silence_group)
        else:
            gr = [None] * len(ph)
        return [{"phoneme_timestamps": ph[b], "group_timestamps": gr[b]} for b in range(len(ph))]
