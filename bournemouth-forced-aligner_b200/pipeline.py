"""PhonemeTimestampAligner without its acoustic model and phonemizer ("the shim that accepts injected
posteriors"; north_star: "keeping the PhonemeTimestampAligner / process_sentence API surface").

The reference class (core.py:34) owns three things: the CUPE acoustic model (out of scope), the espeak phonemizer (out of
scope) and everything between their outputs and the per-segment result dict.  This class is the third part with the first two
INJECTED:

  * `posterior_provider(wavs [B, S], wav_lens) -> (logits_class [B, T, Cp], logits_group [B, T, Cg], embeddings | None,
    spectral_lens)` stands where `_cupe_prediction_batch` (core.py:370-460) stands; it may instead return a dict with the
    un-stitched per-window logits (`window_logits_class/_group` [B, W, frames_per_window, C], `original_audio_length`,
    `window_size_ms`, `stride_ms`, `spectral_lens`), in which case stich_window_predictions + log_softmax run as one kernel;
  * `phonemizer.phonemize_sentence(text)` stands where core.py:1122-1138 stands (a record with the ph66 / pg16 / eipa / words /
    word_num keys) and carries the label tables (`index_to_plabel`, `index_to_glabel`, `phoneme_id_to_group_id`).

Method names, arguments, return values and error behaviour follow the reference: extract_timestamps_from_segment_batch
(core.py:811-992), extract_timestamps_from_segment_simplified (:995-1044), process_segments (:1212-1484), process_sentence
(:1553-1584), process_sentences_batch (:1586-1616).  Per batch the phoneme head and the group head are both enqueued before
the host looks at either result; alignment, boundary extension and confidences are kernels of libbfa_b200, the list surgery
(ensure_target_coverage, convert_to_ms, post_process_segment) stays on the host like in the reference.  Pooled embeddings
(weighted_pool_embeddings) belong to the acoustic side and are not produced: the embedding lists are lists of None."""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import torch

from . import _cabi
from .aligner import AlignmentUtils, _calculate_confidences_batch, _ptr, _stream, extend_soft_boundaries_func, log_softmax_rows
from .postprocess import convert_to_ms, ensure_target_coverage, post_process_segment


def stitch_log_softmax(window_logits: torch.Tensor, original_audio_length: int, cnn_output_size: Optional[int] = None,
                       sample_rate: int = 16000, window_size_ms: int = 160, stride_ms: int = 80) -> torch.Tensor:
    """F.log_softmax(stich_window_predictions(window_logits, ...), dim=2) -- cupe2i/windowing.py:103-173 + core.py:898-899 -- in
    one kernel (bfa_stitch_log_softmax).  window_logits: CUDA [B, num_windows, frames_per_window, C].  Returns [B, total_frames, C]."""
    if not window_logits.is_cuda:
        raise ValueError("window_logits must be a CUDA tensor (there is no CPU path)")
    x = window_logits if (window_logits.dtype == torch.float32 and window_logits.is_contiguous()) else window_logits.contiguous().float()
    B, W, fpw, C_ = x.shape
    cnn_output_size = fpw if cnn_output_size is None else cnn_output_size
    window_size_samples = int(window_size_ms * sample_rate / 1000)                       # :125-129
    stride_samples = int(stride_ms * sample_rate / 1000)
    num_windows_total = ((original_audio_length - window_size_samples) // stride_samples) + 1
    total_frames = (num_windows_total * cnn_output_size) // 2
    weights = torch.cos(torch.linspace(-math.pi / 2, math.pi / 2, fpw, device=x.device))  # :132, the reference's own expression
    out = torch.empty((B, max(total_frames, 0), C_), dtype=torch.float32, device=x.device)
    if W >= 2 and (W - 2) * (fpw // 2) + fpw > total_frames:      # a full window would not fit: the reference's slice assignment (:152) fails
        raise ValueError(f"stitch: window {W - 2} of {fpw} frames ends beyond the {total_frames} output frames")
    rc = _cabi.lib().bfa_stitch_log_softmax(B, W, fpw, C_, total_frames, _ptr(x), W * fpw * C_, _ptr(weights), _ptr(out),
                                            total_frames * C_, _stream(x.device))
    _cabi.check(rc)
    return out


class PhonemeTimestampAligner:
    """Constructor arguments carry the reference's names and defaults where it has them (core.py:41, :124-176, :230-257)."""

    def __init__(self, blank_class=66, silence_class=0, blank_group=16, silence_group=0, silence_anchors=10, ignore_noise=True,
                 enforce_all_targets=True, boost_targets=True, enforce_minimum=True, extend_soft_boundaries=True, boundary_softness=3,
                 ensure_completeness=False, resampler_sample_rate=16000, *, posterior_provider: Optional[Callable] = None,
                 phonemizer=None, duration_max=30, bad_confidence_threshold=0.6, device="cuda"):
        self.device = torch.device(device)
        self.warn_level = 1
        self.posterior_provider = posterior_provider
        self.phonemizer = phonemizer
        self.resampler_sample_rate = resampler_sample_rate
        self.sample_rate = 16000                                                          # :234
        self.ph_seq_min = 1                                                               # :131
        self.seg_duration_min = 0.05                                                      # :134-137
        self.seg_duration_min_samples = int(self.seg_duration_min * self.resampler_sample_rate)
        self.seg_duration_max = duration_max
        self.wav_len_max = int(self.seg_duration_max * self.resampler_sample_rate)
        if phonemizer is not None:                                                        # :147-154, :243-248
            self.phonemes_key, self.phoneme_groups_key = phonemizer.phonemes_key, phonemizer.phoneme_groups_key
            self.phoneme_id_to_label, self.group_id_to_label = phonemizer.index_to_plabel, phonemizer.index_to_glabel
            self.phoneme_id_to_group_id = phonemizer.phoneme_id_to_group_id
            p2i = {label: idx for idx, label in self.phoneme_id_to_label.items()}
            g2i = {label: idx for idx, label in self.group_id_to_label.items()}
            blank_class, blank_group = p2i["noise"], g2i["noise"]
            silence_class, silence_group = p2i["SIL"], g2i["SIL"]
        else:
            self.phonemes_key, self.phoneme_groups_key = "ph66", "pg16"
            self.phoneme_id_to_label, self.group_id_to_label, self.phoneme_id_to_group_id = {}, {}, {}
        self.blank_class, self.silence_class = blank_class, silence_class
        self.blank_group, self.silence_group = blank_group, silence_group
        self.silence_anchors = silence_anchors
        self.boost_targets, self.enforce_minimum = boost_targets, enforce_minimum
        self.enforce_all_targets, self.ignore_noise = enforce_all_targets, ignore_noise
        self.extend_soft_boundaries, self.boundary_softness = extend_soft_boundaries, boundary_softness
        self.ensure_completeness = ensure_completeness
        self.bad_confidence_threshold = bad_confidence_threshold
        self.alignment_utils_g = AlignmentUtils(blank_id=blank_group, silence_id=silence_group, silence_anchors=silence_anchors,   # :256-257
                                                ignore_noise=ignore_noise, truly_forced=enforce_all_targets)
        self.alignment_utils_p = AlignmentUtils(blank_id=blank_class, silence_id=silence_class, silence_anchors=silence_anchors,
                                                ignore_noise=ignore_noise, truly_forced=enforce_all_targets)
        self.stats = {}
        self.reset_counters()

    def reset_counters(self):                                                             # :187-197
        self.total_segments_processed = 0
        self.total_segments_bad = 0
        self.total_segments_failed = 0
        self.perfect_matches = 0
        self.stats.clear()

    # ---- the post-acoustic half of extract_timestamps_from_segment_batch --------------------------------------------------------
    def _prepare(self, utils, log_probs, seqs, seq_lens, spectral_lens, input_is_logits=False):
        return utils.decode_alignments_prepare(log_probs, true_seqs=seqs, pred_lens=spectral_lens, true_seqs_lens=seq_lens,         # :902-922
                                               boost_targets=self.boost_targets, enforce_minimum=self.enforce_minimum,
                                               input_is_logits=input_is_logits)

    def _finish(self, utils, handle, log_probs, seqs, seq_lens, spectral_lens, wav_lens, offsets, silence, input_is_logits=False):
        frames = utils.decode_alignments_finish(handle)
        frames = ensure_target_coverage(seqs, frames, seq_lens=seq_lens, _silence_class=silence,                               # :925-926
                                        ensure_completeness=self.ensure_completeness, stats=self.stats)
        lse_kw = {}
        if input_is_logits:
            # aligned straight from the logits: the steps below form log-probabilities from the rows' log-sum-exp on the fly;
            # otherwise the facade normalised the batch on its way to the full chain and that copy is used
            if utils.last_row_lse is not None:
                lse_kw = dict(row_lse=utils.last_row_lse, pred_lens=spectral_lens)
            else:
                log_probs = utils.last_log_probs
        if self.extend_soft_boundaries:                                                                                        # :928-931
            frames = extend_soft_boundaries_func(log_probs, frames, boundary_softness=self.boundary_softness, **lse_kw)
        frames = _calculate_confidences_batch(log_probs, frames, **lse_kw)                                                     # :936-937 (padded rows, like log_probs[b])
        out = []
        for b, fs in enumerate(frames):
            off = offsets[b] if isinstance(offsets, (list, tuple)) else offsets
            fs = convert_to_ms(fs, int(spectral_lens[b]), off, int(wav_lens[b]), self.resampler_sample_rate)                   # :939-952
            out.append(sorted(fs, key=lambda x: x[6]))                                                                         # :955-956
        return out

    def timestamps_from_posteriors(self, log_probs_p: torch.Tensor, ph_seqs: torch.Tensor, ph_seq_lens, spectral_lens, wav_lens,
                                   start_offset_times=0.0, log_probs_g: Optional[torch.Tensor] = None,
                                   grp_seqs: Optional[torch.Tensor] = None, input_is_logits: bool = False) -> List[dict]:
        """log_probs_p [B, T, C_p] (CUDA, log-softmaxed like core.py:898), ph_seqs [B, S] padded targets, ph_seq_lens / spectral_lens /
        wav_lens per utterance.  Returns the reference's `timestamp_dicts` (core.py:958-964): 8-tuples
        (id, start_frame, end_frame, target_idx, is_estimated, confidence, start_ms, end_ms) per head.
        Both heads are enqueued on the device before the host waits for either (core.py:900-922 runs them one after the other).
        input_is_logits=True: the tensors hold the acoustic model's un-normalised logits (core.py:898-899 skipped); a batch without
        silence_id in its targets is then aligned, stretched and scored without the log-probabilities ever being written."""
        groups = log_probs_g is not None and grp_seqs is not None
        # both heads are prepared (targets, offsets, plans on the device) before either alignment is enqueued: the two alignment
        # kernels sit next to each other in the stream and the second may start while the first drains
        lg = input_is_logits and self.boost_targets
        if input_is_logits and not lg:       # without boosting nothing re-normalises the rows: do what the reference does first
            log_probs_p = log_softmax_rows(log_probs_p)
            log_probs_g = log_softmax_rows(log_probs_g) if groups else log_probs_g
        hp = self._prepare(self.alignment_utils_p, log_probs_p, ph_seqs, ph_seq_lens, spectral_lens, lg)
        hg = self._prepare(self.alignment_utils_g, log_probs_g, grp_seqs, ph_seq_lens, spectral_lens, lg) if groups else None
        self.alignment_utils_p.decode_alignments_enqueue(hp)
        if groups:
            self.alignment_utils_g.decode_alignments_enqueue(hg, after_sibling=True)
        ph = self._finish(self.alignment_utils_p, hp, log_probs_p, ph_seqs, ph_seq_lens, spectral_lens, wav_lens, start_offset_times,
                          self.silence_class, lg)
        gr = (self._finish(self.alignment_utils_g, hg, log_probs_g, grp_seqs, ph_seq_lens, spectral_lens, wav_lens, start_offset_times,
                           self.silence_group, lg) if groups else [None] * len(ph))
        return [{"phoneme_timestamps": ph[b], "group_timestamps": gr[b]} for b in range(len(ph))]

    # ---- acoustic side (injected) ------------------------------------------------------------------------------------------------
    def _log_posteriors(self, wavs, wav_lens, extract_embeddings, keep_logits=False):
        """`_cupe_prediction_batch` (core.py:370-460) + log_softmax (:898-899): from the injected provider."""
        if self.posterior_provider is None:
            raise AssertionError("posterior provider is not set (the reference asserts that the CUPE extractor is loaded, core.py:890)")
        r = self.posterior_provider(wavs, wav_lens)
        if isinstance(r, dict):          # un-stitched window logits: stitch + log_softmax in one pass
            kw = dict(original_audio_length=r["original_audio_length"], sample_rate=r.get("sample_rate", self.sample_rate),
                      window_size_ms=r["window_size_ms"], stride_ms=r["stride_ms"])
            lp_p = stitch_log_softmax(r["window_logits_class"].to(self.device), **kw)
            lp_g = stitch_log_softmax(r["window_logits_group"].to(self.device), **kw) if r.get("window_logits_group") is not None else None
            return lp_p, lp_g, r["spectral_lens"], False
        logits_class, logits_group, _emb, spectral_lens = r
        if keep_logits:                  # the aligner takes the stitched logits as they are (timestamps_from_posteriors(input_is_logits=True))
            f32 = lambda t: t.to(self.device).contiguous().float()
            return f32(logits_class), (f32(logits_group) if logits_group is not None else None), spectral_lens, True
        lp_p = log_softmax_rows(logits_class.to(self.device))
        lp_g = log_softmax_rows(logits_group.to(self.device)) if logits_group is not None else None
        return lp_p, lp_g, spectral_lens, False

    def _map_phonemes_to_groups(self, phoneme_sequence):                                  # :1048-1059
        return torch.tensor([self.phoneme_id_to_group_id.get(int(p), self.blank_group) for p in phoneme_sequence], dtype=torch.long)

    def extract_timestamps_from_segment_batch(self, wavs, wav_lens, phoneme_sequences, start_offset_times=0, group_sequences=None,
                                              extract_embeddings=True, do_groups=True, debug=True):
        """core.py:811-992.  Returns (timestamp_dicts, [None] * B, [None] * B)."""
        if isinstance(phoneme_sequences, torch.Tensor):                                   # :841-846
            ph_seq_lens = [(seq != self.blank_class).sum().item() for seq in phoneme_sequences]
        else:
            ph_seq_lens = [len(seq) for seq in phoneme_sequences]
        if not isinstance(phoneme_sequences, torch.Tensor):                               # :849-852
            max_len = max(len(seq) for seq in phoneme_sequences)
            phoneme_sequences = torch.tensor([list(seq) + [self.blank_class] * (max_len - len(seq)) for seq in phoneme_sequences], dtype=torch.long)
        if group_sequences is not None and not isinstance(group_sequences, torch.Tensor):   # :854-857
            max_len = max(len(seq) for seq in group_sequences)
            group_sequences = torch.tensor([list(seq) + [self.blank_group] * (max_len - len(seq)) for seq in group_sequences], dtype=torch.long)
        if group_sequences is None:                                                       # :861-886
            mapped = [self._map_phonemes_to_groups(row.tolist()) for row in phoneme_sequences]
            if len(mapped) == 0:
                group_sequences = torch.empty(0, dtype=torch.long)
            else:
                max_len = max(m.size(0) for m in mapped)
                group_sequences = torch.stack([m if m.size(0) == max_len else torch.nn.functional.pad(m, (0, max_len - m.size(0)), value=self.blank_group)
                                               for m in mapped], dim=0)
        log_probs_p, log_probs_g, spectral_lens, are_logits = self._log_posteriors(wavs, wav_lens, extract_embeddings, keep_logits=True)
        spectral_lens = [int(v) for v in (spectral_lens.tolist() if isinstance(spectral_lens, torch.Tensor) else spectral_lens)]
        # the reference aligns the group head whatever do_groups says (:914-922); its result is dropped by process_segments then
        dicts = self.timestamps_from_posteriors(log_probs_p, phoneme_sequences, torch.tensor(ph_seq_lens, dtype=torch.long), spectral_lens,
                                                wav_lens, start_offset_times, log_probs_g, group_sequences if log_probs_g is not None else None,
                                                input_is_logits=are_logits)
        for d in dicts:
            if d["group_timestamps"] is None:
                d["group_timestamps"] = []
        return dicts, [None] * len(dicts), [None] * len(dicts)

    def extract_timestamps_from_segment_simplified(self, wavs, wav_lens, phoneme_sequences, start_offset_times=0.0, debug=True):
        """core.py:995-1044: log_softmax -> decode_alignments_simple -> convert_to_ms."""
        if isinstance(phoneme_sequences, torch.Tensor):
            ph_seq_lens = [(seq != self.blank_class).sum().item() for seq in phoneme_sequences]
        else:
            ph_seq_lens = [len(seq) for seq in phoneme_sequences]
            max_len = max(len(seq) for seq in phoneme_sequences)
            phoneme_sequences = torch.tensor([list(seq) + [self.blank_class] * (max_len - len(seq)) for seq in phoneme_sequences], dtype=torch.long)
        log_probs_p, _, spectral_lens, _ = self._log_posteriors(wavs, wav_lens, False)
        spectral_lens = [int(v) for v in (spectral_lens.tolist() if isinstance(spectral_lens, torch.Tensor) else spectral_lens)]
        frames = self.alignment_utils_p.decode_alignments_simple(log_probs_p, true_seqs=phoneme_sequences.to(self.device),
                                                                 pred_lens=torch.tensor(spectral_lens, dtype=torch.long),
                                                                 true_seqs_lens=torch.tensor(ph_seq_lens, dtype=torch.long))
        for b in range(len(frames)):
            off = start_offset_times[b] if isinstance(start_offset_times, (list, tuple)) else start_offset_times
            frames[b] = convert_to_ms(frames[b], spectral_lens[b], off, wav_lens[b], self.resampler_sample_rate)
        return [{"phoneme_timestamps": frames[b]} for b in range(len(frames))]

    # ---- text + audio in, per-segment records out ----------------------------------------------------------------------------------
    @staticmethod
    def _rms_normalize(audio):                                                            # :323-330
        rms = torch.sqrt(torch.mean(audio ** 2))
        return audio / rms if rms > 0 else audio

    def chop_wav(self, wav, start_frame, end_frame):
        """core.py:288-320."""
        num_frames = (end_frame - start_frame) if (end_frame != -1) else -1
        if num_frames < self.seg_duration_min_samples:
            return None, None, -1
        wav = wav[:, start_frame:end_frame]
        if wav.shape[1] < self.seg_duration_min_samples:
            return None, None, -2
        wav = self._rms_normalize(wav.mean(dim=0))
        wav_len = wav.shape[0]
        if wav_len > self.wav_len_max:
            wav = wav[:self.wav_len_max]
            wav_len = wav.shape[0]
        else:
            wav = torch.nn.functional.pad(wav, (0, self.wav_len_max - wav.shape[0]), "constant", 0)
        return wav, wav_len, 0

    def phonemize_sentence(self, text):                                                   # :1122-1138
        return self.phonemizer.phonemize_sentence(text)

    def post_process_segment(self, segment, ts, phoneme_sequence, phoneme_timestamps, group_timestamps=None, debug=False):
        return post_process_segment(segment, ts, phoneme_sequence, phoneme_timestamps, group_timestamps,                      # :1140-1210
                                    index_to_plabel=self.phoneme_id_to_label, index_to_glabel=self.group_id_to_label)

    def process_segments(self, srt_data, audio_wavs, extract_embeddings=False, do_groups=False, batch_size=16, debug=False):
        """core.py:1212-1484: clips with time-bounded text segments in, the reference's per-segment records out."""
        if isinstance(audio_wavs, torch.Tensor):                                          # :1236-1241
            if audio_wavs.dim() == 3:
                audio_wavs = [audio_wavs[i] for i in range(audio_wavs.size(0))]
            elif audio_wavs.dim() == 2:
                audio_wavs = [audio_wavs]
            else:
                raise ValueError(f"Expected audio_wavs of 2D (C,T) or 3D (B,C,T), got {audio_wavs.dim()}D")
        if isinstance(srt_data, dict):
            srt_data = [srt_data]
        if len(srt_data) != len(audio_wavs):
            raise ValueError(f"Batch size mismatch: {len(srt_data)} srt items vs {len(audio_wavs)} audio waveforms.")
        for bi, item in enumerate(srt_data):                                              # :1250-1255
            if "segments" not in item:
                raise ValueError(f"Batch item {bi} missing 'segments' key. Keys: {list(item.keys())}")
            for si, seg in enumerate(item["segments"]):
                if not all(k in seg for k in ("start", "end", "text")):
                    raise ValueError(f"Batch {bi}, segment {si} missing required keys (start/end/text). Has: {list(seg.keys())}")
        num_batch = len(srt_data)
        flat_items = [(bi, seg, clip) for bi, (item, clip) in enumerate(zip(srt_data, audio_wavs)) for seg in item["segments"]]
        empty3 = lambda res: (res, [[] for _ in range(num_batch)], [[] for _ in range(num_batch)]) if extract_embeddings else res
        if not flat_items:
            return empty3([{"segments": []} for _ in range(num_batch)])
        ts_outs = [self.phonemize_sentence(seg["text"]) for _, seg, _ in flat_items]      # :1272-1279
        phoneme_sequences = [ts[self.phonemes_key] for ts in ts_outs]
        group_sequences = [ts[self.phoneme_groups_key] for ts in ts_outs] if do_groups else [None] * len(flat_items)
        for (_, seg, _), ph_seq, grp_seq in zip(flat_items, phoneme_sequences, group_sequences):
            seg[self.phonemes_key] = ph_seq
            seg[self.phoneme_groups_key] = grp_seq
        valid = [i for i, ph_seq in enumerate(phoneme_sequences) if ph_seq and len(ph_seq) >= self.ph_seq_min]   # :1282-1288
        batch_results = [{"segments": []} for _ in range(num_batch)]
        if not valid:
            return empty3(batch_results)
        flat_f = [flat_items[i] for i in valid]
        ph_f = [phoneme_sequences[i] for i in valid]
        grp_f = [group_sequences[i] for i in valid]
        ts_f = [ts_outs[i] for i in valid]
        chopped = [self.chop_wav(clip, int(seg["start"] * self.resampler_sample_rate), int(seg["end"] * self.resampler_sample_rate))   # :1307-1315
                   for _, seg, clip in flat_f]
        wavs, wav_lens, codes = zip(*chopped)
        ok = [i for i, c in enumerate(codes) if c == 0]
        if len(ok) < len(flat_f):                                                         # :1318-1339
            flat_f, ph_f, grp_f, ts_f = [flat_f[i] for i in ok], [ph_f[i] for i in ok], [grp_f[i] for i in ok], [ts_f[i] for i in ok]
            wavs, wav_lens = [wavs[i] for i in ok], [wav_lens[i] for i in ok]
        if not wavs:
            raise ValueError("All segments have audio chopping errors. Cannot proceed with timestamp extraction.")
        wavs = torch.stack(list(wavs), dim=0)
        wav_lens = list(wav_lens)
        start_times = [seg["start"] for _, seg, _ in flat_f]
        call = lambda sl: self.extract_timestamps_from_segment_batch(
            wavs[sl], wav_lens[sl], ph_f[sl], start_offset_times=start_times[sl], group_sequences=grp_f[sl] if do_groups else None,
            extract_embeddings=extract_embeddings, do_groups=do_groups, debug=debug)[0]
        if batch_size < len(flat_items):                                                  # :1350-1393
            results = []
            for i in range(0, len(flat_f), batch_size):
                try:
                    results.extend(call(slice(i, i + batch_size)))
                except ValueError:       # "Audio too short to align": the whole slice comes back empty, like the reference (:1372-1390)
                    results.extend([{"phoneme_timestamps": [], "group_timestamps": []} for _ in range(batch_size)])
        else:
            results = call(slice(0, len(flat_f)))
        for (bi, seg, _), result, ts in zip(flat_f, results, ts_f):                       # :1406-1414
            batch_results[bi]["segments"].append(self.post_process_segment(
                seg, ts, seg[self.phonemes_key], result["phoneme_timestamps"], result["group_timestamps"] if do_groups else None, debug=debug))
        for bi, item in enumerate(batch_results):                                         # :1427-1468 confidence analysis
            for si, seg_out in enumerate(item["segments"]):
                self.total_segments_processed += 1
                if not seg_out.get("phoneme_ts"):
                    self.total_segments_failed += 1
                    continue
                phoneme_ts = seg_out["phoneme_ts"]
                if [t["phoneme_id"] for t in phoneme_ts] == seg_out[self.phonemes_key]:
                    self.perfect_matches += 1
                if len(phoneme_ts) > 60:
                    conf = [t["confidence"] for t in phoneme_ts]
                    if sum(1 for c in conf if c < 0.5) / len(conf) > self.bad_confidence_threshold:
                        seg_out["coverage_analysis"]["bad_alignment"] = True
                        self.total_segments_bad += 1
                    first_20, last_20 = sum(conf[10:30]) / 20, sum(conf[-30:-10]) / 20
                    if first_20 > 0.1 and last_20 < 0.1:
                        if self.silence_anchors == 0:
                            raise Exception(f"Bad confidence pattern in clip {bi}, segment {si+1}: first 20 avg {first_20:.3f} vs last 20 avg "
                                            f"{last_20:.3f}. Consider setting `silence_anchors=3`.")
                        seg_out["coverage_analysis"]["bad_alignment"] = True
                        self.total_segments_bad += 1
        return empty3(batch_results)

    def process_sentence(self, text, audio_wav, extract_embeddings=False, do_groups=False, debug=False):
        """core.py:1553-1584."""
        duration = audio_wav.shape[1] / self.sample_rate
        srt_data = [{"segments": [{"start": 0.0, "end": duration, "text": text.strip()}]}]
        result = self.process_segments(srt_data, [audio_wav], extract_embeddings=extract_embeddings, do_groups=do_groups, debug=debug)
        if extract_embeddings:
            out, p_emb, g_emb = result
            return out[0], p_emb[0], g_emb[0]
        return result[0]

    def process_sentences_batch(self, texts, audio_wavs, extract_embeddings=False, do_groups=False, debug=False):
        """core.py:1586-1616."""
        assert len(texts) == len(audio_wavs), f"Number of texts ({len(texts)}) must match number of audio waveforms ({len(audio_wavs)})"
        srt_data = [{"segments": [{"start": 0.0, "end": wav.shape[1] / self.sample_rate, "text": text.strip()}]} for text, wav in zip(texts, audio_wavs)]
        return self.process_segments(srt_data, audio_wavs, extract_embeddings=extract_embeddings, do_groups=do_groups, debug=debug)
