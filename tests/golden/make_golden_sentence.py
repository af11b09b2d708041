#!/usr/bin/env python
"""tests/golden/sentence.json + stitch.npz: the UNMODIFIED reference PhonemeTimestampAligner (core.py) driven through
process_sentence / process_sentences_batch / process_segments with its two out-of-scope parts replaced by deterministic
stand-ins -- `_cupe_prediction_batch` is bfa_b200.synth.PlantedPosteriorProvider (seeded logits), the phonemizer is bfa_b200.synth.FakePhonemizer -- and
stich_window_predictions + F.log_softmax on seeded window logits (build container only: imports /root/reference with the
`phonemizer` package stubbed, it is not installed here and not on the path)."""
import json, sys, types
from pathlib import Path
import numpy as np
import torch

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
sys.path.insert(0, "/root/reference")
for name, attr in (("phonemizer", None), ("phonemizer.backend", "EspeakBackend"), ("phonemizer.separator", "Separator")):
    m = types.ModuleType(name)
    if attr: setattr(m, attr, type(attr, (), {}))
    sys.modules[name] = m
import bournemouth_aligner.core as core                      # noqa: E402
from bournemouth_aligner.cupe2i.windowing import stich_window_predictions   # noqa: E402
from bfa_b200 import synth                                   # noqa: E402

HERE = Path(__file__).resolve().parent
SR = 16000


def reference_aligner(**kw):
    """The reference object without its constructor's model / espeak loading: the attributes __init__ sets (core.py:124-185)."""
    a = object.__new__(core.PhonemeTimestampAligner)
    a.warn_level = 0; a.device = "cpu"; a.extractor = object(); a.lang = "en-us"
    a.resampler_sample_rate = SR; a.padding_ph_label = -100; a.ph_seq_min = 1
    a.seg_duration_min = 0.05; a.seg_duration_min_samples = int(0.05 * SR); a.seg_duration_max = 30; a.wav_len_max = 30 * SR
    a.phonemizer = synth.FakePhonemizer()
    a.phonemes_key, a.phoneme_groups_key = a.phonemizer.phonemes_key, a.phonemizer.phoneme_groups_key
    a.phoneme_id_to_label = a.phonemizer.index_to_plabel
    a.phoneme_label_to_id = {l: i for i, l in a.phoneme_id_to_label.items()}
    a.group_id_to_label = a.phonemizer.index_to_glabel
    a.group_label_to_id = {l: i for i, l in a.group_id_to_label.items()}
    a.phoneme_id_to_group_id = a.phonemizer.phoneme_id_to_group_id
    a._setup_config()
    a.silence_anchors = kw.get("silence_anchors", 10); a.boost_targets = True; a.enforce_minimum = True; a.enforce_all_targets = True
    a.ensure_completeness = kw.get("ensure_completeness", False); a.ignore_noise = True; a.extend_soft_boundaries = True; a.boundary_softness = 3
    a.bad_confidence_threshold = 0.6; a.break_at_low_confidence = False
    a._setup_decoders()
    a.reset_counters()
    return a


TEXTS = ["the quick brown fox, jumps over the lazy dog.", "hello world", "a stitch in time, saves nine. so they say, every day.",
         "one two three four five six seven eight nine ten eleven twelve", "short, text."]
out_json, out_npz = {}, {}
k = 0


def run(name, fn, texts, durs, **kw):
    global k
    a = reference_aligner(**kw)
    prov = synth.PlantedPosteriorProvider(a.phoneme_id_to_group_id, seed=1000 * (k + 1))
    a._cupe_prediction_batch = prov
    g = torch.Generator().manual_seed(77 + k)
    wavs = [torch.randn(1, int(d * SR), generator=g) * 0.1 for d in durs]
    # the provider needs each utterance's targets: process_segments phonemizes first, in order
    seqs = [a.phonemizer.phonemize_sentence(t.strip())["ph66"] for t in texts]
    orig = a.extract_timestamps_from_segment_batch
    def wrapped(w, wl, phs, **kws):
        prov.pending = [list(p) for p in phs]
        return orig(w, wl, phs, **kws)
    a.extract_timestamps_from_segment_batch = wrapped
    res = fn(a, texts, wavs)
    out_json[name] = {"texts": texts, "durs": durs, "kw": kw, "result": res, "n_calls": prov.n_calls, "provider_seed": prov.seed, "wav_seed": 77 + k,
                      "counters": [a.total_segments_processed, a.total_segments_failed, a.total_segments_bad, a.perfect_matches]}
    k += 1


run("sentence", lambda a, t, w: a.process_sentence(t[0], w[0], do_groups=True), TEXTS[:1], [4.0])
run("sentence_nogroups", lambda a, t, w: a.process_sentence(t[0], w[0], do_groups=False), TEXTS[2:3], [6.5], ensure_completeness=True)
run("batch", lambda a, t, w: a.process_sentences_batch(t, w, do_groups=True), TEXTS, [4.0, 1.2, 6.0, 5.0, 1.5])
run("batch_chunks", lambda a, t, w: a.process_segments([{"segments": [{"start": 0.0, "end": x.shape[1] / SR, "text": s}]} for s, x in zip(t, w)], w,
                                                       do_groups=True, batch_size=2), TEXTS, [4.0, 1.2, 6.0, 5.0, 1.5], silence_anchors=3)
run("too_short", lambda a, t, w: a.process_segments([{"segments": [{"start": 0.0, "end": x.shape[1] / SR, "text": s}]} for s, x in zip(t, w)], w,
                                                    do_groups=False, batch_size=1), [TEXTS[3], TEXTS[1]], [0.2, 1.0])

# ---- stich_window_predictions + log_softmax (cupe2i/windowing.py:103-173, core.py:898-899)
for j, (B, audio_len, fpw, Cc, wms, sms) in enumerate([(3, 48000, 10, 67, 120, 80), (2, 16000 * 7 + 123, 10, 17, 120, 80), (2, 32000, 7, 30, 160, 80),
                                                       (1, 4000, 10, 67, 120, 80)]):
    ws, ss = int(wms * SR / 1000), int(sms * SR / 1000)
    W = (audio_len - ws) // ss + 1
    x = torch.randn(B, W, fpw, Cc, generator=torch.Generator().manual_seed(500 + j)) * 3.0
    st = stich_window_predictions(x, original_audio_length=audio_len, cnn_output_size=fpw, sample_rate=SR, window_size_ms=wms, stride_ms=sms)
    out_npz[f"stitch{j}/x"] = x.numpy(); out_npz[f"stitch{j}/meta"] = np.asarray([audio_len, fpw, wms, sms], np.int64)
    out_npz[f"stitch{j}/stitched"] = st.numpy(); out_npz[f"stitch{j}/logp"] = torch.log_softmax(st, dim=2).numpy()

(HERE / "sentence.json").write_text(json.dumps(out_json))
np.savez_compressed(HERE / "stitch.npz", **out_npz)
print("wrote sentence.json / stitch.npz:", {n: v["n_calls"] for n, v in out_json.items()},
      "segments:", {n: (len(v["result"]["segments"]) if isinstance(v["result"], dict) else [len(r["segments"]) for r in v["result"]]) for n, v in out_json.items()})
