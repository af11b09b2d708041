"""The reference-shaped entry points above the alignment: PhonemeTimestampAligner.process_sentence / process_sentences_batch /
process_segments with the acoustic model and the phonemizer injected (core.py:1212-1616), the fused
stich_window_predictions + log_softmax kernel (cupe2i/windowing.py:103-173, core.py:898-899) and the alignment score
(forced_alignment.py:767-773).  Expected values come from the UNMODIFIED reference class run in the build container with the
same stand-ins (tests/golden/make_golden_sentence.py)."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
SR = 16000


def _same(got, want, path="result"):
    """Deep comparison: structure, ints, strings and booleans exact; times to 1e-6 relative; confidences to 1e-4."""
    if isinstance(want, dict):
        assert isinstance(got, dict) and sorted(got) == sorted(want), f"{path}: keys {sorted(got)} vs {sorted(want)}"
        for k in want:
            _same(got[k], want[k], f"{path}.{k}")
    elif isinstance(want, (list, tuple)):
        assert len(got) == len(want), f"{path}: {len(got)} vs {len(want)} entries"
        for i, (g, w) in enumerate(zip(got, want)):
            _same(g, w, f"{path}[{i}]")
    elif isinstance(want, bool) or isinstance(want, (int, str)) or want is None:
        assert got == want, f"{path}: {got!r} vs {want!r}"
    else:
        tol = 1e-4 if path.endswith("confidence") else 1e-6
        assert abs(float(got) - float(want)) <= tol * max(1.0, abs(float(want))), f"{path}: {got} vs {want}"


def _aligner(bfa, case):
    from bfa_b200 import synth
    ph = synth.FakePhonemizer()
    prov = synth.PlantedPosteriorProvider(ph.phoneme_id_to_group_id, seed=case["provider_seed"])
    kw = dict(silence_anchors=case["kw"].get("silence_anchors", 10), ensure_completeness=case["kw"].get("ensure_completeness", False))
    a = bfa.PhonemeTimestampAligner(posterior_provider=prov, phonemizer=ph, **kw)
    inner = a.extract_timestamps_from_segment_batch

    def wrapped(w, wl, phs, **kws):          # the stand-in acoustic model plants each utterance's own targets
        prov.pending = [list(p) for p in phs]
        return inner(w, wl, phs, **kws)
    a.extract_timestamps_from_segment_batch = wrapped
    g = torch.Generator().manual_seed(case["wav_seed"])
    wavs = [torch.randn(1, int(d * SR), generator=g) * 0.1 for d in case["durs"]]
    return a, prov, wavs


def _segments_call(a, t, w, **kw):
    return a.process_segments([{"segments": [{"start": 0.0, "end": x.shape[1] / SR, "text": s}]} for s, x in zip(t, w)], w, **kw)


CALLS = {
    "sentence": lambda a, t, w: a.process_sentence(t[0], w[0], do_groups=True),
    "sentence_nogroups": lambda a, t, w: a.process_sentence(t[0], w[0], do_groups=False),
    "batch": lambda a, t, w: a.process_sentences_batch(t, w, do_groups=True),
    "batch_chunks": lambda a, t, w: _segments_call(a, t, w, do_groups=True, batch_size=2),
    "too_short": lambda a, t, w: _segments_call(a, t, w, do_groups=False, batch_size=1),
}


@pytest.mark.parametrize("name", sorted(CALLS))
def test_process_sentence_api_vs_reference(bfa, dev, name):
    case = json.loads((GOLD / "sentence.json").read_text())[name]
    a, prov, wavs = _aligner(bfa, case)
    got = CALLS[name](a, case["texts"], wavs)
    assert prov.n_calls == case["n_calls"]
    _same(json.loads(json.dumps(got)), case["result"])
    assert [a.total_segments_processed, a.total_segments_failed, a.total_segments_bad, a.perfect_matches] == case["counters"]


def test_process_sentence_takes_the_logits_path(bfa, dev):
    """Text without punctuation (no silence_id in the targets): both heads are aligned straight from the provider's logits (one
    kernel each, core.py:898-899 never run; soft boundaries and confidences from logits + row_lse).  Same record as when the
    provider's logits are normalised first and everything runs on log-probabilities."""
    from bfa_b200 import synth
    import bfa_b200.pipeline as pl
    case = json.loads((GOLD / "sentence.json").read_text())["sentence"]
    text = "a plain sentence without any pause marks in it at all"
    a, prov, wavs = _aligner(bfa, case)
    got = a.process_sentence(text, wavs[0], do_groups=True)
    assert a.alignment_utils_p.last_row_lse is not None and a.alignment_utils_g.last_row_lse is not None
    assert a.alignment_utils_p.last_log_probs is None
    b, prov_b, _ = _aligner(bfa, case)
    keep = b._log_posteriors
    b._log_posteriors = lambda w, wl, e, keep_logits=False: keep(w, wl, e, keep_logits=False)      # normalise first, like the reference
    want = b.process_sentence(text, wavs[0], do_groups=True)
    assert b.alignment_utils_p.last_row_lse is None
    _same(json.loads(json.dumps(got)), json.loads(json.dumps(want)))
    assert len(got["segments"][0]["phoneme_ts"]) > 10


def test_process_segments_argument_errors(bfa, dev):
    from bfa_b200 import synth
    a = bfa.PhonemeTimestampAligner(posterior_provider=lambda w, l: None, phonemizer=synth.FakePhonemizer())
    with pytest.raises(ValueError, match="Batch size mismatch"):
        a.process_segments([{"segments": []}], [torch.zeros(1, 100), torch.zeros(1, 100)])
    with pytest.raises(ValueError, match="missing 'segments' key"):
        a.process_segments([{"x": 1}], [torch.zeros(1, 100)])
    with pytest.raises(ValueError, match="missing required keys"):
        a.process_segments([{"segments": [{"start": 0.0, "text": "a"}]}], [torch.zeros(1, 100)])
    with pytest.raises(ValueError, match="Expected audio_wavs"):
        a.process_segments([{"segments": []}], torch.zeros(100))
    assert a.process_segments([{"segments": []}], [torch.zeros(1, 100)]) == [{"segments": []}]
    with pytest.raises(ValueError, match="chopping errors"):      # 10 samples: shorter than seg_duration_min
        a.process_segments([{"segments": [{"start": 0.0, "end": 10 / SR, "text": "hello"}]}], [torch.zeros(1, 10)])


def test_stitch_log_softmax_vs_reference(bfa, dev):
    g = np.load(GOLD / "stitch.npz")
    cases = sorted({k.split("/")[0] for k in g.files})
    assert len(cases) == 4
    for c in cases:
        audio_len, fpw, wms, sms = (int(v) for v in g[f"{c}/meta"])
        x = torch.from_numpy(g[f"{c}/x"]).to(dev)
        got = bfa.stitch_log_softmax(x, original_audio_length=audio_len, cnn_output_size=fpw, sample_rate=SR, window_size_ms=wms, stride_ms=sms)
        want = g[f"{c}/logp"]
        assert tuple(got.shape) == want.shape, c
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=2e-6, err_msg=c)
        # the cross-fade alone: re-normalising the reference's stitched tensor with the same kernel (only the cos window, which
        # the caller's device computes, separates the two)
        again = bfa.log_softmax_rows(torch.from_numpy(g[f"{c}/stitched"]).to(dev))
        np.testing.assert_allclose(again.cpu().numpy(), got.cpu().numpy(), rtol=1e-5, atol=2e-6, err_msg=c)


def test_log_softmax_rows_vs_torch(bfa, dev):
    for (B, T, Cc) in [(3, 50, 67), (2, 33, 17), (1, 7, 200), (4, 1, 1)]:
        x = (torch.randn(B, T, Cc, generator=torch.Generator().manual_seed(B * T + Cc)) * 4.0).to(dev)
        got = bfa.log_softmax_rows(x)
        np.testing.assert_allclose(got.cpu().numpy(), torch.log_softmax(x.cpu(), dim=2).numpy(), rtol=1e-5, atol=2e-6)


def test_alignment_score_on_device(bfa, orc, dev):
    """return_scores=True (forced_alignment.py:195-197, :767-773): the score is the sum of the ORIGINAL log-probs along the path."""
    from bfa_b200 import synth
    T, N, Cc = 180, 20, 67
    lp, tgt, _ = synth.planted_batch(1, T, N, Cc, seed=77, peak=8.0)
    dec = bfa.AlignmentUtils(Cc - 1, 0).viterbi_decoder
    fp, fi, score = dec.decode_with_forced_alignment(lp[0].to(dev), tgt[0], return_scores=True)
    want = float(sum(float(lp[0, t, int(p)]) for t, p in enumerate(fp.cpu().tolist())))
    assert abs(score - want) <= 1e-9 * max(1.0, abs(want))
    assert dec._calculate_alignment_score(lp[0].to(dev), torch.full((T,), Cc + 5)) == 0.0       # labels >= C are skipped (:771)
