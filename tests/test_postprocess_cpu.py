"""CPU: the host epilogue (frames -> milliseconds) against fixtures from the unmodified reference's utils.convert_to_ms."""
from pathlib import Path

import numpy as np


def test_convert_to_ms_bit_identical_to_reference():
    from bfa_b200.postprocess import convert_to_ms, stamps_to_ms
    g = np.load(Path(__file__).parent / "golden" / "post.npz")
    cases = sorted({k.split("/")[0] for k in g.files})
    assert len(cases) == 12
    for c in cases:
        T, off, wav_len, sr, tl = g[f"{c}/args"]
        raw = g[f"{c}/stamps"]
        stamps = [tuple([int(r[0]), int(r[1]), int(r[2]), int(r[3]), bool(r[4]), float(r[5])][:int(tl)]) for r in raw]
        got = convert_to_ms(stamps, int(T), float(off), int(wav_len), int(sr))
        want = g[f"{c}/full"]
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert len(a) == 8
            assert [float(x) for x in a] == list(b)          # exact, including the defaults of short tuples
        if len(stamps):                                      # batch form on the BfaStamp layout
            arr = np.zeros((1, len(stamps) + 2, 4), np.int32)
            arr[0, :len(stamps), 0] = raw[:, 0]; arr[0, :len(stamps), 1] = raw[:, 1]; arr[0, :len(stamps), 2] = raw[:, 2]
            s_ms, e_ms = stamps_to_ms(arr, np.array([len(stamps)]), [T], [off], [wav_len], sr)
            np.testing.assert_array_equal(s_ms[0, :len(stamps)], g[f"{c}/ms"][:, 0])
            np.testing.assert_array_equal(e_ms[0, :len(stamps)], g[f"{c}/ms"][:, 1])
            assert (s_ms[0, len(stamps):] == 0).all()


def test_ensure_target_coverage_identical_to_reference():
    """core.py:462-679 on 160 seeded, damaged stamp lists (missing runs at the head / middle / tail, trailing silence, repeated
    and out-of-range target indices, completeness off): every repaired list equals the reference's."""
    from bfa_b200.postprocess import ensure_target_coverage
    g = np.load(Path(__file__).parent / "golden" / "coverage.npz")
    cases = sorted({k.split("/")[0] for k in g.files}, key=lambda c: int(c[1:]))
    assert len(cases) == 160
    inserted = 0
    for c in cases:
        tgt = [int(x) for x in g[f"{c}/tgt"]]
        inp = [tuple(int(x) for x in r) for r in g[f"{c}/in"]]
        raised, complete = (int(x) for x in g[f"{c}/meta"])
        try:
            got = ensure_target_coverage([tgt], [list(inp)], seq_lens=[len(tgt)], _silence_class=0, ensure_completeness=bool(complete))[0]
            assert not raised, c
        except Exception:
            assert raised, c
            continue
        want = [(int(r[0]), int(r[1]), int(r[2]), int(r[3]), bool(r[4])) for r in g[f"{c}/out"]]
        assert got == want, c
        inserted += sum(1 for s in got if s[4])
    assert inserted > 400


def test_align_words_identical_to_reference():
    """core.py:1062-1120 on 60 seeded lists (unsorted word indices, word_num shorter / longer than the timestamps, indices
    beyond the word list)."""
    import json
    from bfa_b200.postprocess import align_words
    cases = json.loads((Path(__file__).parent / "golden" / "words.json").read_text())
    assert len(cases) == 60
    for k, c in enumerate(cases):
        if c["err"]:
            try:
                align_words(c["phoneme_ts"], c["word_num"], c["words"])
                assert False, k
            except Exception:
                continue
        assert align_words(c["phoneme_ts"], c["word_num"], c["words"]) == c["out"], k


def test_post_process_segment_identical_to_reference():
    """core.py:1140-1210 (+ :1701-1732, :1062-1120): the per-segment output dict, compared as JSON with the reference's."""
    import json
    from bfa_b200.postprocess import post_process_segment
    cases = json.loads((Path(__file__).parent / "golden" / "segment.json").read_text())
    assert len(cases) == 40
    plabel = {i: f"ph{i}" for i in range(0, 60)}
    glabel = {i: f"g{i}" for i in range(0, 14)}
    for k, c in enumerate(cases):
        got = post_process_segment(dict(c["segment"]), c["ts"], list(c["seq"]), [tuple(p) for p in c["pts"]],
                                   None if c["gts"] is None else [tuple(g) for g in c["gts"]], index_to_plabel=plabel, index_to_glabel=glabel)
        assert json.loads(json.dumps(got)) == c["out"], k
