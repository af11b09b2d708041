#!/usr/bin/env python
"""tests/golden/words.json: PhonemeTimestampAligner._align_words (core.py:1062-1120) of the UNMODIFIED reference on seeded
phoneme timestamp lists (build container only; the method's source text is executed as a plain function)."""
import json, textwrap
from pathlib import Path
import numpy as np

src = Path("/root/reference/bournemouth_aligner/core.py").read_text().split("\n")
i0 = next(i for i, l in enumerate(src) if l.strip().startswith("def _align_words("))
i1 = next(i for i in range(i0 + 1, len(src)) if src[i].startswith("    def ") or src[i].startswith("class "))
ns = {}
exec(textwrap.dedent("\n".join(src[i0:i1])), ns)
fn = ns["_align_words"]
rng = np.random.default_rng(3)
cases = []
for c in range(60):
    n = int(rng.integers(0, 25))
    n_words = int(rng.integers(1, 8))
    wn = sorted(int(x) for x in rng.integers(0, n_words, n)) if c % 5 else [int(x) for x in rng.integers(0, n_words, n)]
    if c % 7 == 0 and n > 2:
        wn = wn[: n - int(rng.integers(1, 3))]        # word_num shorter than the timestamps
    if c % 11 == 0:
        wn = wn + [n_words + 1]                       # longer, and beyond the word list
    t = np.cumsum(rng.random(n + 1) * 80.0)
    ph = [{"phoneme_id": int(rng.integers(1, 66)), "phoneme_label": f"p{int(rng.integers(0, 66))}", "ipa_label": f"i{int(rng.integers(0, 66))}",
           "start_ms": float(t[i]), "end_ms": float(t[i + 1]), "confidence": float(rng.random())} for i in range(n)]
    words = [f"w{k}" for k in range(n_words if c % 4 else max(1, n_words - 2))]
    try:
        out, err = fn(None, ph, wn, words), None
    except Exception as e:   # the reference's own failure modes (e.g. empty word at the very first index)
        out, err = None, type(e).__name__
    cases.append({"phoneme_ts": ph, "word_num": wn, "words": words, "out": out, "err": err})
Path(__file__).resolve().parent.joinpath("words.json").write_text(json.dumps(cases))
print("wrote words.json:", len(cases), "cases,", sum(1 for c in cases if c["err"]), "raise,", sum(len(c["out"] or []) for c in cases), "words")
