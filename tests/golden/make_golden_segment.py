#!/usr/bin/env python
"""tests/golden/segment.json: PhonemeTimestampAligner.post_process_segment (core.py:1140-1210) with its helpers
analyze_alignment_coverage (:1701-1732) and _align_words (:1062-1120) of the UNMODIFIED reference, on seeded inputs (build
container only; the methods' source text is executed as plain functions bound to a stand-in object)."""
import json, textwrap, types
from pathlib import Path
import numpy as np

src = Path("/root/reference/bournemouth_aligner/core.py").read_text().split("\n")


def method(name):
    i0 = next(i for i, l in enumerate(src) if l.strip().startswith(f"def {name}("))
    i1 = next(i for i in range(i0 + 1, len(src)) if src[i].startswith("    def ") or src[i].startswith("class "))
    ns = {}
    exec(textwrap.dedent("\n".join(src[i0:i1])), ns)
    return ns[name]


me = types.SimpleNamespace()
plabel = {i: f"ph{i}" for i in range(0, 60)}          # 60..65 have no label -> UNK_
glabel = {i: f"g{i}" for i in range(0, 14)}
me.phonemizer = types.SimpleNamespace(index_to_plabel=plabel, index_to_glabel=glabel)
for n in ("analyze_alignment_coverage", "_align_words", "post_process_segment"):
    setattr(me, n, types.MethodType(method(n), me))
rng = np.random.default_rng(11)
cases = []
for c in range(40):
    N = int(rng.integers(1, 30))
    seq = [int(x) for x in rng.integers(0, 66, N)]
    n_words = int(rng.integers(1, 7))
    word_num = sorted(int(x) for x in rng.integers(0, n_words, N))
    ts = {"eipa": [f"e{k}" for k in range(N if c % 6 else max(0, N - 2))], "word_num": word_num, "words": [f"w{k}" for k in range(n_words)]}
    if c % 9 == 0:
        ts = {}
    t = np.cumsum(rng.random(N + 1) * 60.0)
    keep = [i for i in range(N) if rng.random() > 0.1 or c % 3]
    pts = [(seq[i] if rng.random() > 0.05 else int(rng.integers(0, 70)), int(t[i] / 20), int(t[i + 1] / 20) + 1, i if rng.random() > 0.03 else N + 1,
            bool(rng.random() < 0.1), float(rng.random()), float(t[i]), float(t[i + 1])) for i in keep]
    gts = None if c % 2 else [(int(rng.integers(0, 17)), p[1], p[2], p[3], p[4], float(rng.random()), p[6], p[7]) for p in pts]
    segment = {"start": float(c), "end": float(c) + 3.0, "text": f"text {c}"}
    out = me.post_process_segment(dict(segment), ts, list(seq), [tuple(p) for p in pts], None if gts is None else [tuple(g) for g in gts])
    cases.append({"segment": segment, "ts": ts, "seq": seq, "pts": pts, "gts": gts, "out": out})
Path(__file__).resolve().parent.joinpath("segment.json").write_text(json.dumps(cases))
print("wrote segment.json:", len(cases), "cases")
