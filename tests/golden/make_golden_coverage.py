#!/usr/bin/env python
"""tests/golden/coverage.npz: PhonemeTimestampAligner.ensure_target_coverage (core.py:462-679) of the UNMODIFIED reference on
seeded, deliberately damaged stamp lists (build container only).  The method's source text is taken from core.py and run as a
plain function with a stand-in `self`; nothing of it is stored here."""
import textwrap, types
from pathlib import Path
import numpy as np

src = Path("/root/reference/bournemouth_aligner/core.py").read_text().split("\n")
i0 = next(i for i, l in enumerate(src) if l.strip().startswith("def ensure_target_coverage("))
i1 = next(i for i in range(i0 + 1, len(src)) if src[i].startswith("    def ") or src[i].startswith("class "))
ns = {}
exec(textwrap.dedent("\n".join(src[i0:i1])), ns)
ref_fn = ns["ensure_target_coverage"]


def fake_self(complete=True):
    s = types.SimpleNamespace(total_phonemes_aligned=0, total_phonemes_target=0, total_phonemes_extra=0, total_phonemes_missed=0,
                              total_phonemes_aligned_easily=0, warn_level=0, ensure_completeness=complete)
    s.phonemizer = types.SimpleNamespace(index_to_plabel={})
    return s


rng = np.random.default_rng(77)
out = {}
n_cases = 0
for case in range(160):
    N = int(rng.integers(1, 30))
    tgt = rng.integers(1, 60, N)
    tgt[rng.random(N) < 0.15] = 0                       # SIL class
    if case % 9 == 0:
        tgt[-int(rng.integers(1, min(N, 3) + 1)):] = 0  # trailing SIL
    T = int(rng.integers(N + 2, 6 * N + 20))
    cuts = np.sort(rng.choice(np.arange(1, T), size=min(2 * N, T - 1), replace=False))[: 2 * N]
    stamps = []
    for i in range(N):
        if 2 * i + 1 < len(cuts):
            s, e = int(cuts[2 * i]), int(cuts[2 * i + 1])
            if e > s:
                stamps.append((int(tgt[i]), s, e, i))
    mode = case % 8
    keep = np.ones(len(stamps), bool)
    if mode in (0, 1, 2, 3) and len(stamps) > 1:        # drop random stamps (missing targets), sometimes runs of them / the tail / the head
        k = int(rng.integers(1, max(2, len(stamps) // 2)))
        start = int(rng.integers(0, len(stamps))) if mode != 2 else max(0, len(stamps) - k)
        if mode == 3:
            start = 0
        idx = (np.arange(start, start + k) % len(stamps)) if mode == 1 else rng.choice(len(stamps), size=k, replace=False)
        if mode in (2, 3):
            idx = np.arange(start, min(start + k, len(stamps)))
        keep[idx] = False
    stamps = [s for s, kf in zip(stamps, keep) if kf]
    if mode in (4, 5) and stamps:                       # repeated targets: split a stamp (adjacent or with a gap)
        for _ in range(int(rng.integers(1, 3))):
            j = int(rng.integers(0, len(stamps)))
            ph, s, e, i = stamps[j]
            if e - s >= 3:
                m = s + int(rng.integers(1, e - s - 1))
                gap = int(rng.integers(0, 2)) if mode == 5 else 0
                stamps[j] = (ph, s, m, i)
                stamps.insert(j + 1, (ph, min(m + gap, e - 1), e, i))
    if mode == 6 and stamps:                            # invalid target indices
        j = int(rng.integers(0, len(stamps)))
        ph, s, e, i = stamps[j]
        stamps[j] = (ph, s, e, -1 if rng.random() < 0.5 else N + int(rng.integers(0, 3)))
    if mode == 7:
        stamps = [] if rng.random() < 0.3 else stamps[: max(1, len(stamps) // 3)]
    complete = (case % 11) != 10
    inp = [tuple(s) for s in stamps]
    try:
        got = ref_fn(fake_self(complete), [list(int(x) for x in tgt)], [list(inp)], seq_lens=[N], _silence_class=0)
        res, raised = got[0], 0
    except Exception:   # the reference's own completeness check (or an IndexError on odd indices)
        res, raised = [], 1
    c = f"c{n_cases}"
    out[f"{c}/tgt"] = np.asarray(tgt, np.int32)
    out[f"{c}/in"] = np.asarray(inp, np.int32).reshape(-1, 4)
    out[f"{c}/out"] = np.asarray([[r[0], r[1], r[2], r[3], int(r[4])] for r in res], np.int32).reshape(-1, 5)
    out[f"{c}/meta"] = np.asarray([raised, int(complete)], np.int32)
    n_cases += 1
np.savez_compressed(Path(__file__).resolve().parent / "coverage.npz", **out)
print("wrote coverage.npz:", n_cases, "cases,", sum(int(out[f"c{i}/meta"][0]) for i in range(n_cases)), "raise,",
      sum(int((out[f"c{i}/out"][:, 4] == 1).sum()) for i in range(n_cases)), "estimated stamps inserted")
