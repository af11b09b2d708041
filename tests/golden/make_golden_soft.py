#!/usr/bin/env python
"""tests/golden/soft.npz: PhonemeTimestampAligner.extend_soft_boundaries_func (core.py:682-809) of the UNMODIFIED reference,
run on seeded posteriors + stamps (build container only).  core.py itself cannot be imported (phonemizer, librosa, ... are
absent), so the method's source text is taken from the file and executed as a plain function; nothing of it is stored here."""
import re, sys, textwrap
from pathlib import Path
import numpy as np
import torch

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
src = Path("/root/reference/bournemouth_aligner/core.py").read_text().split("\n")
i0 = next(i for i, l in enumerate(src) if l.strip().startswith("def extend_soft_boundaries_func("))
i1 = next(i for i in range(i0 + 1, len(src)) if src[i].startswith("    def ") or src[i].startswith("class "))
ns = {"torch": torch}
exec(textwrap.dedent("\n".join(src[i0:i1])), ns)
ref_fn = ns["extend_soft_boundaries_func"]

from bfa_b200 import synth           # noqa: E402  (pure torch, no CUDA needed)
from oracle import oracle as orc     # noqa: E402

out = {}
case = 0
for (B, T, N, Cc, peak, soft, seed) in [(4, 300, 20, 67, 6.0, 3, 1), (3, 600, 40, 66, 8.0, 3, 2), (2, 200, 30, 30, 4.0, 5, 3),
                                         (3, 150, 12, 17, 5.0, 2, 4), (2, 900, 60, 67, 7.0, 7, 5), (1, 60, 8, 67, 3.0, 3, 6)]:
    lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=800 + seed, peak=peak)
    p = orc.params(Cc - 1, 0)
    o = orc.align_batch(p, lp.numpy(), np.arange(B, dtype=np.int64) * T * Cc, np.full(B, T, np.int32), Cc,
                        tgt.numpy().astype(np.int32).reshape(-1), np.arange(B + 1, dtype=np.int64) * N, max_stamps=T)
    stamps = []
    for b in range(B):
        n = int(o["n_stamps"][b])
        stamps.append([(int(o["stamps"][b]["phoneme"][i]), int(o["stamps"][b]["start"][i]), int(o["stamps"][b]["end"][i]),
                        int(o["stamps"][b]["target_idx"][i]), False) for i in range(n)])
    if seed == 4:      # shrink some stamps to single frames and open gaps, like estimated insertions do
        stamps = [[(s[0], s[1], min(s[1] + 1, s[2]), s[3], True) if i % 3 == 0 else s for i, s in enumerate(st)] for st in stamps]
    got = ref_fn(None, lp, [list(s) for s in stamps], boundary_softness=soft)
    for b in range(B):
        out[f"c{case}/lp"] = lp[b].numpy()
        out[f"c{case}/in"] = np.array([[s[0], s[1], s[2], s[3], int(s[4])] for s in stamps[b]], np.int32).reshape(-1, 5)
        out[f"c{case}/out"] = np.array([[s[0], s[1], s[2], s[3], int(s[4])] for s in got[b]], np.int32).reshape(-1, 5)
        out[f"c{case}/soft"] = np.array([soft], np.int32)
        case += 1
np.savez_compressed(Path(__file__).resolve().parent / "soft.npz", **out)
changed = sum(int((out[f"c{c}/in"][:, 1:3] != out[f"c{c}/out"][:, 1:3]).any(1).sum()) for c in range(case))
print("wrote soft.npz:", case, "utterances,", changed, "stamps changed by the reference")
