#!/usr/bin/env python
"""Fuzz the C oracle against the live reference module (build container only: needs
/root/reference).  TEST INFRASTRUCTURE.  Usage: python oracle/validate_against_reference.py [n_cases]

Reports, per branch, how many random cases were compared and how many differed.
"""
import importlib.util
import sys
from collections import Counter
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path[:] = [q for q in sys.path if Path(q or ".").resolve() != REPO / "oracle"]
sys.path.insert(0, str(REPO))
from oracle import oracle as orc  # noqa: E402

sys.path.insert(0, str(REPO / "tests" / "golden"))
import make_golden as mg  # noqa: E402


def main(n_cases=300, seed0=1000):
    torch.set_num_threads(1)
    fa, ut = mg.load_ref()
    S = mg.synth()
    rng = np.random.default_rng(seed0)
    tally, bad = Counter(), []
    for i in range(n_cases):
        C = int(rng.choice([17, 30, 66, 67]))
        blank = C - 1
        T = int(rng.integers(20, 700))
        dens = rng.choice([0.05, 0.1, 0.2, 0.3, 0.5, 0.9, 1.0])
        N = max(1, min(int(T * dens), 250))
        sil_every = int(rng.choice([0, 0, 3, 5, 8, 12]))
        sil_frames = int(rng.choice([4, 11, 14, 25]))
        peak = float(rng.choice([3.0, 6.0, 9.0, 12.0]))
        anchors = int(rng.choice([0, 3, 10, 10]))
        forced = bool(rng.integers(0, 2))
        lp, tgt, _ = S.planted_batch(1, T, N, C, seed=seed0 + i, peak=peak, sil_every=sil_every, sil_frames=sil_frames)
        lp, seq = lp[0], tgt[0]
        if rng.random() < 0.15:   # flat / degenerate
            lp = torch.log_softmax(torch.randn(T, C) * 4, -1)
        au = fa.AlignmentUtils(blank, 0, silence_anchors=anchors, ignore_noise=bool(rng.integers(0, 2)), truly_forced=forced)
        p = orc.params(blank, 0, anchors, au.viterbi_decoder.ignore_noise, forced)
        try:
            fp, fi, _ = au.viterbi_decoder.decode_with_forced_alignment(lp, seq, anchor_pauses=anchors > 0)
            err = False
        except ValueError:
            err = True
        r = orc.decode_forced(lp.numpy(), seq.numpy(), p)
        st = r["status"] & 7
        key = {0: "fallback", 2: "too_short", 3: "proportional", 4: "segmented"}[st] + ("+degenerate" if r["status"] & 8 else "")
        tally[key] += 1
        if err != (st == orc.ORC_TOO_SHORT):
            bad.append((i, "error mismatch")); continue
        if err:
            continue
        if not (np.array_equal(r["frame_ph"], fp.numpy()) and np.array_equal(r["frame_idx"], fi.numpy())):
            nd = int((r["frame_ph"] != fp.numpy()).sum())
            bad.append((i, f"{key}: {nd} frames differ T={T} N={N} C={C} anchors={anchors}")); continue
        stamps_ref = au.viterbi_decoder.assort_frames(fp, fi)
        stamps = orc.assort(r["frame_ph"], r["frame_idx"], blank, au.viterbi_decoder.ignore_noise)
        if [tuple(s) for s in stamps_ref] != stamps:
            bad.append((i, "assort")); continue
        if stamps:
            c_ref = np.array([c[5] for c in ut._calculate_confidences(lp, [s + (False,) for s in stamps_ref])], np.float32)
            if not np.allclose(orc.confidence(lp.numpy(), stamps), c_ref, rtol=1e-5, atol=1e-7):
                bad.append((i, "confidence"))
    print("branches:", dict(tally))
    print("mismatches:", len(bad))
    for b in bad[:20]:
        print("  ", b)
    return len(bad)


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 300) else 0)
