#!/usr/bin/env python
"""tests/golden/pipeline.npz: the post-acoustic part of extract_timestamps_from_segment_batch (core.py:902-937 + :960-975) run with the
UNMODIFIED reference pieces -- AlignmentUtils.decode_alignments, ensure_target_coverage, extend_soft_boundaries_func,
utils._calculate_confidences, utils.convert_to_ms -- on seeded posteriors (build container only)."""
import importlib.util, sys, textwrap, types
from pathlib import Path
import numpy as np
import torch

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
REF = Path("/root/reference/bournemouth_aligner")


def load(name):
    spec = importlib.util.spec_from_file_location(f"bfa_ref_{name}", REF / f"{name}.py")
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); return m


fa, ut = load("forced_alignment"), load("utils")
src = (REF / "core.py").read_text().split("\n")


def method(name):
    i0 = next(i for i, l in enumerate(src) if l.strip().startswith(f"def {name}("))
    i1 = next(i for i in range(i0 + 1, len(src)) if src[i].startswith("    def ") or src[i].startswith("class "))
    ns = {"torch": torch}
    exec(textwrap.dedent("\n".join(src[i0:i1])), ns)
    return ns[name]


cover, soften = method("ensure_target_coverage"), method("extend_soft_boundaries_func")
me = types.SimpleNamespace(total_phonemes_aligned=0, total_phonemes_target=0, total_phonemes_extra=0, total_phonemes_missed=0,
                           total_phonemes_aligned_easily=0, warn_level=0, ensure_completeness=True,
                           phonemizer=types.SimpleNamespace(index_to_plabel={}))
from bfa_b200 import synth   # noqa: E402

out = {}
k = 0
for (B, T, N, Cc, peak, every, seed) in [(4, 300, 24, 67, 6.0, 0, 1), (3, 500, 40, 67, 5.0, 9, 2), (4, 240, 30, 67, 3.5, 0, 3)]:
    lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=900 + seed, peak=peak, sil_every=every, sil_frames=14)
    lens = torch.randint(T - 60, T + 1, (B,), generator=torch.Generator().manual_seed(seed))
    nl = torch.full((B,), N, dtype=torch.long)
    au = fa.AlignmentUtils(blank_id=Cc - 1, silence_id=0, silence_anchors=10, ignore_noise=True, truly_forced=True)
    frames = au.decode_alignments(lp, true_seqs=tgt, pred_lens=lens, true_seqs_lens=nl, forced_alignment=True, boost_targets=True,
                                  enforce_minimum=True)
    frames = cover(me, tgt, frames, seq_lens=nl, _silence_class=0)
    frames = soften(me, lp, frames, boundary_softness=3)
    for b in range(B):
        fs = ut._calculate_confidences(lp[b], frames[b])
        fs = ut.convert_to_ms(fs, int(lens[b]), 1.5 * b, int(lens[b]) * 320, 16000)
        out[f"u{k}/lp"] = lp[b].numpy(); out[f"u{k}/tgt"] = tgt[b].numpy().astype(np.int32)
        out[f"u{k}/meta"] = np.array([int(lens[b]), N, Cc, b], np.int32)
        out[f"u{k}/out"] = np.array([[float(x) for x in f] for f in fs], np.float64).reshape(-1, 8)
        k += 1
np.savez_compressed(Path(__file__).resolve().parent / "pipeline.npz", **out)
print("wrote pipeline.npz:", k, "utterances,", sum(len(out[f"u{i}/out"]) for i in range(k)), "stamps,",
      sum(int(out[f"u{i}/out"][:, 4].sum()) for i in range(k)), "estimated")
