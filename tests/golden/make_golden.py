#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference module.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):      python tests/golden/make_golden.py

The reference package cannot be imported (phonemizer etc. are absent), but
forced_alignment.py and utils.py import only torch, so they are loaded by file
path.  Inputs are stored explicitly in the fixtures, so nothing at test time
depends on the reference or on torch's RNG.
"""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[2]
REF = Path("/root/reference/bournemouth_aligner")
sys.path.insert(0, str(REPO))
OUT = Path(__file__).resolve().parent


def load_ref():
    mods = {}
    for name in ("forced_alignment", "utils"):
        spec = importlib.util.spec_from_file_location(f"bfa_ref_{name}", REF / f"{name}.py")
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["forced_alignment"], mods["utils"]


def synth():
    spec = importlib.util.spec_from_file_location("bfa_synth", REPO / "bournemouth-forced-aligner_b200" / "synth.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def build_path(seq, stride, blank_id):
    N = len(seq)
    L = stride * N + 1
    path = torch.full((L,), blank_id, dtype=torch.long)
    tidx = torch.full((L,), -1, dtype=torch.long)
    path[1::stride] = seq
    tidx[1::stride] = torch.arange(N)
    return path, tidx


def capture_viterbi(dec, lp, path, tidx, band):
    """Run _viterbi_decode and capture dp/backpointers/path_states via sys.setprofile."""
    box = {}

    def prof(frame, event, arg):
        if event == "return" and frame.f_code.co_name == "_viterbi_decode":
            loc = frame.f_locals
            box["dp"] = loc["dp"].clone()
            box["bp"] = loc["backpointers"].clone()
            box["ps"] = loc["path_states"].clone()
            box["final_state"] = int(loc["final_state"])

    sys.setprofile(prof)
    try:
        fp, fi = dec._viterbi_decode(lp, path, len(path), tidx, band_width=band)
    finally:
        sys.setprofile(None)
    return fp, fi, box


def main():
    torch.set_num_threads(1)
    fa, ut = load_ref()
    S = synth()

    # ---------------------------------------------------------------- A: bare DP
    A = {}
    cases = []

    def add_core(name, lp, seq, stride, band, blank_id, truly_forced=True, keep_tables=False, path_override=None):
        dec = fa.ViterbiDecoder(blank_id, 0, silence_anchors=10, ignore_noise=True, truly_forced=truly_forced)
        if path_override is None:
            path, tidx = build_path(seq, stride, blank_id)
        else:
            path, tidx = path_override
        fp, fi, box = capture_viterbi(dec, lp, path, tidx, band)
        A[f"{name}/lp"] = lp.numpy().astype(np.float32)
        A[f"{name}/path"] = path.numpy().astype(np.int32)
        A[f"{name}/tidx"] = tidx.numpy().astype(np.int32)
        A[f"{name}/meta"] = np.array([band, blank_id, int(truly_forced), box["final_state"]], np.int32)
        A[f"{name}/frame_ph"] = fp.numpy().astype(np.int32)
        A[f"{name}/frame_idx"] = fi.numpy().astype(np.int32)
        A[f"{name}/ps"] = box["ps"].numpy().astype(np.int32)
        A[f"{name}/dp_last"] = box["dp"][-1].numpy().astype(np.float32)
        if keep_tables:
            A[f"{name}/dp"] = box["dp"].numpy().astype(np.float32)
            A[f"{name}/bp"] = box["bp"].numpy().astype(np.int32)
        cases.append(name)

    def band_fb(L):
        return max(L // 4, 20) if L > 60 else 0

    lp, tgt, _ = S.planted_batch(1, 60, 8, 67, seed=1)
    add_core("cfg1_T60_N8_C67", lp[0], tgt[0], 4, 0, 66, keep_tables=True)
    lp, tgt, _ = S.planted_batch(1, 600, 40, 66, seed=2)
    add_core("metric_T600_N40_C66", lp[0], tgt[0], 4, band_fb(161), 65)
    lp, tgt, _ = S.planted_batch(1, 600, 40, 67, seed=3)
    add_core("metric_T600_N40_C67_free_end", lp[0], tgt[0], 4, band_fb(161), 66, truly_forced=False)
    lp, tgt, _ = S.planted_batch(1, 100, 30, 66, seed=4, peak=10.0)
    add_core("stride3_T100_N30", lp[0], tgt[0], 3, band_fb(91), 65)
    lp, tgt, _ = S.planted_batch(1, 100, 45, 66, seed=5, peak=10.0)
    add_core("stride2_T100_N45", lp[0], tgt[0], 2, band_fb(91), 65)
    lp, tgt, _ = S.planted_batch(1, 80, 70, 66, seed=6, peak=10.0)
    add_core("stride1_T80_N70", lp[0], tgt[0], 1, band_fb(71), 65, keep_tables=True)
    # repeated phonemes: can_skip False on phoneme states at stride 2 / 1
    lp, tgt, _ = S.planted_batch(1, 64, 12, 17, seed=7, peak=9.0)
    rep = tgt[0].clone(); rep[1::2] = rep[0::2]
    add_core("repeat_stride2_C17", lp[0], rep, 2, 0, 16, keep_tables=True)
    rep1 = tgt[0].clone(); rep1[2:] = tgt[0][:-2]
    add_core("repeat_stride1_C17", lp[0], rep1, 1, 0, 16, keep_tables=True)
    # ties: coarse-quantised log-probs make equal candidates common (first-max order)
    g = torch.Generator().manual_seed(8)
    q = -torch.randint(0, 4, (50, 12), generator=g).float()
    add_core("ties_quantised", q, torch.tensor([1, 2, 3, 2, 1, 4]), 4, 0, 11, keep_tables=True)
    add_core("ties_quantised_free_end", q, torch.tensor([1, 2, 3, 2, 1, 4]), 2, 0, 11, truly_forced=False, keep_tables=True)
    # degenerate: flat random posteriors, every path below -1000 -> negative back-pointers wrap
    g = torch.Generator().manual_seed(9)
    flat = torch.log_softmax(torch.randn(600, 66, generator=g) * 3.0, dim=-1)
    add_core("degenerate_flat_T600", flat, torch.randint(1, 65, (40,), generator=g), 4, 40, 65)
    flat2 = torch.log_softmax(torch.randn(300, 20, generator=g) * 6.0, dim=-1)
    add_core("degenerate_flat_T300_small", flat2, torch.randint(1, 19, (12,), generator=g), 4, 0, 19, keep_tables=True)
    add_core("degenerate_free_end", flat2, torch.randint(1, 19, (20,), generator=g), 4, 20, 19, truly_forced=False)
    # band edge: L=61 just over the band threshold, T small so the band bites
    lp, tgt, _ = S.planted_batch(1, 64, 15, 66, seed=10, peak=10.0)
    add_core("band_L61", lp[0], tgt[0], 4, 20, 65, keep_tables=True)
    # band makes the last state unreachable -> truly_forced fallbacks (:673-682)
    lp, tgt, _ = S.planted_batch(1, 200, 16, 30, seed=11, peak=10.0)
    add_core("narrow_band_forced_fallback", lp[0], tgt[0], 4, 3, 29)
    # T == 1, T == 2
    lp, tgt, _ = S.planted_batch(1, 2, 1, 10, seed=12)
    add_core("T2_N1", lp[0], tgt[0], 1, 0, 9, keep_tables=True)
    add_core("T1_N1", lp[0][:1], tgt[0], 1, 0, 9, keep_tables=True)
    # long-form unsegmented shape (config 3 fallback): L=801, band 200
    lp, tgt, _ = S.planted_batch(1, 3600, 200, 66, seed=13, peak=12.0)
    add_core("long_T3600_N200", lp[0], tgt[0], 4, 200, 65)
    A["__cases__"] = np.array(cases)
    np.savez_compressed(OUT / "viterbi_core.npz", **A)
    print("viterbi_core:", len(cases), "cases")

    # ------------------------------------------------- B: decode_with_forced_alignment
    Bz = {}
    bcases = []

    def add_full(name, lp, seq, blank_id, silence_id=0, silence_anchors=10, truly_forced=True,
                 boost=True, floor=True):
        au = fa.AlignmentUtils(blank_id, silence_id, silence_anchors=silence_anchors, ignore_noise=True,
                               truly_forced=truly_forced)
        dec = au.viterbi_decoder
        err = 0
        segmented = -1
        try:
            fp, fi, _ = dec.decode_with_forced_alignment(lp, seq, boost_targets=boost, enforce_minimum=floor,
                                                         anchor_pauses=silence_anchors > 0)
            # was the segmented branch taken?
            m = lp.clone()
            if boost: m = dec._boost_target_phonemes(m, seq)
            if floor: m = dec._enforce_minimum_probabilities(m, seq)
            if silence_anchors > 0 and len(seq) > 0:
                r = dec._segmented_viterbi_decode(m, seq, torch.arange(len(seq)))
                segmented = int(r is not None and len(r[0]) == lp.shape[0])
            if lp.shape[0] <= 300:
                Bz[f"{name}/modified"] = m.numpy().astype(np.float32)
            stamps = dec.assort_frames(fp, fi)
            conf = ut._calculate_confidences(lp, [s + (False,) for s in stamps])
            Bz[f"{name}/frame_ph"] = fp.numpy().astype(np.int32)
            Bz[f"{name}/frame_idx"] = fi.numpy().astype(np.int32)
            Bz[f"{name}/stamps"] = np.array(stamps, np.int32).reshape(-1, 4)
            Bz[f"{name}/conf"] = np.array([c[5] for c in conf], np.float32)
        except ValueError:
            err = 1
        Bz[f"{name}/lp"] = lp.numpy().astype(np.float32)
        Bz[f"{name}/seq"] = seq.numpy().astype(np.int32)
        Bz[f"{name}/meta"] = np.array([blank_id, -1 if silence_id is None else silence_id, silence_anchors,
                                       int(truly_forced), int(boost), int(floor), err, segmented], np.int32)
        bcases.append(name)
        print(f"  {name}: T={lp.shape[0]} N={len(seq)} err={err} segmented={segmented}")

    lp, tgt, _ = S.planted_batch(1, 60, 8, 67, seed=21)
    add_full("cfg1", lp[0], tgt[0], 66)
    lp, tgt, _ = S.planted_batch(1, 600, 40, 66, seed=22)
    add_full("metric", lp[0], tgt[0], 65)
    add_full("metric_noboost", lp[0], tgt[0], 65, boost=False, floor=False)
    add_full("metric_noanchor", lp[0], tgt[0], 65, silence_anchors=0)
    lp, tgt, _ = S.planted_batch(1, 900, 60, 67, seed=23, peak=12.0, sil_every=12, sil_frames=22)
    add_full("sil_segmented_T900", lp[0], tgt[0], 66)
    add_full("sil_segmented_T900_free", lp[0], tgt[0], 66, truly_forced=False)
    lp, tgt, _ = S.planted_batch(1, 400, 60, 67, seed=24, peak=12.0, sil_every=10, sil_frames=14)
    add_full("sil_dense_stride_lt4", lp[0], tgt[0], 66)
    lp, tgt, _ = S.planted_batch(1, 420, 100, 67, seed=25, peak=12.0, sil_every=9, sil_frames=12)
    add_full("sil_bailout_candidate", lp[0], tgt[0], 66)
    lp, tgt, _ = S.planted_batch(1, 300, 24, 17, seed=26, peak=10.0, sil_every=6, sil_frames=16, blank_id=16)
    add_full("sil_groups_C17", lp[0], tgt[0], 16)
    # target SILs present but no audio silence -> segmentation returns [] -> fallback
    lp, tgt, _ = S.planted_batch(1, 300, 30, 67, seed=27, peak=10.0)
    t2 = tgt[0].clone(); t2[10] = 0; t2[20] = 0
    add_full("sil_in_target_only", lp[0], t2, 66)
    # double SIL tokens (". ," style) and leading/trailing SIL
    lp, tgt, _ = S.planted_batch(1, 700, 40, 67, seed=28, peak=12.0, sil_every=8, sil_frames=25)
    t3 = tgt[0].clone(); t3[17] = 0  # adjacent to the SIL at 16 -> group of 2
    add_full("sil_double_group", lp[0], t3, 66)
    lp, tgt, _ = S.planted_batch(1, 40, 40, 67, seed=29, peak=10.0)
    add_full("T_eq_N_proportional", lp[0], tgt[0], 66)
    add_full("T_lt_N_raises", lp[0][:30], tgt[0], 66)
    lp, tgt, _ = S.planted_batch(1, 50, 40, 67, seed=30, peak=10.0)
    add_full("stride1_fallback", lp[0], tgt[0], 66)
    lp, tgt, _ = S.planted_batch(1, 130, 40, 67, seed=31, peak=10.0)
    add_full("stride3_fallback", lp[0], tgt[0], 66)
    lp, tgt, _ = S.planted_batch(1, 90, 40, 67, seed=32, peak=10.0)
    add_full("stride2_fallback", lp[0], tgt[0], 66)
    add_full("empty_target", lp[0], tgt[0][:0], 66)
    # >200 phonemes with weak silences -> lowered-threshold retries (:298-308)
    lp, tgt, _ = S.planted_batch(1, 2400, 220, 67, seed=33, peak=3.0, sil_every=40, sil_frames=20)
    add_full("long_retry_thresholds", lp[0], tgt[0], 66)
    lp, tgt, _ = S.planted_batch(1, 3600, 200, 66, seed=34, peak=12.0, sil_every=40, sil_frames=18)
    add_full("cfg3_long_T3600_N200", lp[0], tgt[0], 65)
    Bz["__cases__"] = np.array(bcases)
    np.savez_compressed(OUT / "decode_forced.npz", **Bz)
    print("decode_forced:", len(bcases), "cases")

    # ------------------------------------------------- C: batch API + simple + silence scan
    Cz = {}
    lp, tgt, _ = S.planted_batch(6, 200, 20, 67, seed=41, peak=10.0, sil_every=7, sil_frames=14)
    pred_lens = torch.tensor([200, 180, 150, 200, 90, 120])
    seq_lens = torch.tensor([20, 18, 0, 12, 20, 5])
    au = fa.AlignmentUtils(66, 0, silence_anchors=10, ignore_noise=True, truly_forced=True)
    res = au.decode_alignments(lp, true_seqs=tgt, pred_lens=pred_lens, true_seqs_lens=seq_lens)
    Cz["batch/lp"] = lp.numpy(); Cz["batch/tgt"] = tgt.numpy().astype(np.int32)
    Cz["batch/pred_lens"] = pred_lens.numpy().astype(np.int32); Cz["batch/seq_lens"] = seq_lens.numpy().astype(np.int32)
    for i, r in enumerate(res):
        Cz[f"batch/stamps{i}"] = np.array(r, np.int32).reshape(-1, 4)
        conf = ut._calculate_confidences(lp[i], [s + (False,) for s in r])
        Cz[f"batch/conf{i}"] = np.array([c[5] for c in conf], np.float32)
    res = au.decode_alignments_simple(lp, tgt, pred_lens=pred_lens, true_seqs_lens=torch.tensor([20, 18, 1, 12, 20, 5]))
    for i, r in enumerate(res):
        Cz[f"simple/stamps{i}"] = np.array(r, np.int32).reshape(-1, 4)
    Cz["simple/seq_lens"] = np.array([20, 18, 1, 12, 20, 5], np.int32)
    au_noise = fa.AlignmentUtils(66, 0, silence_anchors=0, ignore_noise=False, truly_forced=True)
    res = au_noise.decode_alignments(lp, true_seqs=tgt, pred_lens=pred_lens, true_seqs_lens=torch.tensor([20, 18, 1, 3, 4, 5]))
    for i, r in enumerate(res):
        Cz[f"noise/stamps{i}"] = np.array(r, np.int32).reshape(-1, 4)
    Cz["noise/seq_lens"] = np.array([20, 18, 1, 3, 4, 5], np.int32)
    # group head C=17, blank 16
    lpg, tgg, _ = S.planted_batch(3, 150, 14, 17, seed=42, peak=9.0, blank_id=16)
    aug = fa.AlignmentUtils(16, 0, silence_anchors=10, ignore_noise=True, truly_forced=True)
    res = aug.decode_alignments(lpg, true_seqs=tgg, pred_lens=torch.tensor([150, 140, 100]), true_seqs_lens=torch.tensor([14, 14, 9]))
    Cz["group/lp"] = lpg.numpy(); Cz["group/tgt"] = tgg.numpy().astype(np.int32)
    for i, r in enumerate(res):
        Cz[f"group/stamps{i}"] = np.array(r, np.int32).reshape(-1, 4)
    # silence scan on its own (several k / thresholds)
    dec = au.viterbi_decoder
    lps, _, _ = S.planted_batch(1, 500, 30, 67, seed=43, peak=9.0, sil_every=5, sil_frames=13)
    Cz["silscan/lp"] = lps[0].numpy()
    for j, (thr, k) in enumerate([(0.9, 10), (0.8, 10), (0.9, 3), (0.5, 1), (0.1, 12), (0.9, 600)]):
        segs = dec._detect_silence_segments(lps[0], sil_prob_threshold=thr, min_silence_frames=k)
        Cz[f"silscan/segs{j}"] = np.array(segs, np.int32).reshape(-1, 2)
        Cz[f"silscan/args{j}"] = np.array([thr, k], np.float64)
    # torch constants the port must reproduce
    Cz["const/min_log_prob"] = torch.log(torch.tensor(1e-8)).numpy()
    np.savez_compressed(OUT / "batch_api.npz", **Cz)
    print("batch_api written")


if __name__ == "__main__":
    main()
