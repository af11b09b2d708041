"""GPU parity tests: the CUDA path (through the C-ABI) against the golden vectors from the
reference and against the C oracle on seeded corpora.

Bar: per-frame indices, stamps and final states bit-exact; fp32 DP scores bit-exact when both
sides consume identical log-probs (the DP is single fp32 adds), 1e-4 relative where expf/logf
are involved (boost+log_softmax, confidences) -- the tolerance north_star states."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _cases(npz):
    return [str(c) for c in npz["__cases__"]]


# ------------------------------------------------------------------------------------------------
def test_viterbi_core_golden_bit_exact(golden, bfa, dev):
    """bfa_viterbi_paths == _viterbi_decode (forced_alignment.py:563-703) incl. degenerate wrap cases."""
    g = golden("viterbi_core")
    for forced in (True, False):
        names = [n for n in _cases(g) if bool(g[f"{n}/meta"][2]) == forced]
        by_c = {}
        for n in names:
            by_c.setdefault((g[f"{n}/lp"].shape[1], int(g[f"{n}/meta"][1])), []).append(n)
        for (Cc, blank), group in by_c.items():
            dec = bfa.ViterbiDecoder(blank, 0, silence_anchors=10, ignore_noise=True, truly_forced=forced)
            lps = [torch.from_numpy(g[f"{n}/lp"]) for n in group]
            T = [int(x.shape[0]) for x in lps]
            row_off = np.concatenate([[0], np.cumsum([t * Cc for t in T])[:-1]])
            flat = torch.cat([x.reshape(-1) for x in lps]).to(dev)
            path = torch.cat([torch.from_numpy(g[f"{n}/path"]) for n in group])
            tidx = torch.cat([torch.from_numpy(g[f"{n}/tidx"]) for n in group])
            L = [int(g[f"{n}/path"].shape[0]) for n in group]
            band = [int(g[f"{n}/meta"][0]) for n in group]
            r = dec.viterbi_paths(flat.view(-1, Cc), T, row_off, path, tidx, L, band)
            fo = r["frame_off"]
            fph, fix = r["frame_ph"].cpu().numpy(), r["frame_idx"].cpu().numpy()
            dpf, fs = r["dp_final"].cpu().numpy(), r["final_state"].cpu().numpy()
            for i, n in enumerate(group):
                fstate = int(g[f"{n}/meta"][3])
                assert fs[i] == fstate, n
                np.testing.assert_array_equal(fph[fo[i]:fo[i + 1]], g[f"{n}/frame_ph"], err_msg=n)
                np.testing.assert_array_equal(fix[fo[i]:fo[i + 1]], g[f"{n}/frame_idx"], err_msg=n)
                assert dpf[i].tobytes() == g[f"{n}/dp_last"][fstate].tobytes(), n


def test_single_viterbi_decode_method(golden, bfa, dev):
    g = golden("viterbi_core")
    n = "cfg1_T60_N8_C67"
    dec = bfa.ViterbiDecoder(66, 0, truly_forced=True)
    fp, fi = dec._viterbi_decode(torch.from_numpy(g[f"{n}/lp"]).to(dev), torch.from_numpy(g[f"{n}/path"]).long(),
                                 len(g[f"{n}/path"]), torch.from_numpy(g[f"{n}/tidx"]).long(), band_width=0)
    assert fp.dtype == torch.int64 and fi.dtype == torch.int64
    np.testing.assert_array_equal(fp.cpu().numpy(), g[f"{n}/frame_ph"])
    np.testing.assert_array_equal(fi.cpu().numpy(), g[f"{n}/frame_idx"])


def test_decode_forced_golden(golden, bfa, dev):
    """decode_with_forced_alignment (forced_alignment.py:87-199): every branch in the fixture set."""
    g = golden("decode_forced")
    for name in _cases(g):
        blank, sil, anchors, forced, boost, floor, err, segmented = (int(v) for v in g[f"{name}/meta"])
        dec = bfa.ViterbiDecoder(blank, None if sil < 0 else sil, silence_anchors=anchors, ignore_noise=True, truly_forced=bool(forced))
        lp = torch.from_numpy(g[f"{name}/lp"]).to(dev)
        seq = torch.from_numpy(g[f"{name}/seq"]).long()
        if err:
            with pytest.raises(ValueError, match="Audio too short to align"):
                dec.decode_with_forced_alignment(lp, seq, boost_targets=bool(boost), enforce_minimum=bool(floor), anchor_pauses=anchors > 0)
            continue
        fp, fi, sc = dec.decode_with_forced_alignment(lp, seq, boost_targets=bool(boost), enforce_minimum=bool(floor),
                                                      anchor_pauses=anchors > 0)
        assert sc is None and fp.dtype == torch.int64
        np.testing.assert_array_equal(fp.cpu().numpy(), g[f"{name}/frame_ph"], err_msg=name)
        np.testing.assert_array_equal(fi.cpu().numpy(), g[f"{name}/frame_idx"], err_msg=name)
        stamps = dec.assort_frames(fp, fi)
        np.testing.assert_array_equal(np.array(stamps, np.int32).reshape(-1, 4), g[f"{name}/stamps"], err_msg=name)
        conf = bfa._calculate_confidences(lp, [s + (False,) for s in stamps])
        np.testing.assert_allclose([c[5] for c in conf], g[f"{name}/conf"], rtol=RTOL, atol=1e-6, err_msg=name)


def test_batch_api_golden(golden, bfa, dev):
    """AlignmentUtils.decode_alignments / decode_alignments_simple on a padded batch with ragged lengths,
    an empty target, ignore_noise=False and the C=17 group head."""
    g = golden("batch_api")
    lp = torch.from_numpy(g["batch/lp"]).to(dev); tgt = torch.from_numpy(g["batch/tgt"]).long().to(dev)
    pred = torch.from_numpy(g["batch/pred_lens"]).long()
    au = bfa.AlignmentUtils(66, 0, silence_anchors=10, ignore_noise=True, truly_forced=True)
    res = au.decode_alignments(lp, true_seqs=tgt, pred_lens=pred, true_seqs_lens=torch.from_numpy(g["batch/seq_lens"]).long(),
                               with_confidence=True)
    for i, r in enumerate(res):
        np.testing.assert_array_equal(np.array([x[:4] for x in r], np.int32).reshape(-1, 4), g[f"batch/stamps{i}"], err_msg=str(i))
        np.testing.assert_allclose([x[4] for x in r], g[f"batch/conf{i}"], rtol=RTOL, atol=1e-6)
    res = au.decode_alignments_simple(lp, tgt, pred_lens=pred, true_seqs_lens=torch.from_numpy(g["simple/seq_lens"]).long())
    for i, r in enumerate(res):
        np.testing.assert_array_equal(np.array(r, np.int32).reshape(-1, 4), g[f"simple/stamps{i}"], err_msg=f"simple{i}")
    au2 = bfa.AlignmentUtils(66, 0, silence_anchors=0, ignore_noise=False, truly_forced=True)
    res = au2.decode_alignments(lp, true_seqs=tgt, pred_lens=pred, true_seqs_lens=torch.from_numpy(g["noise/seq_lens"]).long())
    for i, r in enumerate(res):
        np.testing.assert_array_equal(np.array(r, np.int32).reshape(-1, 4), g[f"noise/stamps{i}"], err_msg=f"noise{i}")
    aug = bfa.AlignmentUtils(16, 0)
    res = aug.decode_alignments(torch.from_numpy(g["group/lp"]).to(dev), true_seqs=torch.from_numpy(g["group/tgt"]).long(),
                                pred_lens=torch.tensor([150, 140, 100]), true_seqs_lens=torch.tensor([14, 14, 9]))
    for i, r in enumerate(res):
        np.testing.assert_array_equal(np.array(r, np.int32).reshape(-1, 4), g[f"group/stamps{i}"], err_msg=f"group{i}")


def test_missing_sequences_raise(bfa, dev):
    au = bfa.AlignmentUtils(66, 0)
    with pytest.raises(ValueError, match="required for forced alignment"):
        au.decode_alignments(torch.zeros(1, 4, 67, device=dev))


# ------------------------------------------------------------------------------------------------
def _oracle_batch(orc, p, lp, tgt, T, N, Cc, threads=8):
    B, Tm = lp.shape[0], lp.shape[1]
    tflat = np.concatenate([tgt[i, :N[i]] for i in range(B)]).astype(np.int32) if B else np.zeros(0, np.int32)
    toff = np.zeros(B + 1, np.int64); np.cumsum(N, out=toff[1:])
    return orc.align_batch(p, lp, np.arange(B, dtype=np.int64) * Tm * Cc, np.asarray(T, np.int32), Cc, tflat, toff,
                           max_stamps=2 * int(max(N)) + 8, n_threads=threads)


def _compare_with_oracle(bfa, orc, dev, lp, tgt, T, N, Cc, blank, **kw):
    au = bfa.AlignmentUtils(blank, 0, **kw)
    p = orc.params(blank, 0, kw.get("silence_anchors", 10), kw.get("ignore_noise", True), kw.get("truly_forced", True))
    dparams = au.viterbi_decoder._params(True, True, au.silence_anchors > 0)
    B, Tm = lp.shape[0], lp.shape[1]
    lp_d = lp.to(dev)
    seqs = tgt.to(dev)
    Nd = torch.tensor(N, device=dev)
    mask = torch.arange(tgt.shape[1], device=dev)[None, :] < Nd[:, None]
    r = au.viterbi_decoder.align_batch(lp_d, torch.arange(B, dtype=torch.int64, device=dev) * Tm * Cc, T, Cc,
                                       seqs[mask].to(torch.int32).contiguous(), N, params=dparams, max_stamps=2 * int(max(N)) + 8)
    o = _oracle_batch(orc, p, lp.numpy(), tgt.numpy(), T, N, Cc)
    st = r.status[:B].cpu().numpy()
    np.testing.assert_array_equal(st & 15, o["status"] & 15)
    fph, fix = r.frame_ph.cpu().numpy(), r.frame_idx.cpu().numpy()
    fo = o["frame_off"]
    bad = [b for b in range(B) if not (np.array_equal(fph[fo[b]:fo[b + 1]], o["frame_ph"][fo[b]:fo[b + 1]])
                                      and np.array_equal(fix[fo[b]:fo[b + 1]], o["frame_idx"][fo[b]:fo[b + 1]]))]
    assert not bad, f"{len(bad)} of {B} utterances differ from the oracle, first {bad[:5]}"
    nst = r.n_stamps[:B].cpu().numpy()
    np.testing.assert_array_equal(nst, o["n_stamps"])
    stamps = r.stamps.cpu().numpy(); conf = r.conf.cpu().numpy()
    for b in range(B):
        n = nst[b]
        want = np.stack([o["stamps"][b][f][:n] for f in ("phoneme", "start", "end", "target_idx")], 1)
        np.testing.assert_array_equal(stamps[b, :n], want)
        np.testing.assert_allclose(conf[b, :n], o["conf"][b, :n], rtol=RTOL, atol=1e-6)
    unseg = (st & 7) == 0
    np.testing.assert_allclose(r.dp_final[:B].cpu().numpy()[unseg], o["dp_final"][unseg], rtol=RTOL)
    return st


def test_metric_shape_vs_oracle(bfa, orc, dev):
    """B=256 of the metric shape T=600,N=40,C=66 (full mode: boost + floor + anchoring enabled)."""
    from bfa_b200 import synth
    lp, tgt, _ = synth.planted_batch(256, 600, 40, 66, seed=101)
    st = _compare_with_oracle(bfa, orc, dev, lp, tgt, [600] * 256, [40] * 256, 66, 65)
    assert (st == 0).all()


def test_segmented_long_form_vs_oracle(bfa, orc, dev):
    """config 3 shape: T=3600, N=200 with SIL anchors -> silence-anchored segmentation on the device."""
    from bfa_b200 import synth
    lp, tgt, _ = synth.planted_batch(12, 3600, 200, 66, seed=102, peak=12.0, sil_every=40, sil_frames=18)
    st = _compare_with_oracle(bfa, orc, dev, lp, tgt, [3600] * 12, [200] * 12, 66, 65)
    assert ((st & 7) == 4).sum() >= 10


def test_segmented_mixed_vs_oracle(bfa, orc, dev):
    from bfa_b200 import synth
    for seed, (T, N, Cc, every, frames, anchors, forced) in enumerate([
            (900, 60, 67, 12, 22, 10, True), (400, 60, 67, 10, 14, 10, True), (420, 100, 67, 9, 12, 10, False),
            (300, 24, 17, 6, 16, 10, True), (700, 40, 67, 8, 25, 3, True), (500, 50, 30, 5, 8, 3, True)]):
        lp, tgt, _ = synth.planted_batch(24, T, N, Cc, seed=200 + seed, peak=10.0, sil_every=every, sil_frames=frames)
        _compare_with_oracle(bfa, orc, dev, lp, tgt, [T] * 24, [N] * 24, Cc, Cc - 1, silence_anchors=anchors, truly_forced=forced)


def test_ragged_lengths_and_strides_vs_oracle(bfa, orc, dev):
    """Padded batch with ragged pred_lens / true_seqs_lens hitting strides 4/3/2/1, T==N, N==0."""
    from bfa_b200 import synth
    rng = np.random.default_rng(7)
    B, Tm, Nm, Cc = 96, 400, 120, 67
    lp, tgt, _ = synth.planted_batch(B, Tm, Nm, Cc, seed=300, peak=10.0)
    T = rng.integers(60, Tm + 1, B).tolist()
    N = [int(min(Nm, max(1, t // d))) for t, d in zip(T, rng.choice([1, 2, 3, 4, 6, 12], B))]
    N[5] = 0; T[6] = N[6] = 77
    # planted alignment does not match the shortened targets; re-plant per utterance
    for b in range(B):
        if N[b] > 0:
            l, t2, _ = synth.planted_batch(1, T[b], N[b], Cc, seed=1000 + b, peak=10.0)
            lp[b, :T[b]] = l[0]; tgt[b, :N[b]] = t2[0]
    st = _compare_with_oracle(bfa, orc, dev, lp, tgt, T, N, Cc, Cc - 1)
    assert (st & 7 == 3).any() and (st & 7 == 1).any()


def test_degenerate_inputs_vs_oracle(bfa, orc, dev):
    """Flat random posteriors: every path below -1000, masked candidates win, back-pointers wrap."""
    g = torch.Generator().manual_seed(11)
    B, T, N, Cc = 32, 600, 40, 66
    lp = torch.log_softmax(torch.randn(B, T, Cc, generator=g) * 3.0, -1)
    tgt = torch.randint(1, Cc - 1, (B, N), generator=g)
    st = _compare_with_oracle(bfa, orc, dev, lp, tgt, [T] * B, [N] * B, Cc, Cc - 1, silence_anchors=0)
    assert (st & 8).all()


def test_too_short_raises_like_reference(bfa, dev):
    from bfa_b200 import synth
    lp, tgt, _ = synth.planted_batch(2, 30, 40, 67, seed=5)
    au = bfa.AlignmentUtils(66, 0)
    with pytest.raises(ValueError, match="40 phonemes cannot be fit into 30 frames"):
        au.decode_alignments(lp.to(dev), true_seqs=tgt, pred_lens=torch.tensor([30, 30]), true_seqs_lens=torch.tensor([40, 40]))


# ---- size-independent properties at BASELINE.json's full batch -------------------------------------
def test_full_batch_properties(bfa, orc, dev):
    """B=4096, T=600, N=40, C=66: (i) dp_final equals the sequential fp32 sum of the modified log-probs along
    the returned path, (ii) target indices are monotone and cover all 40 phonemes, (iii) a
    random sample of utterances is bit-identical to the oracle."""
    from bfa_b200 import synth
    B, T, N, Cc = 4096, 600, 40, 66
    lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=77, device=dev)
    au = bfa.AlignmentUtils(Cc - 1, 0)
    dec = au.viterbi_decoder
    p = dec._params(True, True, True)
    r = dec.align_batch(lp, torch.arange(B, dtype=torch.int64, device=dev) * T * Cc, [T] * B, Cc,
                        tgt.to(torch.int32).reshape(-1).contiguous(), [N] * B, params=p)
    assert (r.status[:B] == 0).all()
    fph = r.frame_ph.view(B, T).long(); fix = r.frame_idx.view(B, T).long()
    # (ii)
    idx = fix.clone(); idx[idx < 0] = 0
    run = torch.cummax(idx, dim=1).values
    assert ((fix < 0) | (fix == run)).all()
    assert (r.n_stamps[:B] == N).all()
    assert (r.stamps[:, :N, 3] == torch.arange(N, device=dev, dtype=torch.int32)[None, :]).all()
    # (iii) + (i) on a sample, oracle-side
    sample = list(range(0, B, 173))
    lp_s = lp[sample].cpu(); tgt_s = tgt[sample].cpu()
    o = _oracle_batch(orc, orc.params(Cc - 1, 0), lp_s.numpy(), tgt_s.numpy(), [T] * len(sample), [N] * len(sample), Cc)
    np.testing.assert_array_equal(fph[sample].cpu().numpy().reshape(-1), o["frame_ph"])
    np.testing.assert_array_equal(fix[sample].cpu().numpy().reshape(-1), o["frame_idx"])
    dpf = r.dp_final[sample].cpu().numpy()
    for i in range(len(sample)):
        m = orc.prep(lp_s[i].numpy(), tgt_s[i].numpy().astype(np.int32), orc.params(Cc - 1, 0))
        acc = np.float32(m[0, o["frame_ph"][i * T]])
        for t in range(1, T):
            acc = np.float32(acc + m[t, o["frame_ph"][i * T + t]])
        assert abs(dpf[i] - acc) <= RTOL * abs(acc)


def test_host_entry_matches_device_entry(bfa, dev):
    """bfa_align_batch_host (host buffers, chunked copies) == bfa_align_batch on the same data."""
    from bfa_b200 import synth
    B, T, N, Cc = 300, 200, 20, 67
    lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=55, peak=10.0, sil_every=7, sil_frames=14)
    au = bfa.AlignmentUtils(Cc - 1, 0)
    p = au.viterbi_decoder._params(True, True, True)
    r = au.viterbi_decoder.align_batch(lp.to(dev), torch.arange(B, dtype=torch.int64, device=dev) * T * Cc, [T] * B, Cc,
                                       tgt.to(torch.int32).reshape(-1).to(dev), [N] * B, params=p)
    h = bfa.align_host(p, lp.numpy(), np.arange(B, dtype=np.int64) * T * Cc, np.full(B, T, np.int32), Cc,
                       tgt.numpy().astype(np.int32).reshape(-1), np.arange(B + 1, dtype=np.int64) * N, chunk_utts=64)
    np.testing.assert_array_equal(h["frame_ph"], r.frame_ph.cpu().numpy())
    np.testing.assert_array_equal(h["frame_idx"], r.frame_idx.cpu().numpy())
    np.testing.assert_array_equal(h["status"], r.status.cpu().numpy())
    np.testing.assert_array_equal(h["n_stamps"], r.n_stamps.cpu().numpy())
    n = h["n_stamps"]
    for b in range(B):
        np.testing.assert_array_equal(h["stamps"][b, :n[b]], r.stamps[b, :n[b]].cpu().numpy())
        np.testing.assert_array_equal(h["conf"][b, :n[b]], r.conf[b, :n[b]].cpu().numpy())


def test_no_sil_hint_is_only_a_hint(bfa, dev):
    """BFA_HINT_NO_SIL skips the row-statistics pass; asserting it wrongly (targets DO hold SIL) must not
    change any result (the planner recomputes what it needs)."""
    from bfa_b200 import synth, _cabi
    B, T, N, Cc = 16, 400, 40, 67
    lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=91, peak=10.0, sil_every=8, sil_frames=16)
    dec = bfa.AlignmentUtils(Cc - 1, 0).viterbi_decoder
    row_off = torch.arange(B, dtype=torch.int64, device=dev) * T * Cc
    tg = tgt.to(torch.int32).reshape(-1).to(dev)
    outs = []
    for hint in (0, _cabi.HINT_NO_SIL):
        p = dec._params(True, True, True)
        p.reserved |= hint
        r = dec.align_batch(lp.to(dev), row_off, [T] * B, Cc, tg, [N] * B, params=p)
        outs.append((r.frame_ph.cpu().numpy(), r.frame_idx.cpu().numpy(), r.status.cpu().numpy(), r.conf.cpu().numpy(), r.n_stamps.cpu().numpy()))
    assert ((outs[0][2] & 7) == 4).sum() >= B // 2            # segmentation really happened
    for a_, b_ in zip(outs[0][:3], outs[1][:3]):
        np.testing.assert_array_equal(a_, b_)
    n = outs[0][4]
    for b in range(B):
        np.testing.assert_allclose(outs[0][3][b, :n[b]], outs[1][3][b, :n[b]], rtol=1e-5)


def test_exact_only_flag_matches_fast_path(bfa, dev):
    """BFA_FLAG_EXACT_ONLY routes everything through the generic exact kernel; on valid inputs the banded fast
    kernel must produce the same frames (scores within fp tolerance of the fused log-sum-exp)."""
    from bfa_b200 import synth, _cabi
    B, T, N, Cc = 64, 600, 40, 66
    lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=92)
    dec = bfa.AlignmentUtils(Cc - 1, 0).viterbi_decoder
    row_off = torch.arange(B, dtype=torch.int64, device=dev) * T * Cc
    tg = tgt.to(torch.int32).reshape(-1).to(dev)
    res = []
    for flag in (0, _cabi.FLAG_EXACT_ONLY):
        p = dec._params(True, True, True)
        p.reserved |= flag
        r = dec.align_batch(lp.to(dev), row_off, [T] * B, Cc, tg, [N] * B, params=p)
        res.append(r)
    np.testing.assert_array_equal(res[0].frame_ph.cpu().numpy(), res[1].frame_ph.cpu().numpy())
    np.testing.assert_array_equal(res[0].frame_idx.cpu().numpy(), res[1].frame_idx.cpu().numpy())
    np.testing.assert_allclose(res[0].dp_final.cpu().numpy(), res[1].dp_final.cpu().numpy(), rtol=RTOL)
    np.testing.assert_allclose(res[0].conf.cpu().numpy()[:, :N], res[1].conf.cpu().numpy()[:, :N], rtol=1e-5)


# ---- packed corpora, arbitrary row alignment, every mode -------------------------------------------
def _packed_vs_oracle(bfa, orc, dev, utts, Cc, *, gap_floats, boost=True, floor=True, mode=0, anchors=10, base_shift=0):
    """Rows packed back to back with `gap_floats` floats between utterances (so that row starts land on every
    residue of the 16-byte grid the bulk copies need) -> bfa_align_batch vs the oracle, utterance by utterance."""
    from bfa_b200 import _cabi
    import ctypes as C
    B = len(utts)
    Ts = [int(l.shape[0]) for l, _ in utts]; Ns = [int(t.shape[0]) for _, t in utts]
    offs, cur = [], base_shift
    for t, g in zip(Ts, gap_floats):
        offs.append(cur); cur += t * Cc + g
    flat = torch.zeros(cur + 8, dtype=torch.float32)
    for (l, _), o, t in zip(utts, offs, Ts):
        flat[o:o + t * Cc] = l.reshape(-1)
    au = bfa.AlignmentUtils(Cc - 1, 0, silence_anchors=anchors)
    dec = au.viterbi_decoder
    p = dec._params(boost, floor, anchors > 0, mode=mode)
    tg = torch.cat([t for _, t in utts]).to(torch.int32).contiguous()
    r = dec.align_batch(flat.to(dev), torch.tensor(offs, dtype=torch.int64, device=dev), Ts, Cc, tg.to(dev), Ns, params=p)
    torch.cuda.synchronize()
    ic = (C.c_int32 * 4)(); _cabi.lib().bfa_debug_item_counts(ic)
    po = orc.params(Cc - 1, 0, anchors, True, True, boost, floor, mode)
    fo = np.zeros(B + 1, np.int64); np.cumsum(np.asarray(Ts, np.int64), out=fo[1:])
    fph, fix, st = r.frame_ph.cpu().numpy(), r.frame_idx.cpu().numpy(), r.status.cpu().numpy()
    nst, stamps, conf = r.n_stamps.cpu().numpy(), r.stamps.cpu().numpy(), r.conf.cpu().numpy()
    for u, (l, t) in enumerate(utts):
        T, N = Ts[u], Ns[u]
        o = orc.align_batch(po, l.numpy().reshape(1, T, Cc), np.zeros(1, np.int64), np.asarray([T], np.int32), Cc,
                            t.numpy().astype(np.int32), np.asarray([0, N], np.int64), max_stamps=r.max_stamps, n_threads=1)
        assert (int(o["status"][0]) & 15) == (int(st[u]) & 15), f"utterance {u}: status"
        np.testing.assert_array_equal(fph[fo[u]:fo[u + 1]], o["frame_ph"], err_msg=f"utterance {u} (T={T}, N={N})")
        np.testing.assert_array_equal(fix[fo[u]:fo[u + 1]], o["frame_idx"], err_msg=f"utterance {u} (T={T}, N={N})")
        n = int(o["n_stamps"][0])
        assert n == int(nst[u])
        if n:
            want = np.stack([o["stamps"][0][f][:n] for f in ("phoneme", "start", "end", "target_idx")], 1)
            np.testing.assert_array_equal(stamps[u, :n], want)
            np.testing.assert_allclose(conf[u, :n], o["conf"][0, :n], rtol=RTOL, atol=1e-6)
    return [int(x) for x in ic]


@pytest.mark.parametrize("Cc", [66, 67, 30])
def test_packed_rows_on_every_alignment_vs_oracle(bfa, orc, dev, Cc):
    """Row starts on all four residues of the 16-byte grid (Item::lead 0..3): the banded kernel keeps its bulk copies by
    starting them at the boundary before the item; what it cannot take (odd lead at C=66) runs in the exact kernel."""
    from bfa_b200 import synth
    rng = np.random.default_rng(17 + Cc)
    utts = synth.ragged_batch(48, C=Cc, t_range=(60, 500), n_range=(4, 40), seed=400 + Cc)
    gaps = rng.integers(0, 4, len(utts)).tolist()
    counts = _packed_vs_oracle(bfa, orc, dev, utts, Cc, gap_floats=gaps, base_shift=int(rng.integers(0, 4)))
    assert sum(counts[1:]) > len(utts) // 3, f"the banded kernels were hardly used: {counts}"


@pytest.mark.parametrize("boost,floor,mode", [(False, True, 0), (False, False, 0), (True, False, 0), (True, True, 1)])
def test_packed_rows_all_modes_vs_oracle(bfa, orc, dev, boost, floor, mode):
    """boost off / floor off / decode_alignments_simple (mode 1): without boost the banded kernel runs its EXACT instantiation."""
    from bfa_b200 import synth
    utts = synth.ragged_batch(40, C=66, t_range=(80, 700), n_range=(4, 60), seed=500 + 2 * int(boost) + int(floor) + 10 * mode)
    gaps = [0, 2] * (len(utts) // 2)
    counts = _packed_vs_oracle(bfa, orc, dev, utts, 66, gap_floats=gaps, boost=boost, floor=floor, mode=mode, anchors=10 if mode == 0 else 0)
    assert sum(counts[1:]) > 0


def test_segmented_packed_misaligned_vs_oracle(bfa, orc, dev):
    """Silence-anchored segments of utterances that themselves start off the 16-byte grid."""
    from bfa_b200 import synth
    utts = []
    for u in range(12):
        lp, tgt, _ = synth.planted_batch(1, 900 + 37 * u, 60, 66, seed=600 + u, peak=10.0, sil_every=12, sil_frames=20)
        utts.append((lp[0], tgt[0]))
    counts = _packed_vs_oracle(bfa, orc, dev, utts, 66, gap_floats=[2, 0, 6, 2] * 3, base_shift=2)
    assert sum(counts[1:]) > 12


def test_stamp_overflow_is_flagged_and_the_reference_api_recovers(bfa, orc, dev):
    """Degenerate posteriors produce more runs than targets: the low-level call flags ST_STAMP_OVERFLOW at the default stamp
    pitch (N + 8); decode_alignments repeats the call with the safe pitch and returns what the oracle returns."""
    from bfa_b200 import _cabi
    g = torch.Generator().manual_seed(12)
    B, T, N, Cc = 8, 600, 40, 66
    lp = torch.log_softmax(torch.randn(B, T, Cc, generator=g) * 3.0, -1)
    tgt = torch.randint(1, Cc - 1, (B, N), generator=g)
    au = bfa.AlignmentUtils(Cc - 1, 0, silence_anchors=0)
    dec = au.viterbi_decoder
    p = dec._params(True, True, False)
    r = dec.align_batch(lp.to(dev), torch.arange(B, dtype=torch.int64, device=dev) * T * Cc, [T] * B, Cc,
                        tgt.to(torch.int32).reshape(-1).to(dev), [N] * B, params=p)
    o = _oracle_batch(orc, orc.params(Cc - 1, 0, 0), lp.numpy(), tgt.numpy(), [T] * B, [N] * B, Cc)
    big = o["n_stamps"] > r.max_stamps
    st = r.status[:B].cpu().numpy()
    np.testing.assert_array_equal((st & _cabi.ST_STAMP_OVERFLOW) != 0, big)
    got = au.decode_alignments(lp.to(dev), true_seqs=tgt, pred_lens=torch.full((B,), T), true_seqs_lens=torch.full((B,), N))
    o2 = orc.align_batch(orc.params(Cc - 1, 0, 0), lp.numpy(), np.arange(B, dtype=np.int64) * T * Cc, np.full(B, T, np.int32), Cc,
                         tgt.numpy().astype(np.int32).reshape(-1), np.arange(B + 1, dtype=np.int64) * N, max_stamps=T, n_threads=4)
    for b in range(B):
        n = int(o2["n_stamps"][b])
        want = [tuple(int(o2["stamps"][b][f][i]) for f in ("phoneme", "start", "end", "target_idx")) for i in range(n)]
        assert got[b] == want


@pytest.mark.parametrize("seed", list(range(1, 13)))
def test_random_configurations_vs_oracle(bfa, orc, dev, seed):
    """Seeded sweep over class counts, length ranges (partial chunks, T < 8, N = 1, paths at the window-class boundaries),
    decoder switches and row alignment: every utterance must match the oracle."""
    from bfa_b200 import synth
    rng = np.random.default_rng(9000 + seed)
    Cc = int(rng.choice([9, 17, 33, 48, 66, 67, 72]))
    t_hi = int(rng.choice([40, 200, 700, 1300]))
    n_hi = int(rng.choice([3, 20, 60, 130]))
    utts = synth.ragged_batch(36, C=Cc, t_range=(max(2, t_hi // 12), t_hi), n_range=(1, max(1, min(n_hi, Cc * 4))), seed=7000 + seed,
                              peak=float(rng.choice([7.0, 10.0, 13.0])))
    gaps = rng.integers(0, 4, len(utts)).tolist()
    boost, floor = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    mode = int(rng.integers(0, 2))
    _packed_vs_oracle(bfa, orc, dev, utts, Cc, gap_floats=gaps, boost=boost, floor=floor, mode=mode,
                      anchors=int(rng.choice([0, 3, 10])) if mode == 0 else 0, base_shift=int(rng.integers(0, 4)))


def test_soft_boundaries_golden_and_oracle(golden, bfa, orc, dev):
    """extend_soft_boundaries_func (core.py:682-809): the batch kernel against fixtures from the unmodified method, one
    ragged batch per class count (utterances padded to the longest, shorter ones exercise T[u] < T_max via the C ABI)."""
    import ctypes as C
    from pathlib import Path
    from bfa_b200 import _cabi
    g = np.load(Path(__file__).parent / "golden" / "soft.npz")
    cases = sorted({k.split("/")[0] for k in g.files}, key=lambda c: int(c[1:]))
    # (a) reference-style call, one utterance at a time
    for c in cases:
        lp = torch.from_numpy(g[f"{c}/lp"]).to(dev)[None]
        fs = [[(int(r[0]), int(r[1]), int(r[2]), int(r[3]), bool(r[4])) for r in g[f"{c}/in"]]]
        got = bfa.extend_soft_boundaries_func(lp, fs, boundary_softness=int(g[f"{c}/soft"][0]))
        want = [(int(r[0]), int(r[1]), int(r[2]), int(r[3]), bool(r[4])) for r in g[f"{c}/out"]]
        assert got[0] == want, c
    # (b) one launch over utterances of different length and stamp count that share C and the softness (C ABI, packed rows)
    groups = {}
    for c in cases:
        groups.setdefault((g[f"{c}/lp"].shape[1], int(g[f"{c}/soft"][0])), []).append(c)
    for (Cc, soft), cs in groups.items():
        lps = [g[f"{c}/lp"] for c in cs]
        Ts = [l.shape[0] for l in lps]
        offs = np.concatenate([[0], np.cumsum([l.size for l in lps])]).astype(np.int64)
        flat = torch.from_numpy(np.concatenate([l.reshape(-1) for l in lps])).to(dev)
        ms = max(len(g[f"{c}/in"]) for c in cs)
        st = np.zeros((len(cs), ms, 4), np.int32)
        for b, c in enumerate(cs):
            st[b, :len(g[f"{c}/in"])] = g[f"{c}/in"][:, :4]
        st_d = torch.from_numpy(st).to(dev)
        n_d = torch.tensor([len(g[f"{c}/in"]) for c in cs], dtype=torch.int32, device=dev)
        off_d = torch.from_numpy(offs[:-1].copy()).to(dev)          # keep the device arrays alive across the call
        T_d = torch.tensor(Ts, dtype=torch.int32, device=dev)
        rc = _cabi.lib().bfa_soft_boundaries_batch(len(cs), Cc, C.c_void_p(flat.data_ptr()), C.c_void_p(off_d.data_ptr()),
                                                   C.c_void_p(T_d.data_ptr()), C.c_void_p(st_d.data_ptr()),
                                                   C.c_void_p(n_d.data_ptr()), ms, soft, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert rc == 0
        out = st_d.cpu().numpy()
        for b, c in enumerate(cs):
            np.testing.assert_array_equal(out[b, :len(g[f"{c}/out"])], g[f"{c}/out"][:, :4], err_msg=c)
            o = orc.soft_boundaries(g[f"{c}/lp"], [tuple(r) for r in g[f"{c}/in"]], soft)
            np.testing.assert_array_equal(out[b, :len(o)], np.array(o, np.int32).reshape(-1, 4))


def test_post_acoustic_pipeline_vs_reference(bfa, dev):
    """decode_alignments -> ensure_target_coverage -> extend_soft_boundaries_func -> _calculate_confidences -> convert_to_ms
    (core.py:902-937, :960-975) on this library against the same chain of the unmodified reference (tests/golden/make_golden_pipeline.py):
    phonemes, frames, target indices, flags and milliseconds identical, confidences within 1e-4."""
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "pipeline.npz")
    utts = sorted({k.split("/")[0] for k in g.files}, key=lambda c: int(c[1:]))
    assert len(utts) == 11
    for u in utts:
        T, N, Cc, b = (int(x) for x in g[f"{u}/meta"])
        lp = torch.from_numpy(g[f"{u}/lp"]).to(dev)[None]
        tgt = torch.from_numpy(g[f"{u}/tgt"]).long()[None]
        au = bfa.AlignmentUtils(blank_id=Cc - 1, silence_id=0, silence_anchors=10, ignore_noise=True, truly_forced=True)
        frames = au.decode_alignments(lp, true_seqs=tgt, pred_lens=torch.tensor([T]), true_seqs_lens=torch.tensor([N]))
        frames = bfa.ensure_target_coverage(tgt, frames, seq_lens=torch.tensor([N]), _silence_class=0)
        frames = bfa.extend_soft_boundaries_func(lp, frames, boundary_softness=3)
        fs = bfa._calculate_confidences(lp[0], frames[0])
        fs = bfa.convert_to_ms(fs, T, 1.5 * b, T * 320, 16000)
        want = g[f"{u}/out"]
        assert len(fs) == len(want), u
        got = np.array([[float(x) for x in f] for f in fs], np.float64).reshape(-1, 8)
        np.testing.assert_array_equal(got[:, :5], want[:, :5], err_msg=u)
        np.testing.assert_allclose(got[:, 5], want[:, 5], rtol=RTOL, atol=1e-7, err_msg=u)
        np.testing.assert_array_equal(got[:, 6:], want[:, 6:], err_msg=u)


def test_confidences_batch_equals_per_utterance(bfa, dev):
    """One launch over the batch (bfa_confidence_batch) == the reference-style per-utterance calls, padded lengths included."""
    from bfa_b200 import synth
    B, T, N, Cc = 6, 220, 18, 67
    lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=31, peak=5.0)
    lens = torch.tensor([220, 200, 180, 220, 150, 199])
    au = bfa.AlignmentUtils(Cc - 1, 0)
    frames = au.decode_alignments(lp.to(dev), true_seqs=tgt, pred_lens=lens, true_seqs_lens=torch.full((B,), N))
    frames = bfa.ensure_target_coverage(tgt, frames, seq_lens=torch.full((B,), N))
    one = bfa._calculate_confidences_batch(lp.to(dev), frames, pred_lens=lens)
    for b in range(B):
        ref = bfa._calculate_confidences(lp[b, :int(lens[b])].to(dev), frames[b])
        assert [f[:5] for f in one[b]] == [f[:5] for f in ref]
        np.testing.assert_array_equal([f[5] for f in one[b]], [f[5] for f in ref])


def test_shim_aligner_from_posteriors_vs_reference(bfa, dev):
    """PhonemeTimestampAligner.timestamps_from_posteriors (the post-acoustic half of core.py:896-957, whole batch at once) against
    the chain of the unmodified reference pieces; the group head runs through the same code at C = 17."""
    from pathlib import Path
    from bfa_b200 import synth
    g = np.load(Path(__file__).parent / "golden" / "pipeline.npz")
    utts = sorted({k.split("/")[0] for k in g.files}, key=lambda c: int(c[1:]))
    al = bfa.PhonemeTimestampAligner(blank_class=66, silence_class=0, ensure_completeness=True)
    batches = {}
    for u in utts:                                   # utterances of one golden batch share (N, C) and the padded length
        T, N, Cc, b = (int(x) for x in g[f"{u}/meta"])
        batches.setdefault((g[f"{u}/lp"].shape[0], N, Cc), []).append((u, T, b))
    for (Tm, N, Cc), items in batches.items():
        lp = torch.from_numpy(np.stack([g[f"{u}/lp"] for u, _, _ in items])).to(dev)
        tgt = torch.from_numpy(np.stack([g[f"{u}/tgt"] for u, _, _ in items])).long()
        lens = torch.tensor([T for _, T, _ in items])
        offs = [1.5 * b for _, _, b in items]
        res = al.timestamps_from_posteriors(lp, tgt, torch.full((len(items),), N), lens, [int(T) * 320 for _, T, _ in items], offs)
        for (u, _, _), r in zip(items, res):
            want = g[f"{u}/out"]
            got = np.array([[float(x) for x in f] for f in r["phoneme_timestamps"]], np.float64).reshape(-1, 8)
            np.testing.assert_array_equal(got[:, :5], want[:, :5], err_msg=u)
            np.testing.assert_allclose(got[:, 5], want[:, 5], rtol=RTOL, atol=1e-7, err_msg=u)
            np.testing.assert_array_equal(got[:, 6:], want[:, 6:], err_msg=u)
            assert r["group_timestamps"] is None
    # both heads
    B, T, N = 3, 200, 16
    lp_p, tgt_p, _ = synth.planted_batch(B, T, N, 67, seed=41, peak=6.0)
    lp_g, tgt_g, _ = synth.planted_batch(B, T, N, 17, seed=42, peak=6.0)
    res = al.timestamps_from_posteriors(lp_p.to(dev), tgt_p, torch.full((B,), N), torch.full((B,), T), [T * 320] * B, 0.0,
                                        log_probs_g=lp_g.to(dev), grp_seqs=tgt_g)
    for r in res:
        assert [f[3] for f in r["phoneme_timestamps"]] == list(range(N)) and [f[3] for f in r["group_timestamps"]] == list(range(N))
        assert all(len(f) == 8 and f[7] >= f[6] for f in r["group_timestamps"])


def test_align_batch_replays_from_a_cuda_graph(bfa, dev):
    """The whole launch sequence of bfa_align_batch (planner, banded kernel with its dependent launches, the exact kernel's
    pass on the forked side streams, stamps + confidences) is capturable: a CUDA graph captured on one batch and replayed on
    new posteriors in the same buffers gives exactly what a direct call gives."""
    from bfa_b200 import synth
    Cc = 66
    Ts = [600] * 24 + [90] * 4 + [300] * 4          # 90 frames for 40 phonemes: stride 2 -> the exact kernel (hint > 0)
    Ns = [40] * 32
    B = len(Ts)
    dec = bfa.AlignmentUtils(Cc - 1, 0).viterbi_decoder
    params = dec._params(True, True, True)

    def batch(seed):
        rows, tg = [], []
        for b in range(B):
            lp, tgt, _ = synth.planted_batch(1, Ts[b], Ns[b], Cc, seed=seed + b, peak=9.0)
            rows.append(lp.reshape(-1)); tg.append(tgt.reshape(-1))
        return torch.cat(rows).to(dev), torch.cat(tg).to(torch.int32).to(dev)

    lp0, tg0 = batch(1000)
    lp1, tg1 = batch(2000)
    row_off = torch.tensor(np.concatenate([[0], np.cumsum(np.asarray(Ts[:-1], np.int64) * Cc)]), dtype=torch.int64, device=dev)
    plan = dec.plan_batch(Ts, Ns, Cc, params=params, device=dev)
    assert plan.shape.reserved > 0                   # the side-stream pass is part of what gets captured
    lp_buf, tg_buf = lp0.clone(), tg0.clone()
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        res = dec.align_batch(lp_buf, row_off, Ts, Cc, tg_buf, Ns, params=params, plan=plan)        # warm-up: allocates everything
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            dec.align_batch(lp_buf, row_off, Ts, Cc, tg_buf, Ns, params=params, plan=plan, out=res)
    torch.cuda.current_stream(dev).wait_stream(s)
    lp_buf.copy_(lp1); tg_buf.copy_(tg1)
    res.arena.zero_(); res.frame_ph.fill_(-7); res.frame_idx.fill_(-7)
    g.replay()
    torch.cuda.synchronize()
    got = [t.cpu().numpy().copy() for t in (res.frame_ph, res.frame_idx, res.status, res.n_stamps, res.stamps, res.conf, res.dp_final)]
    ref = dec.align_batch(lp1, row_off, Ts, Cc, tg1, Ns, params=params, plan=plan)
    torch.cuda.synchronize()
    want = [t.cpu().numpy() for t in (ref.frame_ph, ref.frame_idx, ref.status, ref.n_stamps, ref.stamps, ref.conf, ref.dp_final)]
    assert (want[2] & 7 == 0).all()
    for a_, b_ in zip(got[:4], want[:4]):
        np.testing.assert_array_equal(a_, b_)
    for b in range(B):
        n = want[3][b]
        np.testing.assert_array_equal(got[4][b, :n], want[4][b, :n])
        np.testing.assert_array_equal(got[5][b, :n], want[5][b, :n])
    np.testing.assert_array_equal(got[6], want[6])
