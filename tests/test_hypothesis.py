"""Property-based tests (hypothesis), SURVEY.md section 4.
  * CPU: the C oracle's DP core against an independent numpy restatement of _viterbi_decode (forced_alignment.py:563-703)
    written from the reference's tensor expressions, on drawn paths / bands / posteriors; path invariants of decode_forced;
    and, when /root/reference is present (build container), the oracle's full path against the live reference module.
  * GPU: the CUDA library against the oracle on drawn batch shapes (ragged lengths, class counts, silence layout, modes)."""
import os
from pathlib import Path

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

SET = dict(deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
NEG = np.float32(-1000.0)


def _viterbi_numpy(lp, path, band, blank, truly_forced=True):
    """_viterbi_decode restated with numpy, expression by expression (:579-703)."""
    T, L = lp.shape[0], len(path)
    dp = np.full((T, L), NEG, np.float32)
    bp = np.zeros((T, L), np.int64)
    use_band = band > 0 and T > 1 and L > 1
    pace = (L - 1) / (T - 1) if use_band else 0.0
    idx = np.arange(L, dtype=np.float32)
    dp[0, 0] = lp[0, blank]
    if L > 1:
        dp[0, 1] = lp[0, path[1]]
    can_adv = np.zeros(L, bool); can_adv[1:] = True
    can_skip = np.zeros(L, bool)
    for s in range(2, L):
        can_skip[s] = path[s] != path[s - 2]
    for t in range(1, T):
        e = lp[t, path]
        prev = dp[t - 1]
        stay = prev + e
        adv = np.full(L, NEG, np.float32); adv[1:] = prev[:-1] + e[1:]
        skp = np.full(L, NEG, np.float32); skp[2:] = np.where(can_skip[2:], prev[:-2] + e[2:], NEG)
        allv = np.stack([stay, adv, skp], 1)
        mask = np.stack([np.ones(L, bool), can_adv, can_skip], 1)
        allv = np.where(mask, allv, NEG).astype(np.float32)
        k = np.argmax(allv, 1)                       # first max wins, like torch.argmax
        dp[t] = allv[np.arange(L), k]
        bp[t] = np.arange(L) - k
        if use_band:
            center = t * pace
            out = (idx < np.float32(center - band)) | (idx > np.float32(center + band))
            dp[t, out] = NEG
    last = dp[T - 1]
    if not truly_forced:
        valid = last > NEG
        f = int(np.nonzero(valid)[0][np.argmax(last[valid])]) if valid.any() else int(np.argmax(last))
    else:
        f = L - 1
        if last[f] <= NEG and L >= 2:
            f = L - 2
        if last[f] <= NEG:
            valid = last > NEG
            f = int(np.nonzero(valid)[0][-1]) if valid.any() else L - 1
    states = np.zeros(T, np.int64)
    states[T - 1] = f
    for t in range(T - 2, -1, -1):
        states[t] = bp[t + 1, states[t + 1]]        # negative values wrap like python indexing
    return np.asarray(path)[states], states, float(last[f])


@settings(max_examples=60, **SET)
@given(st.integers(2, 40), st.integers(1, 12), st.sampled_from([1, 2, 3, 4]), st.integers(0, 12), st.booleans(), st.integers(0, 10_000),
       st.sampled_from([0.0, 2.0, 8.0]))
def test_oracle_dp_core_vs_numpy_restatement(T, N, stride, band, forced, seed, peak):
    from oracle import oracle as orc
    rng = np.random.default_rng(seed)
    Cc = 9
    blank = Cc - 1
    tgt = rng.integers(0, blank, N)
    if rng.random() < 0.4 and N > 2:
        tgt[1] = tgt[0]                              # repeated phonemes: can_skip off
    path = np.full(stride * N + 1, blank, np.int64); path[1::stride] = tgt
    x = rng.standard_normal((T, Cc)).astype(np.float32)
    if peak:
        x[np.arange(T), path[np.minimum((np.arange(T) * len(path)) // T, len(path) - 1)]] += peak
    lp = (x - np.log(np.exp(x).sum(1, keepdims=True))).astype(np.float32) if peak else np.full((T, Cc), np.float32(-2.0))   # peak 0: every score ties
    tidx = np.full(len(path), -1, np.int64); tidx[1::stride] = np.arange(N)
    r = orc.viterbi(lp, path.astype(np.int32), tidx.astype(np.int32), band, blank, truly_forced=forced)
    ph, states, score = _viterbi_numpy(lp, path, band, blank, forced)
    np.testing.assert_array_equal(r["frame_ph"], ph)
    np.testing.assert_array_equal(r["frame_idx"], tidx[states])
    assert int(r["final_state"]) == int(states[-1])
    assert np.float32(r["dp_final"]) == np.float32(score)


@settings(max_examples=40, **SET)
@given(st.integers(30, 300), st.floats(0.05, 0.24), st.sampled_from([17, 30, 67]), st.sampled_from([0, 4, 7]), st.integers(0, 10_000))
def test_oracle_forced_path_invariants(T, dens, Cc, sil_every, seed):
    """A well-peaked utterance with room for stride 4: every target is visited in order, exactly once per run, and the
    run-length stamps tile the labelled frames."""
    from bfa_b200 import synth
    from oracle import oracle as orc
    N = max(1, int(T * dens))
    lp, tgt, _ = synth.planted_batch(1, T, N, Cc, seed=seed, peak=10.0, sil_every=sil_every, sil_frames=12)
    p = orc.params(Cc - 1, 0)
    r = orc.decode_forced(lp[0].numpy(), tgt[0].numpy(), p)
    assert (r["status"] & 7) in (0, 4)
    ix = r["frame_idx"][r["frame_idx"] >= 0]
    assert (np.diff(ix) >= 0).all()                                   # monotone
    if r["status"] == 0:                                               # one DP problem, truly forced: all targets, in order
        assert sorted(set(ix.tolist())) == list(range(N))             # (a segmented utterance may lose targets: core.py's ensure_target_coverage)
    stamps = orc.assort(r["frame_ph"], r["frame_idx"], Cc - 1, True)
    assert [s[3] for s in stamps] == sorted(s[3] for s in stamps)
    for (ph, s, e, i) in stamps:
        assert (r["frame_ph"][s:e] == ph).all() and (r["frame_idx"][s:e] == i).all() and e > s


@pytest.mark.skipif(not Path("/root/reference/bournemouth_aligner/forced_alignment.py").exists(), reason="the live reference only exists in the build container")
@settings(max_examples=40, **SET)
@given(st.integers(20, 260), st.sampled_from([0.05, 0.1, 0.2, 0.3, 0.5, 0.9, 1.0]), st.sampled_from([17, 66, 67]), st.sampled_from([0, 3, 5, 9]),
       st.sampled_from([3.0, 6.0, 12.0]), st.sampled_from([0, 3, 10]), st.booleans(), st.integers(0, 10_000))
def test_oracle_vs_live_reference(T, dens, Cc, sil_every, peak, anchors, forced, seed):
    import importlib.util
    from bfa_b200 import synth
    from oracle import oracle as orc
    spec = importlib.util.spec_from_file_location("bfa_ref_fa_hyp", "/root/reference/bournemouth_aligner/forced_alignment.py")
    fa = importlib.util.module_from_spec(spec); spec.loader.exec_module(fa)
    torch.set_num_threads(1)
    N = max(1, min(int(T * dens), 250))
    lp, tgt, _ = synth.planted_batch(1, T, N, Cc, seed=seed, peak=peak, sil_every=sil_every, sil_frames=11)
    au = fa.AlignmentUtils(Cc - 1, 0, silence_anchors=anchors, ignore_noise=True, truly_forced=forced)
    p = orc.params(Cc - 1, 0, anchors, True, forced)
    r = orc.decode_forced(lp[0].numpy(), tgt[0].numpy(), p)
    try:
        fp, fi, _ = au.viterbi_decoder.decode_with_forced_alignment(lp[0], tgt[0], anchor_pauses=anchors > 0)
    except ValueError:
        assert (r["status"] & 7) == orc.ORC_TOO_SHORT
        return
    assert (r["status"] & 7) != orc.ORC_TOO_SHORT
    np.testing.assert_array_equal(r["frame_ph"], fp.numpy())
    np.testing.assert_array_equal(r["frame_idx"], fi.numpy())


@pytest.mark.gpu
@settings(max_examples=25, **SET)
@given(st.integers(1, 24), st.sampled_from([9, 17, 30, 66, 67, 72]), st.integers(8, 420), st.sampled_from([0.03, 0.1, 0.2, 0.26, 0.5, 1.0]),
       st.sampled_from([0, 4, 9]), st.sampled_from([0, 3, 10]), st.booleans(), st.booleans(), st.integers(0, 10_000))
def test_cuda_vs_oracle_drawn_batches(bfa, orc, dev, B, Cc, Tmax, dens, sil_every, anchors, boost, simple, seed):
    """Ragged batch, packed rows (row starts on every 16-byte residue), drawn decoder mode: frames, stamps and statuses bit for bit."""
    from bfa_b200 import synth, _cabi
    rng = np.random.default_rng(seed)
    utts = []
    for b in range(B):
        T = int(rng.integers(max(2, Tmax // 4), Tmax + 1))
        N = max(1, min(int(T * dens), 120))
        l, t, _ = synth.planted_batch(1, T, N, Cc, seed=seed * 31 + b, peak=9.0, sil_every=sil_every, sil_frames=12)
        utts.append((l[0], t[0]))
    flat, row_off, Ts, tg, Ns = synth.pack_ragged(utts, Cc, align_floats=1)
    au = bfa.AlignmentUtils(Cc - 1, 0, silence_anchors=anchors)
    dec = au.viterbi_decoder
    p = dec._params(boost, boost, anchors > 0)
    if simple:
        p.mode = _cabi.MODE_SIMPLE
    r = dec.align_batch(flat.to(dev), row_off.to(dev), Ts, Cc, tg.to(dev), Ns, params=p)
    torch.cuda.synchronize()
    po = orc.params(Cc - 1, 0, anchors, True, True)
    po.boost_targets = po.enforce_minimum = int(boost)
    po.mode = p.mode
    toff = np.zeros(B + 1, np.int64); np.cumsum(np.asarray(Ns, np.int64), out=toff[1:])
    o = orc.align_batch(po, flat.numpy(), row_off.numpy(), np.asarray(Ts, np.int32), Cc, tg.numpy(), toff, max_stamps=r.max_stamps, n_threads=2)
    st_g = r.status[:B].cpu().numpy()
    np.testing.assert_array_equal(st_g & 15, o["status"] & 15)
    live = np.repeat((st_g & 7) != 2, np.asarray(Ts))
    n_tot = int(sum(Ts))
    np.testing.assert_array_equal(r.frame_ph.cpu().numpy()[:n_tot][live], o["frame_ph"][live])
    np.testing.assert_array_equal(r.frame_idx.cpu().numpy()[:n_tot][live], o["frame_idx"][live])
    np.testing.assert_array_equal(r.n_stamps[:B].cpu().numpy(), o["n_stamps"])
    for b in range(B):
        n = int(o["n_stamps"][b])
        if n:
            np.testing.assert_array_equal(r.stamps[b, :n, 1].cpu().numpy(), o["stamps"]["start"][b][:n])
            np.testing.assert_allclose(r.conf[b, :n].cpu().numpy(), o["conf"][b, :n], rtol=1e-4, atol=1e-6)


@pytest.mark.gpu
@settings(max_examples=20, **SET)
@given(st.integers(1, 24), st.sampled_from([9, 17, 30, 66, 67, 72, 80]), st.integers(8, 420), st.sampled_from([0.03, 0.1, 0.2, 0.26, 0.5, 1.0]),
       st.sampled_from([0, 4, 9]), st.sampled_from([0, 3, 10]), st.sampled_from([0, 64, 128]), st.integers(0, 10_000))
def test_logits_equal_log_probs_drawn_batches(bfa, dev, B, Cc, Tmax, dens, sil_every, anchors, flags, seed):
    """bfa_align_batch_logits on ragged, packed batches (rows on every 16-byte residue; every stride, T == N, silence layouts,
    with the one-kernel pass alone / skipped / followed by the chain): whatever the call finishes must equal the ordinary call on
    the normalised rows -- statuses, frame labels, stamps bit for bit (an utterance may differ on a last-bit tie of the DP score),
    confidences to 1e-4, row_lse to 3e-5 -- and what it cannot take must come back flagged (or the call refused), never wrong."""
    from bfa_b200 import synth, _cabi
    from bfa_b200._cabi import BfaError
    rng = np.random.default_rng(seed)
    utts = []
    for b in range(B):
        T = int(rng.integers(max(2, Tmax // 4), Tmax + 1))
        N = max(1, min(int(T * dens), 120))
        l, t, _ = synth.planted_batch(1, T, N, Cc, seed=seed * 37 + b, peak=9.0, sil_every=sil_every, sil_frames=12)
        utts.append((l[0], t[0]))
    flat, row_off, Ts, tg, Ns = synth.pack_ragged(utts, Cc, align_floats=1)
    n_tot = int(sum(Ts))
    # per-row shifts (what makes the rows "logits"), applied row by row on the packed buffer
    logit = flat.clone()
    g = torch.Generator().manual_seed(seed)
    for b in range(B):
        o = int(row_off[b])
        rows = logit[o:o + Ts[b] * Cc].view(Ts[b], Cc)
        rows += torch.randn(Ts[b], 1, generator=g) * 5.0 + 1.0
    dec = bfa.AlignmentUtils(Cc - 1, 0, silence_anchors=anchors).viterbi_decoder
    ref = dec.align_batch(flat.to(dev), row_off.to(dev), Ts, Cc, tg.to(dev), Ns, params=dec._params(True, True, anchors > 0))
    p = dec._params(True, True, anchors > 0)
    p.reserved |= flags                                    # 0 | BFA_FLAG_DIRECT_ONLY | BFA_FLAG_NO_DIRECT
    try:
        r = dec.align_batch(logit.to(dev), row_off.to(dev), Ts, Cc, tg.to(dev), Ns, params=p, logits=True)
    except BfaError as e:
        assert e.code == _cabi.BFA_E_UNSUPPORTED           # no pass can take logits for this shape: said so, nothing computed
        return
    torch.cuda.synchronize()
    st_r, st_g = ref.status[:B].cpu().numpy(), r.status[:B].cpu().numpy()
    done = (st_g & 7) != _cabi.ST_DEFERRED
    np.testing.assert_array_equal((st_g & 15)[done], (st_r & 15)[done])
    fo = np.zeros(B + 1, np.int64); np.cumsum(np.asarray(Ts, np.int64), out=fo[1:])
    fph_r, fix_r = ref.frame_ph.cpu().numpy(), ref.frame_idx.cpu().numpy()
    fph_g, fix_g = r.frame_ph.cpu().numpy(), r.frame_idx.cpu().numpy()
    lse = r.row_lse.cpu().numpy()
    ties = 0
    for b in np.nonzero(done)[0]:
        if (st_g[b] & 7) == 2:                             # TOO_SHORT: no frames
            continue
        a, e = int(fo[b]), int(fo[b + 1])
        if not (np.array_equal(fph_g[a:e], fph_r[a:e]) and np.array_equal(fix_g[a:e], fix_r[a:e])):
            ties += 1
            assert abs(float(r.dp_final[b]) - float(ref.dp_final[b])) <= 4e-6 * max(1.0, abs(float(ref.dp_final[b])))
            continue
        n = int(ref.n_stamps[b])
        assert int(r.n_stamps[b]) == n
        if n:
            assert torch.equal(r.stamps[b, :n], ref.stamps[b, :n])
            np.testing.assert_allclose(r.conf[b, :n].cpu().numpy(), ref.conf[b, :n].cpu().numpy(), rtol=1e-4, atol=1e-6)
        if (st_g[b] & 7) in (0, 4) and Ns[b] > 0:          # rows were read by a pass that leaves row_lse
            o = int(row_off[b])
            want = torch.logsumexp(logit[o:o + Ts[b] * Cc].view(Ts[b], Cc).double(), dim=1).numpy()
            np.testing.assert_allclose(lse[a:e], want, rtol=0, atol=3e-5)
    assert ties <= 1
