"""GPU parity at BASELINE.json's full sizes and on the one-kernel (direct) path.

Everything here goes through the C ABI (ViterbiDecoder.align_batch -> bfa_align_batch) and is compared with the C oracle
(oracle/bfa_oracle.c, pinned to the reference) on EVERY utterance: frame labels, timestamps and statuses bit-exact, DP scores
and confidences within 1e-4 relative (the tolerance north_star states; the kernel's fused log-softmax uses ex2/lg2, the
reference torch's expf/logf)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _oracle(orc, p, w, Cc, max_stamps, threads=None):
    Ts, Ns = w["Ts"], w["Ns"]
    toff = np.zeros(len(Ns) + 1, np.int64); np.cumsum(np.asarray(Ns, np.int64), out=toff[1:])
    return orc.align_batch(p, w["lp"].cpu().numpy(), w["row_off"].cpu().numpy(), np.asarray(Ts, np.int32), Cc, w["tgt"].cpu().numpy(), toff,
                           max_stamps=max_stamps, n_threads=threads or orc.n_host_threads())


def _compare_all(r, o, B, what, check_score=True):
    """Every utterance of a BatchResult against the oracle's arrays.  Returns the statuses."""
    st = r.status[:B].cpu().numpy()
    np.testing.assert_array_equal(st & 15, o["status"] & 15, err_msg=f"{what}: statuses")
    fph, fix = r.frame_ph.cpu().numpy(), r.frame_idx.cpu().numpy()
    n_tot = int(o["frame_off"][-1])
    live = np.repeat((st & 7) != 2, np.diff(o["frame_off"]))          # TOO_SHORT utterances have no frames
    dph = (fph[:n_tot] != o["frame_ph"]) & live
    dix = (fix[:n_tot] != o["frame_idx"]) & live
    if dph.any() or dix.any():
        bad = np.unique(np.searchsorted(o["frame_off"], np.nonzero(dph | dix)[0], side="right") - 1)
        raise AssertionError(f"{what}: {len(bad)} of {B} utterances differ from the oracle in their frame labels, first {bad[:8].tolist()}")
    nst = r.n_stamps[:B].cpu().numpy()
    np.testing.assert_array_equal(nst, o["n_stamps"], err_msg=f"{what}: stamp counts")
    stamps = r.stamps.cpu().numpy()[:B]; conf = r.conf.cpu().numpy()[:B]
    ms = min(stamps.shape[1], o["stamps"].shape[1])
    valid = np.arange(ms)[None, :] < nst[:, None]
    for i, f in enumerate(("phoneme", "start", "end", "target_idx")):
        assert np.array_equal(stamps[:, :ms, i][valid], o["stamps"][f][:, :ms][valid]), f"{what}: stamp field {f}"
    np.testing.assert_allclose(conf[:, :ms][valid], o["conf"][:, :ms][valid], rtol=RTOL, atol=1e-6, err_msg=f"{what}: confidences")
    if check_score:
        unseg = (st & 15) == 0
        np.testing.assert_allclose(r.dp_final[:B].cpu().numpy()[unseg], o["dp_final"][unseg], rtol=RTOL, err_msg=f"{what}: DP scores")
    return st


def _align(bfa, dev, w, Cc, flags=0, **kw):
    from bfa_b200 import _cabi  # noqa: F401
    au = bfa.AlignmentUtils(Cc - 1, 0, **kw)
    dec = au.viterbi_decoder
    p = dec._params(True, True, au.silence_anchors > 0)
    p.reserved |= flags
    r = dec.align_batch(w["lp"].to(dev), w["row_off"].to(dev), w["Ts"], Cc, w["tgt"].to(dev), w["Ns"], params=p)
    torch.cuda.synchronize()
    return r


# ---- the metric batch and BASELINE configs 2-4 at their full sizes, every utterance ------------------------------
def test_metric_batch_every_utterance_vs_oracle(bfa, orc, dev):
    """B=4096, T=600, N=40, C=66 (the metric shape): all 4096 utterances, default chain and one-kernel path."""
    from bfa_b200 import synth, _cabi
    B, T, N, Cc = 4096, 600, 40, 66
    lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=77)
    w = dict(lp=lp.reshape(-1), row_off=torch.arange(B, dtype=torch.int64) * T * Cc, Ts=[T] * B, tgt=tgt.to(torch.int32).reshape(-1), Ns=[N] * B)
    o = _oracle(orc, orc.params(Cc - 1, 0), w, Cc, N + 8)
    for flags in (0, _cabi.HINT_NO_SIL | _cabi.FLAG_DIRECT_ONLY, _cabi.HINT_NO_SIL | _cabi.FLAG_DIRECT_ONLY | _cabi.FLAG_PIPELINED):
        r = _align(bfa, dev, w, Cc, flags)
        st = _compare_all(r, o, B, f"metric batch, flags {flags}")
        assert (st == 0).all()


def test_baseline_config2_full_vs_oracle(bfa, orc, dev):
    from bfa_b200 import synth
    w = synth.baseline_config(2)
    o = _oracle(orc, orc.params(65, 0), w, 66, 48)
    st = _compare_all(_align(bfa, dev, w, 66), o, 1024, "config 2")
    assert (st == 0).all()


def test_baseline_config3_full_vs_oracle(bfa, orc, dev):
    """B=256, T=3600, N=200 with SIL anchors: silence-anchored segmentation on the device, all 256 utterances."""
    from bfa_b200 import synth
    w = synth.baseline_config(3)
    o = _oracle(orc, orc.params(65, 0), w, 66, 208)
    st = _compare_all(_align(bfa, dev, w, 66), o, 256, "config 3")
    assert ((st & 7) == 4).sum() >= 240          # segmentation accepted nearly everywhere


def test_baseline_config4_full_vs_oracle(bfa, orc, dev):
    """Ragged B=8192, T in [60,1800], N in [4,120], packed rows: all utterances (every stride, both window kernels,
    the exact kernel for the dense tail)."""
    from bfa_b200 import synth
    w = synth.baseline_config(4, device=dev)
    w_cpu = dict(w, lp=w["lp"].cpu(), row_off=w["row_off"].cpu(), tgt=w["tgt"].cpu())
    o = _oracle(orc, orc.params(65, 0), w_cpu, 66, max(w["Ns"]) + 8)
    _compare_all(_align(bfa, dev, w, 66), o, 8192, "config 4")


# ---- the one-kernel path ---------------------------------------------------------------------------------------
def _mixed_batch(synth, Cc):
    """Plain utterances mixed with everything the direct kernel must hand back: silence_id in the target, too dense for
    stride 4, T == N, an empty target, more than 128 phonemes, a band wider than the 24-group window."""
    specs = [(400, 30, 0), (400, 30, 8), (400, 120, 0), (77, 77, 0), (300, 0, 0), (900, 150, 0), (1000, 110, 0), (64, 5, 0),
             (401, 31, 0), (399, 29, 0), (200, 49, 0), (200, 50, 0)] * 3
    utts = []
    for i, (T, N, sil) in enumerate(specs):
        if N == 0:
            l, _, _ = synth.planted_batch(1, T, 4, Cc, seed=900 + i, peak=9.0)
            utts.append((l[0], torch.zeros(0, dtype=torch.long)))
            continue
        l, t, _ = synth.planted_batch(1, T, N, Cc, seed=900 + i, peak=9.0, sil_every=sil, sil_frames=14)
        utts.append((l[0], t[0]))
    return utts


def test_direct_only_flags_what_it_cannot_finish(bfa, orc, dev):
    """BFA_FLAG_DIRECT_ONLY: an utterance is either finished exactly like the full chain finishes it, or reported as
    BFA_ST_DEFERRED with no stamps -- never silently different."""
    from bfa_b200 import synth, _cabi
    Cc = 67
    utts = _mixed_batch(synth, Cc)
    flat, row_off, Ts, tg, Ns = synth.pack_ragged(utts, Cc)
    w = dict(lp=flat, row_off=row_off, Ts=Ts, tgt=tg, Ns=Ns)
    B = len(Ts)
    full = _align(bfa, dev, w, Cc, 0)
    only = _align(bfa, dev, w, Cc, _cabi.FLAG_DIRECT_ONLY)
    st_f, st_o = full.status[:B].cpu().numpy(), only.status[:B].cpu().numpy()
    deferred = (st_o & 7) == _cabi.ST_DEFERRED
    assert deferred.any() and (~deferred).any()
    # what must have been handed back: SIL targets (a segmentation attempt), dense / proportional / empty / long targets
    for b in range(B):
        T, N = Ts[b], Ns[b]
        must_defer = N == 0 or 4 * N + 1 > T or N > 128 or bool((utts[b][1] == 0).any())
        if must_defer:
            assert deferred[b], f"utterance {b} (T={T}, N={N}) should have been deferred"
    fo = np.zeros(B + 1, np.int64); np.cumsum(np.asarray(Ts, np.int64), out=fo[1:])
    for b in np.nonzero(~deferred)[0]:
        assert st_o[b] == st_f[b] == 0
        for x, y in ((only.frame_ph, full.frame_ph), (only.frame_idx, full.frame_idx)):
            assert torch.equal(x[fo[b]:fo[b + 1]], y[fo[b]:fo[b + 1]]), f"utterance {b}: frames"
        n = int(full.n_stamps[b])
        assert int(only.n_stamps[b]) == n
        assert torch.equal(only.stamps[b, :n], full.stamps[b, :n])
        torch.testing.assert_close(only.conf[b, :n], full.conf[b, :n], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(only.dp_final[b], full.dp_final[b], rtol=1e-5, atol=0)
    assert (only.n_stamps[:B].cpu().numpy()[deferred] == 0).all()
    # and the full chain (direct kernel first, planner chain for the rest) equals the oracle on everything
    toff = np.zeros(B + 1, np.int64); np.cumsum(np.asarray(Ns, np.int64), out=toff[1:])
    o = orc.align_batch(orc.params(Cc - 1, 0), flat.numpy(), row_off.numpy(), np.asarray(Ts, np.int32), Cc, tg.numpy(), toff,
                        max_stamps=full.max_stamps, n_threads=4)
    _compare_all(full, o, B, "mixed batch, full chain")
    # A/B switch: without the direct kernel the same
    _compare_all(_align(bfa, dev, w, Cc, _cabi.FLAG_NO_DIRECT), o, B, "mixed batch, no direct kernel")


def test_facade_reruns_deferred_utterances(bfa, orc, dev):
    """decode_alignments with host-side targets picks the one-kernel path when it can and silently falls back to the full
    chain when that path hands utterances back (here: band-edge paths are unlikely, so a dense utterance is slipped in)."""
    from bfa_b200 import synth, _cabi
    Cc, B, T, N = 67, 24, 300, 30
    lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=31, peak=9.0)
    au = bfa.AlignmentUtils(Cc - 1, 0)
    lens, nl = torch.full((B,), T), torch.full((B,), N)
    got = au.decode_alignments(lp.to(dev), true_seqs=tgt, pred_lens=lens, true_seqs_lens=nl, with_confidence=True)
    assert au.last_params_reserved & _cabi.FLAG_DIRECT_ONLY
    o = orc.align_batch(orc.params(Cc - 1, 0), lp.numpy(), np.arange(B, dtype=np.int64) * T * Cc, np.full(B, T, np.int32), Cc,
                        tgt.numpy().astype(np.int32).reshape(-1), np.arange(B + 1, dtype=np.int64) * N, max_stamps=T, n_threads=4)
    for b in range(B):
        n = int(o["n_stamps"][b])
        want = [tuple(int(o["stamps"][b][f][i]) for f in ("phoneme", "start", "end", "target_idx")) for i in range(n)]
        assert [g[:4] for g in got[b]] == want
        np.testing.assert_allclose([g[4] for g in got[b]], o["conf"][b][:n], rtol=RTOL, atol=1e-6)
    # shorter audio for one utterance makes it too dense for stride 4: the facade's host-side check then keeps the full chain
    lens2 = lens.clone(); lens2[3] = 100
    got2 = au.decode_alignments(lp.to(dev), true_seqs=tgt, pred_lens=lens2, true_seqs_lens=nl)
    assert not (au.last_params_reserved & _cabi.FLAG_DIRECT_ONLY)
    o2 = orc.align_batch(orc.params(Cc - 1, 0), lp.numpy(), np.arange(B, dtype=np.int64) * T * Cc, lens2.numpy().astype(np.int32), Cc,
                         tgt.numpy().astype(np.int32).reshape(-1), np.arange(B + 1, dtype=np.int64) * N, max_stamps=T, n_threads=4)
    for b in range(B):
        n = int(o2["n_stamps"][b])
        assert got2[b] == [tuple(int(o2["stamps"][b][f][i]) for f in ("phoneme", "start", "end", "target_idx")) for i in range(n)]


def test_pipelined_calls_back_to_back(bfa, orc, dev):
    """BFA_FLAG_PIPELINED: consecutive launches overlap (programmatic dependent launch, no wait before the fill, per-SM
    scratch).  Twelve back-to-back calls over three different batches into two alternating result sets must each give what
    an ordinary call gives."""
    from bfa_b200 import synth, _cabi
    Cc, B, T, N = 66, 1200, 320, 24
    batches = []
    for s in range(3):
        lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=40 + s, device=dev)
        batches.append((lp, tgt.to(torch.int32).reshape(-1).contiguous()))
    dec = bfa.AlignmentUtils(Cc - 1, 0).viterbi_decoder
    row_off = torch.arange(B, dtype=torch.int64, device=dev) * T * Cc
    p0 = dec._params(True, True, True)
    ref = []
    for lp, tg in batches:
        r = dec.align_batch(lp, row_off, [T] * B, Cc, tg, [N] * B, params=p0)
        assert (r.n_stamps[:B] == N).all()
        ref.append((r.frame_ph.clone(), r.frame_idx.clone(), r.stamps[:B, :N].clone(), r.conf[:B, :N].clone(), r.dp_final[:B].clone()))
    p = dec._params(True, True, True)
    p.reserved |= _cabi.HINT_NO_SIL | _cabi.FLAG_DIRECT_ONLY | _cabi.FLAG_PIPELINED
    plan = dec.plan_batch([T] * B, [N] * B, Cc, params=p, device=dev)
    outs = [None, None]
    torch.cuda.synchronize()
    checks = []
    for i in range(12):
        k = i % 3
        lp, tg = batches[k]
        outs[i & 1] = dec.align_batch(lp, row_off, [T] * B, Cc, tg, [N] * B, params=p, plan=plan, out=outs[i & 1])
        if i >= 10:     # the last two calls are checked (their result sets are not overwritten afterwards)
            checks.append((i & 1, k))
    torch.cuda.synchronize()
    for slot, k in checks:
        r = outs[slot]
        assert (r.status[:B] == 0).all()
        assert torch.equal(r.frame_ph, ref[k][0]) and torch.equal(r.frame_idx, ref[k][1])
        # conf / dp_final come from the same code on the same data: bit-identical too
        assert (r.n_stamps[:B] == N).all()
        assert torch.equal(r.stamps[:B, :N], ref[k][2]) and torch.equal(r.conf[:B, :N], ref[k][3]) and torch.equal(r.dp_final[:B], ref[k][4])


def test_caller_owned_arena(bfa, dev):
    """align_batch(arena=...): the packed per-utterance results go into storage the caller owns (the multi-GPU path hands in this
    rank's slot of the gathering rank's symmetric-memory buffer, sharding.PeerArena).  Same results as a library-allocated arena,
    nothing written outside the arena's words, bad arenas refused."""
    from bfa_b200 import synth
    from bfa_b200.aligner import result_arena_words, BfaError
    Cc, B, T, N = 66, 300, 200, 16
    lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=77, device=dev)
    tg = tgt.to(torch.int32).reshape(-1).contiguous()
    dec = bfa.AlignmentUtils(Cc - 1, 0).viterbi_decoder
    row_off = torch.arange(B, dtype=torch.int64, device=dev) * T * Cc
    p = dec._params(True, True, True)
    ref = dec.align_batch(lp, row_off, [T] * B, Cc, tg, [N] * B, params=p)
    words = result_arena_words(B, ref.max_stamps, True, True)["total"]
    big = torch.full((words + 64,), 0x5a5a5a5a, dtype=torch.int32, device=dev)
    r = dec.align_batch(lp, row_off, [T] * B, Cc, tg, [N] * B, params=p, arena=big[32:32 + words + 8])
    torch.cuda.synchronize()
    assert r.arena.data_ptr() == big[32:].data_ptr() and r.arena.numel() == words
    assert (big[:32] == 0x5a5a5a5a).all() and (big[32 + words:] == 0x5a5a5a5a).all()
    assert torch.equal(r.frame_ph, ref.frame_ph) and torch.equal(r.n_stamps[:B], ref.n_stamps[:B]) and torch.equal(r.status[:B], ref.status[:B])
    assert torch.equal(r.stamps[:B, :N], ref.stamps[:B, :N]) and torch.equal(r.conf[:B, :N], ref.conf[:B, :N])
    assert torch.equal(r.dp_final[:B], ref.dp_final[:B])
    r2 = dec.align_batch(lp, row_off, [T] * B, Cc, tg, [N] * B, params=p, out=r)      # reuse keeps writing into the caller's storage
    assert r2.arena.data_ptr() == big[32:].data_ptr()
    for bad in (big[33:33 + words], big[32:32 + words - 4], big[32:32 + words].to(torch.int64), torch.zeros(words, dtype=torch.int32)):
        with pytest.raises(BfaError):
            dec.align_batch(lp, row_off, [T] * B, Cc, tg, [N] * B, params=p, arena=bad)


@pytest.mark.parametrize("Cc", [66, 67, 17, 30])
def test_logits_in_one_kernel(bfa, orc, dev, Cc):
    """bfa_align_batch_logits: un-normalised logits in (core.py:898-899 skipped).  Frames, timestamps and statuses must equal what
    the ordinary call gives on log_softmax(logits) -- and the oracle on the same log-probabilities --, confidences and the DP
    score agree to 1e-4, row_lse is the rows' log-sum-exp; an utterance the one-kernel pass cannot take goes through the planner
    chain (still on the logits), or is flagged DEFERRED under BFA_FLAG_DIRECT_ONLY."""
    from bfa_b200 import synth, _cabi
    B, T, N = 512, 300, 24
    lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=900 + Cc, peak=7.0, device=dev)
    g = torch.Generator(device="cpu").manual_seed(5)
    shift = (torch.randn(B, T, 1, generator=g) * 6.0 + 3.0).to(dev)        # per-row offsets: what makes log-probs "logits"
    logits = (lp + shift).contiguous()
    tgt = tgt.clone()
    tgt[3, 5] = 0                                                           # silence_id in a target while anchoring is on -> not for this pass
    tg = tgt.to(torch.int32).reshape(-1).contiguous()
    dec = bfa.AlignmentUtils(Cc - 1, 0, silence_anchors=10, ignore_noise=True, truly_forced=True).viterbi_decoder
    row_off = torch.arange(B, dtype=torch.int64, device=dev) * T * Cc
    p = dec._params(True, True, True)
    ref = dec.align_batch(lp, row_off, [T] * B, Cc, tg, [N] * B, params=p)
    r = dec.align_batch(logits, row_off, [T] * B, Cc, tg, [N] * B, params=dec._params(True, True, True), logits=True)
    torch.cuda.synchronize()
    # utterance 3 goes to the planner chain -- on the logits as well: the chain's silence pass supplies its row_lse
    assert torch.equal(r.status[:B] & 7, ref.status[:B] & 7) and int(r.status[3] & 7) != _cabi.ST_DEFERRED
    pd = dec._params(True, True, True)
    pd.reserved |= _cabi.FLAG_DIRECT_ONLY              # the one-kernel pass alone: what it cannot take is flagged
    rd = dec.align_batch(logits, row_off, [T] * B, Cc, tg, [N] * B, params=pd, logits=True)
    st = rd.status[:B].cpu().numpy()
    assert (st[3] & 7) == _cabi.ST_DEFERRED and int(((st & 7) == _cabi.ST_DEFERRED).sum()) == 1
    others = torch.ones(B, dtype=torch.bool, device=dev); others[3] = False
    assert torch.equal(rd.frame_ph[: B * T][others.repeat_interleave(T)], r.frame_ph[: B * T][others.repeat_interleave(T)])
    assert torch.equal(rd.conf[:B, :N][others], r.conf[:B, :N][others])
    keep = torch.ones(B, dtype=torch.bool, device=dev)
    # the two runs see emissions that differ in the last bit (x - lse against (x + s) - (lse + s)): an utterance may differ where
    # two paths tie to 1e-6 of the score; anything else must be identical
    same = ((r.frame_ph[: B * T] == ref.frame_ph[: B * T]) & (r.frame_idx[: B * T] == ref.frame_idx[: B * T])).view(B, T).all(1)
    tied = keep & ~same
    assert int(tied.sum()) <= 2, int(tied.sum())
    assert torch.allclose(r.dp_final[:B][tied], ref.dp_final[:B][tied], rtol=2e-6, atol=0)
    keep = keep & same
    fk = keep.repeat_interleave(T)
    assert torch.equal(r.frame_ph[: B * T][fk], ref.frame_ph[: B * T][fk]) and torch.equal(r.frame_idx[: B * T][fk], ref.frame_idx[: B * T][fk])
    assert torch.equal(r.n_stamps[:B][keep], ref.n_stamps[:B][keep])
    assert torch.equal(r.stamps[:B, :N][keep], ref.stamps[:B, :N][keep])
    assert torch.allclose(r.conf[:B, :N][keep], ref.conf[:B, :N][keep], rtol=1e-4, atol=1e-6)
    assert torch.allclose(r.dp_final[:B][keep], ref.dp_final[:B][keep], rtol=1e-4, atol=1e-3)
    lse = torch.logsumexp(logits.double(), dim=2).float().reshape(-1)
    assert torch.allclose(r.row_lse[: B * T][fk], lse[fk], rtol=0, atol=2e-5)
    # and against the oracle on the normalised rows (a sample)
    po = orc.params(Cc - 1, 0)
    import numpy as np
    for u in [u for u in (0, 100, 511) if bool(keep[u])]:
        o = orc.align_batch(po, lp[u].cpu().numpy().reshape(1, T, Cc), np.zeros(1, np.int64), np.asarray([T], np.int32), Cc,
                            tgt[u].cpu().numpy().astype(np.int32), np.asarray([0, N], np.int64), max_stamps=r.max_stamps, n_threads=1)
        assert np.array_equal(r.frame_ph[u * T:(u + 1) * T].cpu().numpy(), o["frame_ph"])
        assert np.allclose(r.conf[u, :N].cpu().numpy(), o["conf"][0, :N], rtol=1e-4, atol=1e-6)
    # no boosting: nothing re-normalises the rows -> refused
    with pytest.raises(Exception):
        dec.align_batch(logits, row_off, [T] * B, Cc, tg, [N] * B, params=dec._params(False, True, True), logits=True)


@pytest.mark.parametrize("Cc,sil_every", [(66, 10), (67, 9), (17, 5), (80, 10)])
def test_logits_through_the_planner_chain(bfa, dev, Cc, sil_every):
    """Targets with silence_id (what punctuation does): silence-anchored segmentation, list-mode banded kernel / exact kernel and
    the stamp kernel all run on the un-normalised logits, the silence pass leaves row_lse.  Everything must equal the same call on
    log_softmax(logits).  C = 80 takes the paths without the staged row reduction (class table of 72) and without banded kernels."""
    from bfa_b200 import synth, _cabi
    B, T = 384, 600
    N = 40 if Cc > 20 else 14
    lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=1200 + Cc, peak=9.0, sil_every=sil_every, sil_frames=15, device=dev)
    tgt[::7, 1] = 1                                   # (some variety: utterances stay SIL-bearing)
    shift = (torch.randn(B, T, 1, generator=torch.Generator().manual_seed(2)) * 7.0 - 4.0).to(dev)
    logits = (lp + shift).contiguous()
    tg = tgt.to(torch.int32).reshape(-1).contiguous()
    dec = bfa.AlignmentUtils(Cc - 1, 0, silence_anchors=10, ignore_noise=True, truly_forced=True).viterbi_decoder
    row_off = torch.arange(B, dtype=torch.int64, device=dev) * T * Cc
    ref = dec.align_batch(lp, row_off, [T] * B, Cc, tg, [N] * B, params=dec._params(True, True, True))
    for flags in (0, _cabi.FLAG_NO_DIRECT):
        p = dec._params(True, True, True)
        p.reserved |= flags
        r = dec.align_batch(logits, row_off, [T] * B, Cc, tg, [N] * B, params=p, logits=True)
        torch.cuda.synchronize()
        assert torch.equal(r.status[:B] & 7, ref.status[:B] & 7)
        assert int(((r.status[:B] & 7) == _cabi.ST_SEGMENTED).sum()) > B // 2          # anchoring really ran
        same = ((r.frame_ph[: B * T] == ref.frame_ph[: B * T]) & (r.frame_idx[: B * T] == ref.frame_idx[: B * T])).view(B, T).all(1)
        assert int((~same).sum()) <= 2, int((~same).sum())                              # last-bit ties, see test_logits_in_one_kernel
        ns = ref.n_stamps[:B]
        assert torch.equal(r.n_stamps[:B][same], ns[same])
        ms = r.stamps.shape[1]
        valid = (torch.arange(ms, device=dev)[None, :] < ns[:, None]) & same[:, None]
        assert torch.equal(r.stamps[:B][valid], ref.stamps[:B][valid])
        assert torch.allclose(r.conf[:B][valid], ref.conf[:B][valid], rtol=1e-4, atol=1e-6)
        lse = torch.logsumexp(logits.double(), dim=2).float().reshape(-1)
        assert torch.allclose(r.row_lse[: B * T], lse, rtol=0, atol=3e-5)


def test_logits_fall_back_when_no_pass_can_take_them(bfa, dev):
    """80 classes (beyond the banded kernels' class table) and no silence_id in the batch: neither the one-kernel pass nor a silence
    pass exists for this shape; the C entry says BFA_E_UNSUPPORTED and the facade normalises first.  Same lists either way."""
    from bfa_b200 import synth
    Cc, B, T, N = 80, 24, 200, 12
    lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=5, peak=8.0, device=dev)
    logits = (lp + 2.5).contiguous()
    au = bfa.AlignmentUtils(blank_id=Cc - 1, silence_id=0, silence_anchors=10, ignore_noise=True, truly_forced=True)
    lens_t = torch.full((B,), T); lens_n = torch.full((B,), N)
    want = au.decode_alignments(lp, true_seqs=tgt.cpu(), pred_lens=lens_t, true_seqs_lens=lens_n, with_confidence=True)
    got = au.decode_alignments(logits, true_seqs=tgt.cpu(), pred_lens=lens_t, true_seqs_lens=lens_n, with_confidence=True, input_is_logits=True)
    assert au.last_row_lse is None and au.last_log_probs is not None
    for b in range(B):
        assert [x[:4] for x in got[b]] == [x[:4] for x in want[b]]
        assert all(abs(x[4] - y[4]) <= 1e-4 * max(abs(y[4]), 1e-3) for x, y in zip(got[b], want[b]))


def test_decode_alignments_from_logits(bfa, dev):
    """AlignmentUtils.decode_alignments(input_is_logits=True): the caller skips F.log_softmax (core.py:898-899).  Same lists as on
    the normalised tensor -- through the one-kernel pass when the batch qualifies, through the planner chain on the logits when targets
    hold silence_id (the silence pass reads every row anyway and leaves row_lse), through a normalising pass otherwise."""
    from bfa_b200 import synth, _cabi
    Cc, B, T, N = 67, 96, 260, 20
    for case in ("plain", "sil", "one deferred", "handed back"):
        lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=31, peak=8.0, sil_every=7 if case == "sil" else 0, sil_frames=14, device=dev)
        tgt = tgt.cpu()
        lens_t = torch.full((B,), T); lens_n = torch.full((B,), N)
        if case == "one deferred":
            lens_t[5] = 60            # 4 N + 1 > T: stride 3 -- the host-side check sends the batch to the full chain
        if case == "handed back":
            tgt[5, 3] = Cc - 1        # blank_id inside a target: the host-side check passes, the kernel hands the utterance back
        logits = (lp + (torch.randn(B, T, 1, generator=torch.Generator().manual_seed(9)) * 5.0).to(dev)).contiguous()
        au = bfa.AlignmentUtils(blank_id=Cc - 1, silence_id=0, silence_anchors=10, ignore_noise=True, truly_forced=True)
        want = au.decode_alignments(lp, true_seqs=tgt, pred_lens=lens_t, true_seqs_lens=lens_n, with_confidence=True)
        got = au.decode_alignments(logits, true_seqs=tgt, pred_lens=lens_t, true_seqs_lens=lens_n, with_confidence=True, input_is_logits=True)
        one_kernel = bool(au.last_params_reserved & _cabi.FLAG_DIRECT_ONLY)
        assert one_kernel == (case == "plain"), (case, au.last_params_reserved)
        # "sil" goes through the planner chain ON THE LOGITS (its silence pass supplies row_lse); "one deferred" (no silence_id
        # anywhere: that pass is not run) is normalised first, "handed back" started as a one-kernel call and is repeated on
        # normalised rows
        from_logits = case in ("plain", "sil")
        assert (au.last_row_lse is not None) == from_logits and (au.last_log_probs is None) == from_logits
        if from_logits:
            lse = torch.logsumexp(logits.double(), dim=2).float().reshape(-1)
            assert torch.allclose(au.last_row_lse[: B * T], lse, rtol=0, atol=2e-5)
        else:
            assert torch.allclose(au.last_log_probs, torch.log_softmax(logits, dim=2), rtol=0, atol=2e-5)
        assert len(got) == B
        for b in range(B):
            assert [x[:4] for x in got[b]] == [x[:4] for x in want[b]], (case, b)
            assert all(abs(x[4] - y[4]) <= 1e-4 * max(abs(y[4]), 1e-3) for x, y in zip(got[b], want[b])), (case, b)


def test_downstream_steps_from_logits(bfa, dev):
    """Soft boundaries and confidences on un-normalised logits + row_lse (bfa_*_batch_lse) equal the same steps on
    log_softmax(logits): the chain core.py:925-937 runs after decode_alignments, without ever materialising the log-probabilities."""
    from bfa_b200 import synth
    Cc, B, T, N = 67, 40, 300, 22
    lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=77, peak=6.0, device=dev)
    logits = (lp + (torch.randn(B, T, 1, generator=torch.Generator().manual_seed(3)) * 4.0 - 2.0).to(dev)).contiguous()
    lens_t = torch.full((B,), T); lens_n = torch.full((B,), N)
    au = bfa.AlignmentUtils(blank_id=Cc - 1, silence_id=0, silence_anchors=10, ignore_noise=True, truly_forced=True)
    got = au.decode_alignments(logits, true_seqs=tgt.cpu(), pred_lens=lens_t, true_seqs_lens=lens_n, input_is_logits=True)
    rl = au.last_row_lse
    assert rl is not None
    st5 = [[(p, s, e, i, False) for (p, s, e, i) in fs] for fs in got]
    soft_a = bfa.extend_soft_boundaries_func(logits, st5, boundary_softness=3, row_lse=rl, pred_lens=lens_t)
    soft_b = bfa.extend_soft_boundaries_func(lp, st5, boundary_softness=3)
    assert soft_a == soft_b and soft_a != st5          # identical, and the step did move boundaries
    conf_a = bfa._calculate_confidences_batch(logits, soft_a, pred_lens=lens_t, row_lse=rl)
    conf_b = bfa._calculate_confidences_batch(lp, soft_b, pred_lens=lens_t)
    for fa, fb in zip(conf_a, conf_b):
        assert [x[:5] for x in fa] == [x[:5] for x in fb]
        assert all(abs(x[5] - y[5]) <= 1e-4 * max(abs(y[5]), 1e-3) for x, y in zip(fa, fb))


# ---- near-ties: how often does the fused log-softmax flip a back-trace decision? -------------------------------
@pytest.mark.parametrize("Cc,sil", [(66, 0), (67, 0), (17, 0), (67, 9)])
def test_flip_rate_at_low_peaks(bfa, orc, dev, Cc, sil):
    """>= 10^4 utterances at planted peaks 2..5 (posteriors far less decisive than the benchmark's): frame labels must be
    bit-identical to the oracle's.  The kernel's emissions come from ex2.approx / lg2.approx and (1b)/(1c) of the reduced
    lattice read ties off decision bits, so a flip is conceivable where two candidates round to the same float; this test
    measures it.  Any utterance that differs must at least carry the same DP score to 1e-6 relative (an exact tie broken
    the other way) -- and is reported."""
    from bfa_b200 import synth
    B, T = 2560, 240
    N = 24 if Cc > 20 else 12
    flips, total = [], 0
    for peak in (2.0, 3.0, 4.0, 5.0):
        lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=int(peak * 100) + Cc, peak=peak, sil_every=sil, sil_frames=14)
        w = dict(lp=lp.reshape(-1), row_off=torch.arange(B, dtype=torch.int64) * T * Cc, Ts=[T] * B, tgt=tgt.to(torch.int32).reshape(-1), Ns=[N] * B)
        o = _oracle(orc, orc.params(Cc - 1, 0), w, Cc, N + 8)
        r = _align(bfa, dev, w, Cc)
        fph = r.frame_ph.cpu().numpy().reshape(B, T); fix = r.frame_idx.cpu().numpy().reshape(B, T)
        diff = ((fph != o["frame_ph"].reshape(B, T)) | (fix != o["frame_idx"].reshape(B, T))).any(1)
        st = r.status[:B].cpu().numpy()
        np.testing.assert_array_equal(st & 15, o["status"] & 15)
        total += B
        for b in np.nonzero(diff)[0]:
            gap = abs(float(r.dp_final[b]) - float(o["dp_final"][b])) / max(abs(float(o["dp_final"][b])), 1e-9)
            flips.append((peak, int(b), gap, int(st[b])))
    msg = f"C={Cc} sil_every={sil}: {len(flips)} of {total} utterances differ from the oracle: {flips[:10]}"
    assert not flips, msg


# ---- API corners the reference has and round 1 left untested ------------------------------------------------------
def test_return_scores_matches_oracle(bfa, orc, dev):
    """decode_with_forced_alignment(return_scores=True) (forced_alignment.py:195-197, :767-773): sum of the ORIGINAL
    log-probs along the returned path."""
    from bfa_b200 import synth
    Cc, T, N = 67, 300, 25
    lp, tgt, _ = synth.planted_batch(1, T, N, Cc, seed=61, peak=9.0)
    dec = bfa.ViterbiDecoder(Cc - 1, 0, silence_anchors=10, truly_forced=True)
    fp, fi, score = dec.decode_with_forced_alignment(lp[0].to(dev), tgt[0], return_scores=True)
    o = orc.align_batch(orc.params(Cc - 1, 0), lp.numpy(), np.zeros(1, np.int64), np.asarray([T], np.int32), Cc, tgt.numpy().astype(np.int32).reshape(-1),
                        np.asarray([0, N], np.int64), max_stamps=T, n_threads=1)
    np.testing.assert_array_equal(fp.cpu().numpy(), o["frame_ph"])
    want = float(lp[0].double()[torch.arange(T), torch.from_numpy(o["frame_ph"]).long()].sum())
    assert isinstance(score, float) and abs(score - want) <= 1e-5 * abs(want)


def test_free_decoding_matches_reference_semantics(bfa, dev):
    """decode_alignments(forced_alignment=False) (forced_alignment.py:912-928): frame-wise argmax, then assort_frames per item,
    ONE flat list for the whole batch like the reference returns."""
    g = torch.Generator().manual_seed(3)
    B, T, Cc = 3, 50, 12
    lp = torch.log_softmax(torch.randn(B, T, Cc, generator=g) * 3.0, -1)
    lens = torch.tensor([50, 37, 44])
    au = bfa.AlignmentUtils(Cc - 1, 0)
    got = au.decode_alignments(lp.to(dev), pred_lens=lens, forced_alignment=False)
    want = []
    for b in range(B):
        pred = lp[b, :int(lens[b])].argmax(1).tolist()
        t = 0
        while t < len(pred):
            e = t
            while e < len(pred) and pred[e] == pred[t]:
                e += 1
            if pred[t] != Cc - 1:           # ignore_noise: blank runs never become stamps
                want.append((pred[t], t, e, -1))
            t = e
    assert got == want


# ---- round-1 advisor findings --------------------------------------------------------------------------------------
def test_long_targets_wide_kernel_and_per_utterance_refusal(bfa, orc, dev):
    """Targets of more than 255 phonemes need more than 1024 path states: they are aligned by the one-CTA-per-problem exact kernel
    (the reference aligns them whole when silence anchoring does not split them, forced_alignment.py:298-308, :153-190).  Beyond
    BFA_MAX_L = 8192 states an utterance is reported as BFA_ST_UNSUPPORTED (blank frames, no stamps) and every other utterance of
    the batch is aligned as usual."""
    from bfa_b200 import synth, _cabi
    Cc = 67
    utts = []
    shapes = [(300, 30), (1900, 350), (500, 60), (1300, 260), (200, 12), (8600, 2100), (2500, 600), (1105, 276)]
    for i, (T, N) in enumerate(shapes):
        l, t, _ = synth.planted_batch(1, T, N, Cc, seed=70 + i, peak=10.0)
        utts.append((l[0], t[0]))
    flat, row_off, Ts, tg, Ns = synth.pack_ragged(utts, Cc)
    w = dict(lp=flat, row_off=row_off, Ts=Ts, tgt=tg, Ns=Ns)
    r = _align(bfa, dev, w, Cc)
    nb = len(shapes)
    st = r.status[:nb].cpu().numpy()
    assert (st & 7).tolist() == [0, 0, 0, 0, 0, _cabi.ST_UNSUPPORTED, 0, 0]
    assert r.n_stamps[:nb].cpu().tolist()[5] == 0
    fo = np.zeros(nb + 1, np.int64); np.cumsum(np.asarray(Ts, np.int64), out=fo[1:])
    assert (r.frame_ph[fo[5]:fo[6]] == Cc - 1).all() and (r.frame_idx[fo[5]:fo[6]] == -1).all()
    for b in (0, 1, 2, 3, 4, 6, 7):
        l, t = utts[b]
        o = orc.align_batch(orc.params(Cc - 1, 0), l.numpy().reshape(1, Ts[b], Cc), np.zeros(1, np.int64), np.asarray([Ts[b]], np.int32), Cc,
                            t.numpy().astype(np.int32), np.asarray([0, Ns[b]], np.int64), max_stamps=r.max_stamps, n_threads=1)
        np.testing.assert_array_equal(r.frame_ph[fo[b]:fo[b + 1]].cpu().numpy(), o["frame_ph"], err_msg=f"utterance {b} {shapes[b]}")
        np.testing.assert_array_equal(r.frame_idx[fo[b]:fo[b + 1]].cpu().numpy(), o["frame_idx"], err_msg=f"utterance {b} {shapes[b]}")
        n = int(o["n_stamps"][0])
        assert int(r.n_stamps[b]) == n
        np.testing.assert_array_equal(r.stamps[b, :n, 1].cpu().numpy(), o["stamps"]["start"][0][:n])
        np.testing.assert_allclose(float(r.dp_final[b]), float(o["dp_final"][0]), rtol=1e-4)
    # degenerate long problems (flat posteriors: ties, unreachable end states) follow the same exact semantics
    T, N = 1500, 300
    flat_lp = torch.log_softmax(torch.zeros(1, T, Cc), -1)
    tg1 = torch.randint(1, Cc - 1, (1, N), generator=torch.Generator().manual_seed(5))
    w1 = dict(lp=flat_lp.reshape(-1), row_off=torch.zeros(1, dtype=torch.int64), Ts=[T], tgt=tg1.to(torch.int32).reshape(-1), Ns=[N])
    r1 = _align(bfa, dev, w1, Cc)
    o1 = orc.align_batch(orc.params(Cc - 1, 0), flat_lp.numpy(), np.zeros(1, np.int64), np.asarray([T], np.int32), Cc, tg1.numpy().astype(np.int32).reshape(-1),
                         np.asarray([0, N], np.int64), max_stamps=r1.max_stamps, n_threads=1)
    np.testing.assert_array_equal(r1.frame_ph[:T].cpu().numpy(), o1["frame_ph"])
    np.testing.assert_array_equal(r1.frame_idx[:T].cpu().numpy(), o1["frame_idx"])
    # the reference-shaped entry warns and returns no stamps for the refused utterance
    au = bfa.AlignmentUtils(Cc - 1, 0)
    sel = [0, 5, 3]
    Tm, Nm = max(Ts[b] for b in sel), max(Ns[b] for b in sel)
    lp = torch.full((3, Tm, Cc), -20.0); tg2 = torch.zeros((3, Nm), dtype=torch.long)
    for k, b in enumerate(sel):
        l, t = utts[b]
        lp[k, :Ts[b]] = l; tg2[k, :Ns[b]] = t
    with pytest.warns(UserWarning, match="were not aligned"):
        got = au.decode_alignments(lp.to(dev), true_seqs=tg2, pred_lens=torch.tensor([Ts[b] for b in sel]), true_seqs_lens=torch.tensor([Ns[b] for b in sel]))
    assert got[1] == [] and len(got[0]) == 30 and len(got[2]) == 260


def test_two_host_threads_one_device(bfa, dev):
    """Two host threads aligning different ragged batches on the same device at the same time, each on its own stream (the
    side-stream pass of the exact kernel is shared per device and is serialised by a mutex): same results as one after the other."""
    import threading
    from bfa_b200 import synth
    Cc = 66
    batches = []
    for s in range(2):
        utts = synth.ragged_batch(96, C=Cc, t_range=(60, 500), n_range=(4, 120), seed=810 + s)
        batches.append(synth.pack_ragged(utts, Cc, device=dev))
    dec = bfa.AlignmentUtils(Cc - 1, 0).viterbi_decoder

    def run(k, out, stream=None):
        flat, row_off, Ts, tg, Ns = batches[k]
        p = dec._params(True, True, True)
        ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream())
        with ctx:
            for _ in range(6):
                r = dec.align_batch(flat, row_off, Ts, Cc, tg, Ns, params=p)
            (stream or torch.cuda.current_stream()).synchronize()
        out[k] = (r.frame_ph.clone(), r.frame_idx.clone(), r.status.clone(), r.n_stamps.clone())
    ref, got = {}, {}
    for k in range(2):
        run(k, ref)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    th = [threading.Thread(target=run, args=(k, got, streams[k])) for k in range(2)]
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    for k in range(2):
        for a_, b_ in zip(ref[k], got[k]):
            assert torch.equal(a_, b_), f"batch {k} differs when two threads share the device"


def test_confidence_rejects_out_of_range_stamps(bfa, dev):
    """utils.py:89 indexes probs[start_frame, phoneme_id]: a stamp outside the matrix raises (IndexError in the reference)."""
    lp = torch.log_softmax(torch.randn(20, 9), -1).to(dev)
    with pytest.raises(IndexError):
        bfa._calculate_confidences(lp, [(3, 25, 30, 0, False)])
    with pytest.raises(IndexError):
        bfa._calculate_confidences(lp, [(9, 2, 5, 0, False)])
    ok = bfa._calculate_confidences(lp, [(3, 2, 5, 0, False)])
    assert len(ok) == 1 and 0.0 < ok[0][5] <= 1.0
