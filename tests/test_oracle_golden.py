"""CPU: the C oracle (oracle/bfa_oracle.c) against vectors produced by the unmodified
reference module (tests/golden/make_golden.py).  Bit-exact for indices and for DP values fed
identical log-probs; 1e-5 relative where torch's log_softmax/exp numerics are involved."""
import numpy as np
import pytest

from oracle import oracle as orc


def _cases(npz):
    return [str(c) for c in npz["__cases__"]]


def test_min_log_prob_constant(golden):
    g = golden("batch_api")
    p = orc.params(66)
    assert np.float32(p.min_log_prob) == g["const/min_log_prob"]


def test_viterbi_core_bit_exact(golden):
    g = golden("viterbi_core")
    for name in _cases(g):
        band, blank, forced, fstate = (int(v) for v in g[f"{name}/meta"])
        keep = f"{name}/dp" in g.files
        r = orc.viterbi(g[f"{name}/lp"], g[f"{name}/path"], g[f"{name}/tidx"], band, blank, bool(forced), dump=keep)
        assert r["final_state"] == fstate, name
        np.testing.assert_array_equal(r["frame_ph"], g[f"{name}/frame_ph"], err_msg=name)
        np.testing.assert_array_equal(r["frame_idx"], g[f"{name}/frame_idx"], err_msg=name)
        assert r["dp_final"].tobytes() == g[f"{name}/dp_last"][fstate].tobytes(), name
        if keep:
            assert r["dp"].tobytes() == g[f"{name}/dp"].tobytes(), name
            np.testing.assert_array_equal(r["bp"][1:], g[f"{name}/bp"][1:], err_msg=name)


def test_degenerate_cases_really_wrap(golden):
    g = golden("viterbi_core")
    assert (g["degenerate_flat_T600/ps"] < 0).any()


def test_decode_forced(golden):
    g = golden("decode_forced")
    seen_seg = 0
    for name in _cases(g):
        blank, sil, anchors, forced, boost, floor, err, segmented = (int(v) for v in g[f"{name}/meta"])
        p = orc.params(blank, None if sil < 0 else sil, anchors, True, bool(forced), bool(boost), bool(floor))
        lp, seq = g[f"{name}/lp"], g[f"{name}/seq"]
        r = orc.decode_forced(lp, seq, p)
        st = r["status"] & 7
        if err:
            assert st == orc.ORC_TOO_SHORT, name
            continue
        assert st != orc.ORC_TOO_SHORT, name
        if segmented >= 0:
            assert (st == orc.ORC_SEGMENTED) == bool(segmented), name
            seen_seg += segmented
        np.testing.assert_array_equal(r["frame_ph"], g[f"{name}/frame_ph"], err_msg=name)
        np.testing.assert_array_equal(r["frame_idx"], g[f"{name}/frame_idx"], err_msg=name)
        if f"{name}/modified" in g.files:
            np.testing.assert_allclose(orc.prep(lp, seq, p), g[f"{name}/modified"], rtol=1e-5, atol=1e-5, err_msg=name)
        stamps = orc.assort(r["frame_ph"], r["frame_idx"], blank)
        np.testing.assert_array_equal(np.array(stamps, np.int32).reshape(-1, 4), g[f"{name}/stamps"], err_msg=name)
        conf = orc.confidence(lp, stamps)
        np.testing.assert_allclose(conf, g[f"{name}/conf"], rtol=1e-5, atol=1e-7, err_msg=name)
    assert seen_seg >= 5


def _batch(g, key, seq_lens, **kw):
    lp = g[f"{key}/lp"]; tgt = g[f"{key}/tgt"]
    B, Tm, C = lp.shape
    return lp, tgt, B, Tm, C


def test_batch_api(golden):
    g = golden("batch_api")
    lp = g["batch/lp"]; tgt = g["batch/tgt"]; T = g["batch/pred_lens"]; B, Tm, C = lp.shape
    row_off = np.arange(B, dtype=np.int64) * Tm * C
    for key, seq_lens, kw in (("batch", g["batch/seq_lens"], {}),
                              ("simple", g["simple/seq_lens"], dict(mode=1)),
                              ("noise", g["noise/seq_lens"], dict(silence_anchors=0, ignore_noise=False))):
        p = orc.params(66, 0, **kw)
        tflat = np.concatenate([tgt[i, :seq_lens[i]] for i in range(B)]).astype(np.int32)
        toff = np.zeros(B + 1, np.int64); np.cumsum(seq_lens, out=toff[1:])
        r = orc.align_batch(p, lp, row_off, T, C, tflat, toff, max_stamps=Tm, n_threads=3)
        for i in range(B):
            n = r["n_stamps"][i]
            got = np.stack([r["stamps"][i][f][:n] for f in ("phoneme", "start", "end", "target_idx")], 1)
            np.testing.assert_array_equal(got, g[f"{key}/stamps{i}"], err_msg=f"{key}{i}")
            if key == "batch":
                np.testing.assert_allclose(r["conf"][i][:n], g[f"batch/conf{i}"], rtol=1e-5, atol=1e-7)


def test_group_head(golden):
    g = golden("batch_api")
    lp = g["group/lp"]; tgt = g["group/tgt"]; B, Tm, C = lp.shape
    T = np.array([150, 140, 100], np.int32); sl = np.array([14, 14, 9])
    p = orc.params(16, 0)
    tflat = np.concatenate([tgt[i, :sl[i]] for i in range(B)]).astype(np.int32)
    toff = np.zeros(B + 1, np.int64); np.cumsum(sl, out=toff[1:])
    r = orc.align_batch(p, lp, np.arange(B, dtype=np.int64) * Tm * C, T, C, tflat, toff, max_stamps=Tm)
    for i in range(B):
        n = r["n_stamps"][i]
        got = np.stack([r["stamps"][i][f][:n] for f in ("phoneme", "start", "end", "target_idx")], 1)
        np.testing.assert_array_equal(got, g[f"group/stamps{i}"])


def test_silence_scan(golden):
    g = golden("batch_api")
    lp = g["silscan/lp"]
    nonempty = 0
    for j in range(6):
        thr, k = g[f"silscan/args{j}"]
        segs = orc.detect_silence(lp, 0, float(thr), int(k))
        np.testing.assert_array_equal(np.array(segs, np.int32).reshape(-1, 2), g[f"silscan/segs{j}"])
        nonempty += len(segs) > 0
    assert nonempty >= 3


def test_soft_boundaries_oracle_vs_reference():
    """extend_soft_boundaries_func (core.py:682-809): the C restatement against fixtures from the unmodified method."""
    import numpy as np
    from pathlib import Path
    from oracle import oracle as orc
    g = np.load(Path(__file__).parent / "golden" / "soft.npz")
    cases = sorted({k.split("/")[0] for k in g.files}, key=lambda c: int(c[1:]))
    assert len(cases) == 15
    changed = 0
    for c in cases:
        st_in, st_out = g[f"{c}/in"], g[f"{c}/out"]
        got = orc.soft_boundaries(g[f"{c}/lp"], [tuple(r) for r in st_in], int(g[f"{c}/soft"][0]))
        np.testing.assert_array_equal(np.array(got, np.int32).reshape(-1, 4), st_out[:, :4], err_msg=c)
        changed += int((st_in[:, 1:3] != st_out[:, 1:3]).any(1).sum())
    assert changed > 200
