"""Host-side mirror of the reference's alignment interface, backed by libbfa_b200.so.

Same class names, constructor arguments, method signatures, return types and error behaviour as
  bournemouth_aligner/forced_alignment.py   ViterbiDecoder (:11), AlignmentUtils (:836)
  bournemouth_aligner/utils.py              _calculate_confidences (:70)
so core.py:256-257 / :902-937 / :1028 can switch imports.  torch is used for device memory and
streams only; every computation on the path runs in the CUDA library.  There is no CPU fallback:
without a CUDA device or without the built library the calls raise.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from ._cabi import BfaError, BfaParams, BfaShape

Stamp = Tuple[int, int, int, int]


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise BfaError(f"{what} must be a CUDA tensor (got {t.device}); use align_host() for host buffers")


class _Workspace:
    """Grow-only device scratch, one per (aligner, device, CUDA stream): calls in flight on different streams never share
    counters, item lists or back-pointer slabs.  A buffer that is outgrown is handed back to the caching allocator, which
    only reuses it for work enqueued later on the same stream."""

    def __init__(self):
        self.buf = {}

    def get(self, nbytes: int, device) -> torch.Tensor:
        key = (str(device), int(torch.cuda.current_stream(device).cuda_stream))
        b = self.buf.get(key)
        if b is None or b.numel() < nbytes:
            b = torch.empty(int(nbytes * 1.1) + 256, dtype=torch.uint8, device=device)
            self.buf[key] = b
        return b


def result_arena_words(B: int, ms: int, want_stamps=True, want_conf=True) -> dict:
    """Word offsets of the per-utterance result arrays inside one int32 allocation (sections 16-byte aligned)."""
    off, o = {}, 0
    def sec(name, n):
        nonlocal o
        off[name] = o
        o += (n + 3) // 4 * 4
    sec("stamps", B * ms * 4 if want_stamps else 0)
    sec("conf", B * ms if (want_stamps and want_conf) else 0)
    sec("n_stamps", B if want_stamps else 0)
    sec("status", B)
    sec("dp_final", B)
    off["total"] = max(o, 4)
    return off


class BatchResult:
    """Device-resident result arrays of one batched alignment call (C-ABI layout)."""

    def __init__(self, frame_ph, frame_idx, frame_off, dp_final, status, stamps, conf, n_stamps, T, max_stamps):
        self.frame_ph, self.frame_idx, self.frame_off = frame_ph, frame_idx, frame_off
        self.dp_final, self.status = dp_final, status
        self.stamps, self.conf, self.n_stamps = stamps, conf, n_stamps
        self.T, self.max_stamps = T, max_stamps
        self.arena = None   # the single allocation behind stamps/conf/n_stamps/status/dp_final (align_batch), or None
        self.row_lse = None  # align_batch(logits=True): float32[total_frames] log-sum-exp of the rows (see bfa_align_batch_logits)

    def stamp_lists(self, with_conf: bool = False) -> List[List[tuple]]:
        """list[B] of list[(phoneme, start, end_exclusive, target_idx[, conf])] -- forced_alignment.py:871-872."""
        B = len(self.T) if self.T is not None else int(self.n_stamps.shape[0])
        if B == 0:
            return []
        # one device->host copy of the packed arena when there is one (stamps | conf | n_stamps live in a single allocation)
        n = self.n_stamps[:B].cpu().numpy()
        st = self.stamps[:B].cpu().numpy()
        cf = self.conf[:B].cpu().numpy() if (with_conf and self.conf is not None) else None
        # vectorised: all valid rows flattened once, converted to Python objects in one go, then cut per utterance
        valid = np.arange(st.shape[1])[None, :] < n[:, None]
        rows = st[valid]                                   # [sum n, 4]
        cols = [rows[:, i].tolist() for i in range(4)]
        if cf is None:
            flat = list(zip(*cols))
        else:
            flat = list(zip(*cols, cf[valid].tolist()))
        out, pos = [], 0
        for k in n.tolist():
            out.append(flat[pos:pos + k])
            pos += k
        return out


def exact_items_hint(T, N, params) -> int:
    """How many utterances will (probably) need the exact kernel instead of the banded stride-4 one: the stride rule picks
    3/2/1 (forced_alignment.py:153-157, :963-968), more than 128 phonemes, or a band wider than the 64-group window.
    Only a scheduling hint (BfaShape.reserved): with silence-anchored segmentation the planner may decide otherwise."""
    T = np.asarray(T, np.int64); N = np.asarray(N, np.int64)
    if T.size == 0:
        return 0
    L = 4 * N + 1
    dense = (L > 0.8 * T) if params.mode == _cabi.MODE_SIMPLE else (L > T)
    band = np.where(L > 60, np.maximum(L // 4, 20), 0)
    adv = (8 * (L - 1) + np.maximum(T - 2, 0)) // np.maximum(T - 1, 1)
    need = np.where(band > 0, np.minimum((2 * band + adv + 6) // 4 + 1, N + 1), N + 1)
    exact = (N > 0) & (T >= N) & (dense | (N > 128) | (need > 64))
    return int(exact.sum())


def log_softmax_rows(logits: torch.Tensor) -> torch.Tensor:
    """F.log_softmax(logits, dim=2) (core.py:898-899) for a CUDA [B, T, C] tensor (bfa_stitch_log_softmax without windows)."""
    if not logits.is_cuda:
        raise ValueError("logits must be a CUDA tensor (there is no CPU path)")
    x = logits if (logits.dtype == torch.float32 and logits.is_contiguous()) else logits.contiguous().float()
    B, T, C_ = x.shape
    out = torch.empty_like(x)
    rc = _cabi.lib().bfa_stitch_log_softmax(B, 0, 0, C_, T, _ptr(x), T * C_, None, _ptr(out), T * C_, _stream(x.device))
    _cabi.check(rc)
    return out


def direct_only_worthwhile(T, N, params) -> bool:
    """True when (nearly) every utterance of the batch is a plain stride-4 DP problem (forced_alignment.py:153: 4N+1 <= T), i.e.
    when launching only the direct kernel (BFA_FLAG_DIRECT_ONLY) will finish the batch.  The caller has already established that no
    target holds silence_id (or anchoring is off).  A wrong guess costs one extra launch of the full chain, never a wrong result."""
    T = np.asarray(T, np.int64); N = np.asarray(N, np.int64)
    if T.size == 0 or not params.ignore_noise:
        return False
    L = 4 * N + 1
    ok = (N >= 1) & (N <= 128) & (T >= 2) & ((L <= 0.8 * T) if params.mode == _cabi.MODE_SIMPLE else (L <= T))
    band = np.where(L > 60, np.maximum(L // 4, 20), 0)
    adv = (8 * (L - 1) + np.maximum(T - 2, 0)) // np.maximum(T - 1, 1)
    need = np.where(band > 0, np.minimum((2 * band + adv + 6) // 4 + 1, N + 1), N + 1)
    ok &= need <= 24
    return bool(ok.all())


class BatchPlan:
    """Device-resident shape metadata of one ragged batch (frame / target offsets, lengths, BfaShape)."""

    def __init__(self, T, N, C_, params, dev, max_stamps=None):
        B = len(T)
        self.T_np = np.asarray(T, np.int32)
        N_np = np.asarray(N, np.int64)
        frame_off_np = np.zeros(B + 1, np.int64); np.cumsum(self.T_np, out=frame_off_np[1:])
        tgt_off_np = np.zeros(B + 1, np.int64); np.cumsum(N_np, out=tgt_off_np[1:])
        self.B, self.total = B, int(frame_off_np[-1])
        max_T = int(self.T_np.max()) if B else 0
        max_N = int(N_np.max()) if B else 0
        if max_stamps is None:
            # with ignore_noise every target index owns one run of frames on a monotone path: N stamps.  Degenerate paths
            # (scores below the -1000 sentinel) can produce more; the library flags that (ST_STAMP_OVERFLOW) and the
            # reference-API entry points below then repeat the call with the pitch nothing can exceed (max_T).
            max_stamps = (max_N + 8) if params.ignore_noise else max(max_T, 1)
        self.max_stamps = int(max_stamps)
        self.shape = BfaShape(B, C_, max_T, max_N, self.total, self.max_stamps, exact_items_hint(self.T_np, N_np, params))
        meta = torch.from_numpy(np.concatenate([frame_off_np, tgt_off_np])).to(dev)
        self.frame_off, self.tgt_off = meta[: B + 1], meta[B + 1:]
        self.T_dev = torch.from_numpy(self.T_np).to(dev)
        self.ws_bytes = None


class ViterbiDecoder:
    """forced_alignment.py:11-834 (same constructor signature, :16)."""

    def __init__(self, blank_id, silence_id, silence_anchors=3, min_phoneme_prob=1e-8, ignore_noise=True, truly_forced=False):
        self.blank_id = blank_id
        self.silence_id = silence_id
        self.silence_anchors = silence_anchors
        self.min_phoneme_prob = min_phoneme_prob
        self.ignore_noise = ignore_noise
        self.truly_forced = truly_forced
        self._neg_inf = -1000.0
        self._ws = _Workspace()

    def set_blank_id(self, blank_id):
        self.blank_id = blank_id

    # ---- parameter block -------------------------------------------------------------------
    def _params(self, boost_targets=True, enforce_minimum=True, anchor_pauses=True, mode=_cabi.MODE_FULL) -> BfaParams:
        if self.blank_id is None:
            raise ValueError("Blank ID not set. Call set_blank_id first.")  # forced_alignment.py:104-105
        p = _cabi.default_params(self.blank_id, self.silence_id)
        p.silence_anchors = int(self.silence_anchors) if anchor_pauses else 0
        p.ignore_noise = int(bool(self.ignore_noise))
        p.truly_forced = int(bool(self.truly_forced))
        p.boost_targets = int(bool(boost_targets))
        p.enforce_minimum = int(bool(enforce_minimum))
        p.min_log_prob = float(np.log(np.float32(self.min_phoneme_prob)))
        p.neg_inf = float(self._neg_inf)
        p.mode = mode
        return p

    # ---- batched core ----------------------------------------------------------------------
    def plan_batch(self, T: Sequence[int], N: Sequence[int], C_: int, *, params: BfaParams, device, max_stamps=None) -> "BatchPlan":
        """Shape metadata of a ragged batch, uploaded once; reusable for every call with the same (T, N)."""
        return BatchPlan(T, N, C_, params, torch.device(device), max_stamps)

    def align_batch(self, log_probs: torch.Tensor, row_off: torch.Tensor, T: Sequence[int], C_: int, tgt: torch.Tensor,
                    N: Sequence[int], *, params: BfaParams, want_stamps=True, want_conf=True, max_stamps=None,
                    plan: Optional["BatchPlan"] = None, out: Optional[BatchResult] = None,
                    arena: Optional[torch.Tensor] = None, logits: bool = False) -> BatchResult:
        """Ragged batch through bfa_align_batch.  log_probs: flat/any-shape fp32 CUDA tensor holding the rows,
        row_off int64[B] element offsets (CUDA), T/N python sequences, tgt flat int32 CUDA targets.
        `plan` (from plan_batch) skips the per-call metadata upload, `out` reuses a previous result's buffers.
        `arena`: caller-owned int32 storage for the packed per-utterance results (>= result_arena_words(..)["total"] words,
        16-byte aligned) instead of a fresh allocation -- e.g. this rank's slice of ANOTHER GPU's symmetric-memory buffer
        (sharding.PeerArena): the kernels then write the timestamp arrays straight into the gathering rank's memory.
        `logits=True`: the rows hold the acoustic model's UN-NORMALISED logits (core.py:898-899 skipped); one kernel aligns them
        and leaves `result.row_lse` (float32[total_frames], log-sum-exp of every row of the utterances it finished) behind;
        utterances it cannot finish come back with status ST_DEFERRED (bfa_align_batch_logits; full mode with boosting only)."""
        _require_cuda(log_probs, "log_probs")
        dev = log_probs.device
        if log_probs.dtype != torch.float32 or not log_probs.is_contiguous():
            raise BfaError("log_probs must be contiguous float32")
        if plan is None:
            plan = BatchPlan(T, N, C_, params, dev, max_stamps)
        B, total, ms = plan.B, plan.total, plan.max_stamps
        if out is None:
            frame_ph = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
            frame_idx = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
            # the per-utterance results live in ONE allocation (stamps | conf | n_stamps | status | dp_final, 4-byte words),
            # so that the multi-GPU gather of the timestamp arrays is a single collective (sharding.gather_packed)
            Bp = max(B, 1)
            words = result_arena_words(Bp, ms, want_stamps, want_conf)
            if arena is None:
                arena = torch.empty(words["total"], dtype=torch.int32, device=dev)
            else:
                if arena.dtype != torch.int32 or arena.dim() != 1 or arena.numel() < words["total"] or not arena.is_cuda or arena.data_ptr() % 16:
                    raise BfaError(f"arena must be a 16-byte aligned 1-D int32 CUDA tensor of >= {words['total']} words")
                arena = arena[:words["total"]]
            dp_final = arena[words["dp_final"]:words["dp_final"] + Bp].view(torch.float32)
            status = arena[words["status"]:words["status"] + Bp]
            stamps = conf = n_stamps = None
            if want_stamps:
                stamps = arena[words["stamps"]:words["stamps"] + Bp * ms * 4].view(Bp, ms, 4)
                n_stamps = arena[words["n_stamps"]:words["n_stamps"] + Bp]
                if want_conf:
                    conf = arena[words["conf"]:words["conf"] + Bp * ms].view(torch.float32).view(Bp, ms)
            out = BatchResult(frame_ph, frame_idx, plan.frame_off, dp_final, status, stamps, conf, n_stamps, plan.T_np, ms)
            out.arena = arena
        if logits and getattr(out, "row_lse", None) is None:
            out.row_lse = torch.zeros(max(total, 1), dtype=torch.float32, device=dev)
        l = _cabi.lib()
        with torch.cuda.device(dev):
            if plan.ws_bytes is None:
                plan.ws_bytes = l.bfa_workspace_bytes(C.byref(params), C.byref(plan.shape))
                if plan.ws_bytes == 0 and B > 0:
                    _cabi.check(_cabi.BFA_E_UNSUPPORTED)
            ws = self._ws.get(max(plan.ws_bytes, 256), dev)
            if logits:
                rc = l.bfa_align_batch_logits(C.byref(params), C.byref(plan.shape), _ptr(log_probs), _ptr(row_off), _ptr(plan.T_dev), _ptr(tgt),
                                              _ptr(plan.tgt_off), _ptr(out.frame_ph), _ptr(out.frame_idx), _ptr(plan.frame_off),
                                              _ptr(out.dp_final), _ptr(out.status), _ptr(out.stamps), _ptr(out.conf), _ptr(out.n_stamps),
                                              _ptr(out.row_lse), _ptr(ws), ws.numel(), _stream(dev))
            else:
                rc = l.bfa_align_batch(C.byref(params), C.byref(plan.shape), _ptr(log_probs), _ptr(row_off), _ptr(plan.T_dev), _ptr(tgt),
                                       _ptr(plan.tgt_off), _ptr(out.frame_ph), _ptr(out.frame_idx), _ptr(plan.frame_off), _ptr(out.dp_final),
                                       _ptr(out.status), _ptr(out.stamps), _ptr(out.conf), _ptr(out.n_stamps), _ptr(ws), ws.numel(),
                                       _stream(dev))
        _cabi.check(rc)
        return out

    @staticmethod
    def _raise_if_too_short(status_np, T, N):
        bad = np.nonzero((status_np & 7) == _cabi.ST_TOO_SHORT)[0]
        if bad.size:
            i = int(bad[0])
            raise ValueError(  # same text as forced_alignment.py:162-165
                f"Audio too short to align: {int(N[i])} phonemes cannot be fit into "
                f"{int(T[i])} frames (need at least 1 frame per phoneme)."
            )

    # ---- reference methods -----------------------------------------------------------------
    def decode_with_forced_alignment(self, log_probs, true_sequence, return_scores=False, boost_targets=True,
                                     enforce_minimum=True, anchor_pauses=True, debug=False):
        """forced_alignment.py:87-199.  Returns (frame_phonemes i64[T], frame_phonemes_idx i64[T], score|None)."""
        _require_cuda(log_probs, "log_probs")
        lp = log_probs.contiguous().float()
        T, C_ = lp.shape
        seq = true_sequence.to(lp.device)
        N = int(seq.shape[0])
        p = self._params(boost_targets, enforce_minimum, anchor_pauses)
        r = self.align_batch(lp, torch.zeros(1, dtype=torch.int64, device=lp.device), [T], C_, seq.to(torch.int32).contiguous(),
                             [N], params=p, want_stamps=False)
        self._raise_if_too_short(r.status[:1].cpu().numpy(), [T], [N])
        fp, fi = r.frame_ph[:T].long(), r.frame_idx[:T].long()
        score = self._calculate_alignment_score(lp, fp) if return_scores else None
        return fp, fi, score

    def _viterbi_decode(self, log_probs, ctc_path, ctc_len, ctc_path_true_idx=None, band_width=0, debug=False):
        """forced_alignment.py:563-703 on an explicit CTC path."""
        r = self.viterbi_paths(log_probs.contiguous().float(), [log_probs.shape[0]], [0], ctc_path[:ctc_len], ctc_path_true_idx,
                               [int(ctc_len)], [int(band_width)])
        fp = r["frame_ph"].long()
        return fp, (r["frame_idx"].long() if ctc_path_true_idx is not None else None)

    def viterbi_paths(self, log_probs, T: Sequence[int], row_off: Sequence[int], path, true_idx, L: Sequence[int],
                      band: Sequence[int]):
        """Batched bfa_viterbi_paths: item i uses rows at row_off[i] (elements), path[path_off[i]: +L[i]]."""
        _require_cuda(log_probs, "log_probs")
        dev = log_probs.device
        C_ = int(log_probs.shape[-1])
        n = len(T)
        T_np = np.asarray(T, np.int32); L_np = np.asarray(L, np.int32)
        f_off = np.zeros(n + 1, np.int64); np.cumsum(T_np, out=f_off[1:])
        p_off = np.zeros(n + 1, np.int64); np.cumsum(L_np, out=p_off[1:])
        i64 = lambda a: torch.from_numpy(np.asarray(a, np.int64)).to(dev)
        i32 = lambda a: torch.from_numpy(np.asarray(a, np.int32)).to(dev)
        path_d = path.to(dev).to(torch.int32).contiguous()
        tidx_d = None if true_idx is None else true_idx.to(dev).to(torch.int32).contiguous()
        total = int(f_off[-1])
        frame_ph = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        frame_idx = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        dp_final = torch.empty(n, dtype=torch.float32, device=dev)
        fstate = torch.empty(n, dtype=torch.int32, device=dev)
        row_off_d, T_d, p_off_d, L_d, band_d, f_off_d = i64(row_off), i32(T_np), i64(p_off), i32(L_np), i32(band), i64(f_off)
        p = self._params(False, False, False)
        l = _cabi.lib()
        with torch.cuda.device(dev):
            max_T, max_L = int(T_np.max()), int(L_np.max())
            if max_L > _cabi.MAX_L:
                _cabi.check(_cabi.BFA_E_UNSUPPORTED)
            ws_bytes = l.bfa_viterbi_paths_workspace_bytes(n, max_T, max_L)
            ws = self._ws.get(max(ws_bytes, 256), dev)
            rc = l.bfa_viterbi_paths(C.byref(p), n, C_, max_T, max_L, _ptr(log_probs), _ptr(row_off_d), _ptr(T_d), _ptr(path_d),
                                     _ptr(tidx_d), _ptr(p_off_d), _ptr(L_d), _ptr(band_d), _ptr(frame_ph), _ptr(frame_idx),
                                     _ptr(f_off_d), _ptr(dp_final), _ptr(fstate), _ptr(ws), ws.numel(), _stream(dev))
        _cabi.check(rc)
        return dict(frame_ph=frame_ph[:total], frame_idx=frame_idx[:total], frame_off=f_off, dp_final=dp_final, final_state=fstate)

    def _calculate_alignment_score(self, log_probs, frame_phonemes):
        """forced_alignment.py:767-773: sum of log_probs[t, frame_phonemes[t]] (bfa_alignment_score_batch, one warp per utterance,
        accumulated in double like the reference's Python float)."""
        _require_cuda(log_probs, "log_probs")
        lp = log_probs if (log_probs.dtype == torch.float32 and log_probs.is_contiguous()) else log_probs.contiguous().float()
        dev = lp.device
        T_, C_ = lp.shape
        ph = torch.as_tensor(frame_phonemes).to(device=dev, dtype=torch.int32).contiguous()
        n = min(T_, int(ph.numel()))
        score = torch.empty(1, dtype=torch.float64, device=dev)
        z = torch.zeros(1, dtype=torch.int64, device=dev)
        rc = _cabi.lib().bfa_alignment_score_batch(1, C_, _ptr(lp), _ptr(z), _ptr(torch.tensor([n], dtype=torch.int32, device=dev)), _ptr(ph),
                                                   _ptr(z), _ptr(score), _stream(dev))
        _cabi.check(rc)
        return float(score.item())

    def assort_frames(self, frame_phonemes, frame_phonemes_idx, max_blanks=10) -> List[Stamp]:
        """forced_alignment.py:777-834."""
        if len(frame_phonemes) == 0:
            return []
        fp = torch.as_tensor(frame_phonemes)
        fi = torch.as_tensor(frame_phonemes_idx)
        dev = fp.device if fp.is_cuda else torch.device("cuda", torch.cuda.current_device())
        if not torch.cuda.is_available():
            raise BfaError("assort_frames needs a CUDA device")
        fp = fp.to(dev).to(torch.int32).contiguous(); fi = fi.to(dev).to(torch.int32).contiguous()
        T = int(fp.shape[0])
        p = self._params(False, False, False)
        p.max_blanks = int(max_blanks)
        ms = T
        stamps = torch.empty((1, ms, 4), dtype=torch.int32, device=dev)
        n_st = torch.empty(1, dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        T_d = torch.tensor([T], dtype=torch.int32, device=dev)
        f_off = torch.tensor([0, T], dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            rc = _cabi.lib().bfa_assort_batch(C.byref(p), 1, _ptr(T_d), _ptr(f_off), _ptr(fp), _ptr(fi), _ptr(status), _ptr(stamps),
                                              _ptr(n_st), ms, _stream(dev))
        _cabi.check(rc)
        n = int(n_st.item())
        return [tuple(r) for r in stamps[0, :n].cpu().tolist()]


class AlignmentUtils:
    """forced_alignment.py:836-986 (same constructor signature, :841)."""

    def __init__(self, blank_id, silence_id, silence_anchors=10, ignore_noise=True, truly_forced=True):
        self.blank_id = blank_id
        self.silence_id = silence_id
        self.silence_anchors = silence_anchors
        self.truly_forced = truly_forced
        self.viterbi_decoder = ViterbiDecoder(blank_id, silence_id, silence_anchors=self.silence_anchors,
                                              ignore_noise=ignore_noise, truly_forced=self.truly_forced)
        self.last_result: Optional[BatchResult] = None
        self.last_params_reserved = 0

    @staticmethod
    def _lens(x, B, default):
        if x is None:
            return [default] * B
        if isinstance(x, torch.Tensor):
            return [int(v) for v in x.tolist()]
        return [int(v) for v in x]

    def _dense_prepare(self, log_probs, true_seqs, pred_lens, true_seqs_lens, params, want_conf, input_is_logits=False):
        """Everything of a padded batch's alignment call that is not the call itself: lengths, flat targets, offsets and the batch
        plan on the device (small torch ops on the current stream), and the choice of launch sequence."""
        _require_cuda(log_probs, "log_probs")
        lp = log_probs if (log_probs.dtype == torch.float32 and log_probs.is_contiguous()) else log_probs.contiguous().float()
        B, T_max, C_ = lp.shape
        dev = lp.device
        T = self._lens(pred_lens, B, T_max)
        S = int(true_seqs.shape[1]) if true_seqs.dim() == 2 else 0
        N = self._lens(true_seqs_lens, B, S)
        # can any utterance take the silence-anchored segmentation attempt (forced_alignment.py:133)?
        may_segment = params.mode == _cabi.MODE_FULL and params.silence_anchors > 0 and params.silence_id >= 0
        if may_segment and not true_seqs.is_cuda and S > 0:
            # targets still live on the host (core.py builds them there): a free check lets the library skip the
            # row-statistics pass that only the silence scan needs (BFA_HINT_NO_SIL; purely a performance hint)
            lens = torch.as_tensor(N)[:, None]
            sil_utts = ((true_seqs == params.silence_id) & (torch.arange(S)[None, :] < lens)).any(dim=1)
            if not bool(sil_utts.any()):
                params.reserved |= _cabi.HINT_NO_SIL
                may_segment = False
            elif float(sil_utts.float().mean()) > 0.75:
                # (nearly) every target holds silence_id: the one-kernel pass would hand (nearly) everything back to the
                # planner chain; skip it (BFA_FLAG_NO_DIRECT; a performance switch, results do not depend on it)
                params.reserved |= _cabi.FLAG_NO_DIRECT
        # Without segmentation every utterance that is a plain stride-4 problem is finished by ONE kernel (in-kernel planning,
        # Viterbi, stamps, confidences): launch only that kernel; whatever it flags as deferred is run again in _dense_finish
        if not may_segment and S > 0 and direct_only_worthwhile(T, N, params):
            params.reserved |= _cabi.FLAG_DIRECT_ONLY
        seqs = true_seqs.to(dev)
        N_dev = torch.tensor(N, dtype=torch.int64, device=dev)
        mask = torch.arange(S, device=dev)[None, :] < N_dev[:, None]
        tgt = seqs[mask].to(torch.int32).contiguous()
        row_off = torch.arange(B, dtype=torch.int64, device=dev) * (T_max * C_)
        plan = self.viterbi_decoder.plan_batch(T, N, C_, params=params, device=dev)
        # Un-normalised logits (core.py:898-899 not run by the caller): when the batch goes to the one-kernel pass and boosting is
        # on, that kernel takes them as they are (bfa_align_batch_logits); otherwise they are normalised first, like the reference.
        #   (... or to the planner chain with its silence pass, which then also leaves the rows' log-sum-exp behind)
        chain_ok = may_segment and not (params.reserved & _cabi.HINT_NO_SIL) and 0 <= params.silence_id < C_
        use_logits = bool(input_is_logits and params.mode == _cabi.MODE_FULL and params.boost_targets
                          and ((params.reserved & _cabi.FLAG_DIRECT_ONLY) or chain_ok))
        if input_is_logits and not use_logits:
            lp = log_softmax_rows(lp)
        return dict(r=None, lp=lp, row_off=row_off, T=T, N=N, C=C_, tgt=tgt, params=params, want_conf=want_conf, B=B, T_max=T_max, plan=plan,
                    logits=use_logits)

    def _dense_enqueue(self, h, after_sibling=False):
        """Enqueue the prepared call (kernels only, no host synchronisation).  after_sibling: the operation before this one in the
        stream is the alignment kernel of ANOTHER prepared call whose preparation was enqueued after this call's preparation (the
        other head of the same batch, core.py:900-920): every input of this call was complete before that kernel was enqueued, so
        a one-kernel call may start while it drains (BFA_FLAG_PIPELINED)."""
        params = h["params"]
        if after_sibling and (params.reserved & _cabi.FLAG_DIRECT_ONLY):
            params.reserved |= _cabi.FLAG_PIPELINED
        run = lambda: self.viterbi_decoder.align_batch(h["lp"], h["row_off"], h["T"], h["C"], h["tgt"], h["N"], params=params, want_stamps=True,
                                                       want_conf=h["want_conf"], plan=h["plan"], logits=h.get("logits", False))
        try:
            h["r"] = run()
        except BfaError as e:
            # logits in, but neither the one-kernel pass nor a chain with a silence pass is available for this shape (more than 72
            # classes without silence_id in the batch, ...): normalise first, like the reference
            if not (h.get("logits") and e.code == _cabi.BFA_E_UNSUPPORTED):
                raise
            h["lp"] = log_softmax_rows(h["lp"])
            h["logits"] = False
            h["r"] = run()
        return h

    def _dense_launch(self, log_probs, true_seqs, pred_lens, true_seqs_lens, params, want_conf, input_is_logits=False):
        """Enqueue the alignment of a padded batch on the current stream and return without waiting for it: the first half of
        decode_alignments.  Two heads (core.py:900-920) can be launched back to back before either result is looked at."""
        return self._dense_enqueue(self._dense_prepare(log_probs, true_seqs, pred_lens, true_seqs_lens, params, want_conf, input_is_logits))

    def _dense_finish(self, h):
        """Second half: wait for the statuses, repeat the call where the library asked for it, raise what the reference raises."""
        r, params, B, T, N = h["r"], h["params"], h["B"], h["T"], h["N"]
        again = lambda **kw: self.viterbi_decoder.align_batch(h["lp"], h["row_off"], T, h["C"], h["tgt"], N, params=params, want_stamps=True,
                                                              want_conf=h["want_conf"], logits=h.get("logits", False), **kw)
        st = r.status[:B].cpu().numpy()
        if ((st & 7) == _cabi.ST_DEFERRED).any():  # the one-kernel path handed utterances back: the full chain takes the batch
            params.reserved &= ~(_cabi.FLAG_DIRECT_ONLY | _cabi.FLAG_PIPELINED)
            if h.get("logits"):                    # ... on normalised rows (no silence pass runs in this configuration to supply row_lse)
                h["lp"] = log_softmax_rows(h["lp"])
                h["logits"] = False
            r = again()
            st = r.status[:B].cpu().numpy()
        if (st & _cabi.ST_STAMP_OVERFLOW).any():   # more runs than the default stamp pitch (degenerate paths): use the safe pitch
            r = again(max_stamps=max(h["T_max"], 1))
            st = r.status[:B].cpu().numpy()
        if ((st & 7) == _cabi.ST_UNSUPPORTED).any():
            import warnings
            bad = np.nonzero((st & 7) == _cabi.ST_UNSUPPORTED)[0]
            warnings.warn(f"{bad.size} utterance(s) (first: #{int(bad[0])}, {int(N[int(bad[0])])} phonemes) need a CTC path of more than "
                          f"{_cabi.MAX_L} states and were not aligned (no timestamps returned for them); split such segments")
        self.last_result = r
        self.last_params_reserved = int(params.reserved)     # which path the batch finally took (FLAG_DIRECT_ONLY: one kernel)
        # input_is_logits: what downstream steps (soft boundaries, confidences of moved stamps) need to form log-probabilities:
        # either the rows' log-sum-exp (aligned straight from the logits; r.row_lse, packed like the frame labels) or the
        # normalised copy the fallback made
        self.last_row_lse = r.row_lse if h.get("logits") else None
        self.last_log_probs = None if h.get("logits") else h["lp"]
        self.viterbi_decoder._raise_if_too_short(st, T, N)
        return r

    def _dense_batch(self, log_probs, true_seqs, pred_lens, true_seqs_lens, params, want_conf, input_is_logits=False):
        return self._dense_finish(self._dense_launch(log_probs, true_seqs, pred_lens, true_seqs_lens, params, want_conf, input_is_logits))

    def decode_alignments_launch(self, log_probs, true_seqs, pred_lens=None, true_seqs_lens=None, boost_targets=True, enforce_minimum=True,
                                 with_confidence=False, input_is_logits=False):
        """decode_alignments split in two: this half only enqueues the kernels (no host synchronisation) and returns a handle for
        decode_alignments_finish.  core.py:900-920 aligns the phoneme head and the group head one after the other; with the two
        halves both heads are in flight before the host looks at either."""
        if (true_seqs is None) or (true_seqs_lens is None):
            raise ValueError("Phoneme sequences and lengths required for forced alignment")  # :878-879
        p = self.viterbi_decoder._params(boost_targets, enforce_minimum, self.silence_anchors > 0)
        h = self._dense_launch(log_probs, true_seqs, pred_lens, true_seqs_lens, p, with_confidence, input_is_logits)
        h["with_confidence"] = with_confidence
        return h

    def decode_alignments_prepare(self, log_probs, true_seqs, pred_lens=None, true_seqs_lens=None, boost_targets=True, enforce_minimum=True,
                                  with_confidence=False, input_is_logits=False):
        """decode_alignments_launch without the launch: a handle for decode_alignments_enqueue.  Preparing both heads of a batch
        first and enqueueing them afterwards puts the two alignment kernels next to each other in the stream."""
        if (true_seqs is None) or (true_seqs_lens is None):
            raise ValueError("Phoneme sequences and lengths required for forced alignment")  # :878-879
        p = self.viterbi_decoder._params(boost_targets, enforce_minimum, self.silence_anchors > 0)
        h = self._dense_prepare(log_probs, true_seqs, pred_lens, true_seqs_lens, p, with_confidence, input_is_logits)
        h["with_confidence"] = with_confidence
        return h

    def decode_alignments_enqueue(self, handle, after_sibling=False):
        return self._dense_enqueue(handle, after_sibling)

    def decode_alignments_finish(self, handle):
        return self._dense_finish(handle).stamp_lists(with_conf=handle["with_confidence"])

    def decode_alignments(self, log_probs, true_seqs=None, pred_lens=None, true_seqs_lens=None, forced_alignment=True,
                          boost_targets=True, enforce_minimum=True, debug=False, with_confidence=False, input_is_logits=False):
        """forced_alignment.py:856-928.  Returns list[B] of list[(phoneme, start, end, target_idx)].
        with_confidence=True (extension) appends utils._calculate_confidences' score to each tuple.
        input_is_logits=True (extension): `log_probs` holds the acoustic model's un-normalised logits, i.e. the caller skipped
        F.log_softmax (core.py:898-899); same results as on the normalised tensor (see bfa_align_batch_logits)."""
        if forced_alignment:
            if (true_seqs is None) or (true_seqs_lens is None):
                raise ValueError("Phoneme sequences and lengths required for forced alignment")  # :878-879
            p = self.viterbi_decoder._params(boost_targets, enforce_minimum, self.silence_anchors > 0)
            r = self._dense_batch(log_probs, true_seqs, pred_lens, true_seqs_lens, p, with_confidence, input_is_logits)
            return r.stamp_lists(with_conf=with_confidence)
        # free decoding (:912-928): frame-wise argmax then assort; returns ONE flat list like the reference
        _require_cuda(log_probs, "log_probs")
        B = log_probs.shape[0]
        T = self._lens(pred_lens, B, log_probs.shape[1])
        pred = torch.argmax(log_probs, dim=2)
        out = []
        for i in range(B):
            fp = pred[i, : T[i]]
            out.extend(self.viterbi_decoder.assort_frames(fp, torch.full_like(fp, -1)))
        return out

    def decode_alignments_simple(self, log_probs, true_seqs, pred_lens=None, true_seqs_lens=None):
        """forced_alignment.py:932-986: no boost, no floor, no anchoring."""
        p = self.viterbi_decoder._params(False, False, False, mode=_cabi.MODE_SIMPLE)
        r = self._dense_batch(log_probs, true_seqs, pred_lens, true_seqs_lens, p, False)
        return r.stamp_lists()


def _calculate_confidences(log_probs: torch.Tensor, framestamps):
    """utils.py:70-113.  framestamps: 5-tuples (phoneme, start, end, target_idx, is_estimated);
    returns 6-tuples with avg_confidence appended."""
    _require_cuda(log_probs, "log_probs")
    lp = log_probs.contiguous().float()
    T, C_ = lp.shape
    n = len(framestamps)
    if n == 0:
        return []
    dev = lp.device
    for (ph, s, e, _i, est) in framestamps:
        s2, e2 = max(0, int(s)), min(T, int(e))
        if est and not (s2 < T and e2 <= T):  # utils.py:88-91
            raise ValueError(f"Invalid frame range for estimated timestamp: start_frame={s2}, end_frame={e2}, "
                             f"log_probs shape={tuple(lp.shape)}, is_estimated={est}, phoneme_id={ph}")
    for (ph, s, e, _i, est) in framestamps:
        if not (-C_ <= int(ph) < C_) or max(0, int(s)) >= T:   # utils.py:89 indexes probs[start_frame, phoneme_id]
            raise IndexError(f"stamp (phoneme_id={ph}, start_frame={s}) is out of bounds for log_probs of shape {(T, C_)}")
    st = torch.tensor([[int(f[0]) % C_, int(f[1]), int(f[2]), int(f[3])] for f in framestamps], dtype=torch.int32, device=dev)
    conf = torch.empty(n, dtype=torch.float32, device=dev)
    i64 = lambda v: torch.tensor(v, dtype=torch.int64, device=dev)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
    row_off, T_d, n_d = i64([0]), i32([T]), i32([n])
    with torch.cuda.device(dev):
        rc = _cabi.lib().bfa_confidence_batch(1, C_, _ptr(lp), _ptr(row_off), _ptr(T_d), _ptr(st), _ptr(n_d), n, _ptr(conf),
                                              _stream(dev))
    _cabi.check(rc)
    c = conf.cpu().tolist()
    return [(f[0], max(0, int(f[1])), min(T, int(f[2])), f[3], f[4], c[i]) for i, f in enumerate(framestamps)]


def _lse_offsets(row_lse, T, dev):
    """Offsets of the utterances inside a packed row_lse (AlignmentUtils.last_row_lse: utterance u at sum(T[:u]))."""
    off = np.zeros(len(T), np.int64)
    np.cumsum(np.asarray(T[:-1], np.int64), out=off[1:])
    if int(off[-1]) + int(T[-1]) > row_lse.numel():
        raise ValueError("row_lse is shorter than the frames of the batch")
    return torch.from_numpy(off).to(dev)


def _calculate_confidences_batch(log_probs: torch.Tensor, framestamps, pred_lens=None, row_lse=None):
    """utils._calculate_confidences (utils.py:70-113) for a whole batch in ONE launch: log_probs [B, T, C] (CUDA), framestamps
    list[B] of lists of 5-tuples -> list[B] of lists of 6-tuples.  The reference loops over the batch (core.py:936-937);
    `pred_lens[b]` plays the role of `log_probs[b].shape[0]` when the batch is padded.
    row_lse (AlignmentUtils.last_row_lse after decode_alignments(input_is_logits=True)): `log_probs` holds un-normalised logits."""
    _require_cuda(log_probs, "log_probs")
    lp = log_probs if (log_probs.dtype == torch.float32 and log_probs.is_contiguous()) else log_probs.contiguous().float()
    B, T_max, C_ = lp.shape
    T = [T_max] * B if pred_lens is None else [int(v) for v in (pred_lens.tolist() if hasattr(pred_lens, "tolist") else pred_lens)]
    dev = lp.device
    for b, fs in enumerate(framestamps):
        for (ph, s, e, _i, est) in fs:
            s2, e2 = max(0, int(s)), min(T[b], int(e))
            if est and not (s2 < T[b] and e2 <= T[b]):  # utils.py:88-91
                raise ValueError(f"Invalid frame range for estimated timestamp: start_frame={s2}, end_frame={e2}, "
                                 f"log_probs shape={(T[b], C_)}, is_estimated={est}, phoneme_id={ph}")
    ms = max(1, max((len(f) for f in framestamps), default=1))
    st = np.zeros((B, ms, 4), np.int32)
    for b, fs in enumerate(framestamps):
        for i, f in enumerate(fs):
            if not (-C_ <= int(f[0]) < C_) or max(0, int(f[1])) >= T[b]:   # utils.py:89 indexes probs[start_frame, phoneme_id]
                raise IndexError(f"stamp (phoneme_id={f[0]}, start_frame={f[1]}) of item {b} is out of bounds for log_probs of shape {(T[b], C_)}")
            st[b, i] = (int(f[0]) % C_, int(f[1]), int(f[2]), int(f[3]))
    st_d = torch.from_numpy(st).to(dev)
    n_d = torch.tensor([len(f) for f in framestamps], dtype=torch.int32, device=dev)
    T_d = torch.tensor(T, dtype=torch.int32, device=dev)
    row_off = torch.arange(B, dtype=torch.int64, device=dev) * (T_max * C_)
    conf = torch.zeros((B, ms), dtype=torch.float32, device=dev)
    lse_off = _lse_offsets(row_lse, T, dev) if row_lse is not None else None
    with torch.cuda.device(dev):
        rc = _cabi.lib().bfa_confidence_batch_lse(B, C_, _ptr(lp), _ptr(row_off), _ptr(T_d), _ptr(st_d), _ptr(n_d), ms, _ptr(conf),
                                                  _ptr(row_lse), _ptr(lse_off), _stream(dev))
    _cabi.check(rc)
    c = conf.cpu().numpy()
    return [[(f[0], max(0, int(f[1])), min(T[b], int(f[2])), f[3], f[4], float(c[b, i])) for i, f in enumerate(fs)]
            for b, fs in enumerate(framestamps)]


def extend_soft_boundaries_func(log_probs: torch.Tensor, framestamps, boundary_softness=3, debug=False, row_lse=None, pred_lens=None):
    """PhonemeTimestampAligner.extend_soft_boundaries_func (core.py:682-809) as a free function: log_probs [B, T, C] (CUDA),
    framestamps list[B] of lists of 5-tuples (phoneme, start, end, target_idx, is_estimated) -> the same structure with the
    stretched boundaries.  One kernel over the whole batch (bfa_soft_boundaries_batch).
    row_lse (+ the pred_lens it was packed with): `log_probs` holds un-normalised logits, see _calculate_confidences_batch."""
    _require_cuda(log_probs, "log_probs")
    lp = log_probs if (log_probs.dtype == torch.float32 and log_probs.is_contiguous()) else log_probs.contiguous().float()
    B, T, C_ = lp.shape
    if len(framestamps) != B:
        raise ValueError("framestamps must hold one list per batch item")
    dev = lp.device
    ms = max(1, max((len(f) for f in framestamps), default=1))
    st = np.zeros((B, ms, 4), np.int32)
    for b, fs in enumerate(framestamps):
        for i, f in enumerate(fs):
            st[b, i] = (int(f[0]), int(f[1]), int(f[2]), int(f[3]))
    st_d = torch.from_numpy(st).to(dev)
    n_d = torch.tensor([len(f) for f in framestamps], dtype=torch.int32, device=dev)
    T_d = torch.full((B,), T, dtype=torch.int32, device=dev)
    row_off = torch.arange(B, dtype=torch.int64, device=dev) * (T * C_)
    lse_off = None
    if row_lse is not None:      # row_lse only covers the frames that were aligned: the stretch stops at each utterance's own length
        Tl = [T] * B if pred_lens is None else [int(v) for v in (pred_lens.tolist() if hasattr(pred_lens, "tolist") else pred_lens)]
        lse_off = _lse_offsets(row_lse, Tl, dev)
        T_d = torch.tensor(Tl, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _cabi.lib().bfa_soft_boundaries_batch_lse(B, C_, _ptr(lp), _ptr(row_off), _ptr(T_d), _ptr(st_d), _ptr(n_d), ms,
                                                       int(boundary_softness), _ptr(row_lse), _ptr(lse_off), _stream(dev))
    _cabi.check(rc)
    out = st_d.cpu().numpy()
    return [[(f[0], int(out[b, i, 1]), int(out[b, i, 2]), f[3], f[4]) for i, f in enumerate(fs)] for b, fs in enumerate(framestamps)]


def align_host(params: BfaParams, log_probs: np.ndarray, row_off: np.ndarray, T: np.ndarray, C_: int, tgt: np.ndarray,
               tgt_off: np.ndarray, *, max_stamps: Optional[int] = None, want_conf=True, device=0, chunk_utts=0, out=None):
    """bfa_align_batch_host: every buffer is a HOST numpy array (pin them for full PCIe speed).
    Returns dict of numpy arrays.  `out` may carry pre-allocated (pinned) output arrays."""
    T = np.ascontiguousarray(T, np.int32); B = T.shape[0]
    row_off = np.ascontiguousarray(row_off, np.int64); tgt_off = np.ascontiguousarray(tgt_off, np.int64)
    tgt = np.ascontiguousarray(tgt, np.int32)
    assert log_probs.dtype == np.float32 and log_probs.flags.c_contiguous
    frame_off = np.zeros(B + 1, np.int64); np.cumsum(T, out=frame_off[1:])
    total = int(frame_off[-1])
    N = np.diff(tgt_off)
    max_T, max_N = (int(T.max()), int(N.max())) if B else (0, 0)
    if max_stamps is None:
        max_stamps = (max_N + 8) if params.ignore_noise else max(max_T, 1)   # see BatchPlan; check status & ST_STAMP_OVERFLOW
    shape = BfaShape(B, C_, max_T, max_N, total, int(max_stamps), 0)
    o = out if out is not None else {}

    def buf(name, shp, dt):
        a = o.get(name)
        if a is None or a.shape != tuple(shp) or a.dtype != dt:
            a = np.empty(shp, dt); o[name] = a
        return a

    frame_ph = buf("frame_ph", (max(total, 1),), np.int32); frame_idx = buf("frame_idx", (max(total, 1),), np.int32)
    dp_final = buf("dp_final", (max(B, 1),), np.float32); status = buf("status", (max(B, 1),), np.int32)
    stamps = buf("stamps", (max(B, 1), max_stamps, 4), np.int32); n_stamps = buf("n_stamps", (max(B, 1),), np.int32)
    conf = buf("conf", (max(B, 1), max_stamps), np.float32) if want_conf else None
    vp = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
    rc = _cabi.lib().bfa_align_batch_host(C.byref(params), C.byref(shape), vp(log_probs), vp(row_off), vp(T), vp(tgt), vp(tgt_off),
                                          vp(frame_ph), vp(frame_idx), vp(frame_off), vp(dp_final), vp(status), vp(stamps), vp(conf),
                                          vp(n_stamps), int(device), int(chunk_utts))
    _cabi.check(rc, host=True)
    o["frame_off"] = frame_off
    o["max_stamps"] = max_stamps
    return o
