"""Seeded synthetic "planted-peaky" log-posteriors.

The reference treats -1000.0 as minus infinity (forced_alignment.py:23), so flat
random posteriors are invalid inputs (every path scores below the sentinel).
Valid synthetic inputs plant a monotone alignment: 2N+1 slots
(blank, ph_0, blank, ph_1, ... blank) with random durations summing to T, a
+peak logit on the planted class of each frame, then log_softmax -- the same
thing core.py:898-899 produces from CUPE logits.

Pure torch, vectorised over the batch, runs on CPU or CUDA.  Used by tests,
bench.py and smoke(); it has no dependency on oracle/.
"""
from __future__ import annotations

import torch


def planted_batch(B, T, N, C, *, seed=0, peak=8.0, blank_id=None, silence_id=0, sil_every=0,
                  sil_frames=15, device="cpu", dtype=torch.float32):
    """Returns (log_probs [B,T,C] f32, targets [B,N] i64, planted [B,T] i64).

    targets are drawn from [1, blank_id) (never SIL=0, never blank) unless
    sil_every>0, in which case every sil_every-th target is silence_id and its
    slot is at least sil_frames long (drives silence-anchored segmentation).
    """
    if blank_id is None:
        blank_id = C - 1
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    tgt = torch.randint(1, blank_id, (B, N), generator=g)
    if sil_every and N > sil_every:
        tgt[:, sil_every::sil_every] = silence_id
    nslot = 2 * N + 1
    w = torch.rand(B, nslot, generator=g) + 0.2
    w[:, 0::2] *= 0.5  # blanks shorter than phonemes
    slot_class = torch.full((B, nslot), blank_id, dtype=torch.long)
    slot_class[:, 1::2] = tgt
    min_len = torch.zeros(B, nslot)
    min_len[:, 1::2] = 1.0  # every phoneme gets at least one frame
    if sil_every:
        min_len[slot_class == silence_id] = float(sil_frames)
    free = (T - min_len.sum(1, keepdim=True)).clamp_min(0.0)
    dur = min_len + w / w.sum(1, keepdim=True) * free
    edges = torch.cumsum(dur, 1)
    edges[:, -1] = T + 1.0
    frames = torch.arange(T, dtype=torch.float32).expand(B, T) + 0.5
    slot_of_frame = torch.searchsorted(edges.contiguous(), frames.contiguous(), right=False).clamp_max(nslot - 1)
    planted = torch.gather(slot_class, 1, slot_of_frame)
    planted = planted.to(device)
    gd = torch.Generator(device=device).manual_seed(int(seed) + 7919)
    logits = torch.randn(B, T, C, generator=gd, device=device, dtype=torch.float32)
    logits.scatter_add_(2, planted.unsqueeze(-1), torch.full((B, T, 1), float(peak), device=device))
    logp = torch.log_softmax(logits, dim=2).to(dtype)
    return logp, tgt.to(device), planted


def ragged_batch(B, *, C=66, t_range=(60, 1800), n_range=(4, 120), seed=0, peak=10.0, device="cpu"):
    """Ragged corpus of config 4: T~U[t_range], N~U[n_range] with N <= T/5 mostly,
    plus a tail (every 16th utterance) that lands on strides 3/2/1.
    Returns a list of (log_probs [T,C], targets [N]) per utterance."""
    g = torch.Generator().manual_seed(int(seed))
    out = []
    for u in range(B):
        T = int(torch.randint(t_range[0], t_range[1] + 1, (1,), generator=g))
        nmax = max(n_range[0], min(n_range[1], T // 5))
        N = int(torch.randint(n_range[0], nmax + 1, (1,), generator=g))
        if u % 16 == 15:  # dense tail: stride 3 / 2 / 1
            N = min(n_range[1], max(n_range[0], int(T / (1.2 + 2.5 * float(torch.rand(1, generator=g))))))
        lp, tgt, _ = planted_batch(1, T, N, C, seed=seed * 100003 + u, peak=peak, device=device)
        out.append((lp[0], tgt[0]))
    return out


def pack_ragged(utts, C, device="cpu", align_floats=4):
    """Pack a list of (log_probs [T,C], targets [N]) back to back (every utterance starts on a multiple of `align_floats`
    floats) -> (flat fp32 rows, row_off int64 [B], Ts, flat int32 targets, Ns), the arguments of ViterbiDecoder.align_batch."""
    Ts = [int(l.shape[0]) for l, _ in utts]
    Ns = [int(t.shape[0]) for _, t in utts]
    offs, cur = [], 0
    for t in Ts:
        offs.append(cur)
        cur += (t * C + align_floats - 1) // align_floats * align_floats
    flat = torch.empty(max(cur, 1), dtype=torch.float32, device=device)
    for (l, _), o, t in zip(utts, offs, Ts):
        flat[o:o + t * C] = l.reshape(-1)
    tg = torch.cat([t for _, t in utts]).to(torch.int32).contiguous() if utts else torch.zeros(0, dtype=torch.int32)
    return flat, torch.tensor(offs, dtype=torch.int64, device=device), Ts, tg.to(device), Ns


def baseline_config(n, *, C=66, device="cpu", B=None, seed=None, sort_ragged=True):
    """The synthetic workloads of BASELINE.json `configs`, as align_batch arguments:
       2: B=1024, T=600, N=40      3: B=256, T=3600, N=200 with SIL anchors (silence-anchored segmentation)
       4: ragged B=8192, T in [60,1800], N in [4,120], rows packed (ordered by length like a bucketing loader)
    Returns dict(lp=flat rows, row_off, Ts, tgt=flat int32 targets, Ns, name)."""
    if n == 2:
        B = B or 1024
        lp, tgt, _ = planted_batch(B, 600, 40, C, seed=21 if seed is None else seed, device=device)
        return dict(name=f"2: B={B} T=600 N=40", lp=lp.reshape(-1), row_off=torch.arange(B, dtype=torch.int64, device=device) * 600 * C,
                    Ts=[600] * B, tgt=tgt.to(torch.int32).reshape(-1).contiguous(), Ns=[40] * B)
    if n == 3:
        B = B or 256
        lp, tgt, _ = planted_batch(B, 3600, 200, C, seed=22 if seed is None else seed, peak=12.0, sil_every=40, sil_frames=18, device=device)
        return dict(name=f"3: B={B} T=3600 N=200 SIL anchors", lp=lp.reshape(-1), row_off=torch.arange(B, dtype=torch.int64, device=device) * 3600 * C,
                    Ts=[3600] * B, tgt=tgt.to(torch.int32).reshape(-1).contiguous(), Ns=[200] * B)
    if n == 4:
        B = B or 8192
        utts = ragged_batch(B, C=C, seed=23 if seed is None else seed, device=device)
        if sort_ragged:
            utts.sort(key=lambda u: -int(u[0].shape[0]))
        flat, row_off, Ts, tg, Ns = pack_ragged(utts, C, device=device)
        return dict(name=f"4: ragged B={B} T in [60,1800] N in [4,120] packed" + (", ordered by length" if sort_ragged else ""),
                    lp=flat, row_off=row_off, Ts=Ts, tgt=tg, Ns=Ns)
    raise ValueError(n)


class FakePhonemizer:
    """Deterministic stand-in for ipamappers/ph66_phonemeizer.Phonemizer (espeak is not part of the path): the same record
    layout (ph66 / pg16 / eipa / mipa / words / word_num / text, ph66_phonemeizer.py:347-470) and label tables; every word maps
    to 2-5 phoneme ids derived from its letters, commas and full stops become SIL like the reference's <SIL> words."""
    phonemes_key, phoneme_groups_key, phonemes_ipa_key = "ph66", "pg16", "eipa"

    def __init__(self, n_ph=66, n_grp=16):
        self.index_to_plabel = {0: "SIL", **{i: f"p{i}" for i in range(1, n_ph)}, n_ph: "noise"}
        self.index_to_glabel = {0: "SIL", **{i: f"g{i}" for i in range(1, n_grp)}, n_grp: "noise"}
        self.phoneme_id_to_group_id = {0: 0, **{i: 1 + (i - 1) % (n_grp - 1) for i in range(1, n_ph)}, n_ph: n_grp}
        self.n_ph = n_ph

    def phonemize_sentence(self, text):
        words = text.replace(",", " <SIL> ").replace(".", " <SIL> ").split()
        out = {"ph66": [], "pg16": [], "eipa": [], "mipa": [], "word_num": [], "words": [], "text": text}
        for wi, w in enumerate(words):
            if w == "<SIL>":
                ids = [0]
                out["words"].append("<sil>")
            else:
                h = sum((i + 1) * ord(ch) for i, ch in enumerate(w))
                ids = [1 + (h * (k + 3) + 7 * k) % (self.n_ph - 1) for k in range(2 + h % 4)]
                out["words"].append(w)
            for p in ids:
                out["ph66"].append(p); out["pg16"].append(self.phoneme_id_to_group_id[p])
                out["eipa"].append(self.index_to_plabel[p]); out["mipa"].append(self.index_to_plabel[p]); out["word_num"].append(wi)
        return out


def planted_logits(targets, T, C, *, seed=0, peak=9.0, sil_frames=12, blank_id=None):
    """[T, C] raw logits (not normalised) whose frame-wise peaks follow `targets` in order, target 0 (SIL) held for
    `sil_frames` frames, the rest spread evenly with blank frames in between: what an acoustic model would hand to
    core.py:898 for an utterance that really contains the targets."""
    blank_id = C - 1 if blank_id is None else blank_id
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(T, C, generator=g)
    n = len(targets)
    n_sil = sum(1 for t in targets if t == 0)
    per = max(2, (T - n_sil * sil_frames) // max(n - n_sil, 1))
    t = 0
    for ph in targets:
        dur = sil_frames if ph == 0 else per
        hold = dur if ph == 0 else max(1, dur // 2)
        for f in range(t, min(t + dur, T)):
            logits[f, ph if f - t < hold else blank_id] += peak
        t += dur
    for f in range(t, T):
        logits[f, blank_id] += peak
    return logits


class PlantedPosteriorProvider:
    """Stands where PhonemeTimestampAligner._cupe_prediction_batch (core.py:370-460) stands in tests: logits planted along each
    utterance's own targets (set `pending` to the batch's phoneme sequences before the call), 20 ms frames, seeded call by
    call.  Returns (logits_class [B, T, 67], logits_group [B, T, 17], None, spectral_lens) on the CPU."""

    def __init__(self, phoneme_id_to_group_id, seed):
        self.p2g, self.seed, self.n_calls, self.pending = phoneme_id_to_group_id, seed, 0, None

    def __call__(self, wavs, wav_lens, extract_embeddings=False):
        B = wavs.shape[0]
        spectral = [max(8, int(wl) // 320) for wl in wav_lens]
        T = max(spectral)
        lp, lg = torch.zeros(B, T, 67), torch.zeros(B, T, 17)
        for b in range(B):
            tg = [int(p) for p in self.pending[b]]
            lp[b, :spectral[b]] = planted_logits(tg, spectral[b], 67, seed=self.seed + 10 * self.n_calls + b)
            lg[b, :spectral[b]] = planted_logits([self.p2g[p] for p in tg], spectral[b], 17, seed=self.seed + 10 * self.n_calls + b + 5)
        self.n_calls += 1
        return lp, lg, None, spectral
