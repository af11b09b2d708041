#!/usr/bin/env python
"""Secondary measurements: the BASELINE.json configs other than the headline one, device-timed through
ViterbiDecoder.align_batch (inputs resident in HBM), each with a parity sample against the oracle.
    config 2: B=1024, T=600, N=40      config 3: B=256, T=3600, N=200 with SIL anchors (segmentation on the device)
    config 4: ragged B=8192, T in [60,1800], N in [4,120], rows packed with row_off
Prints one JSON line per config (not bench.py's contract line; these are evidence for DESIGN.md)."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bfa_b200
from bfa_b200 import synth, _cabi
from oracle import oracle as orc

dev = torch.device("cuda:0")
Cc = 66
au = bfa_b200.AlignmentUtils(blank_id=Cc - 1, silence_id=0, silence_anchors=10, ignore_noise=True, truly_forced=True)
dec = au.viterbi_decoder
lib = _cabi.lib()
STEPS = int(os.environ.get("STEPS", "10"))


def run(name, lp_flat, row_off, Ts, tgt32, Ns, sample, lp_of, tgt_of):
    B = len(Ts)
    p = dec._params(True, True, True)
    plan = dec.plan_batch(Ts, Ns, Cc, params=p, device=dev)
    res = [None]
    def step():
        res[0] = dec.align_batch(lp_flat, row_off, Ts, Cc, tgt32, Ns, params=p, want_stamps=True, want_conf=True, plan=plan, out=res[0])
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    l0 = lib.bfa_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    r = res[0]
    st = r.status[:B].cpu().numpy()
    frames = int(np.sum(Ts))
    # parity sample against the oracle (bit-exact frames, stamps; 1e-4 on confidences)
    po = orc.params(Cc - 1, 0)
    fo = np.zeros(B + 1, np.int64); np.cumsum(np.asarray(Ts, np.int64), out=fo[1:])
    fph, fix = r.frame_ph.cpu().numpy(), r.frame_idx.cpu().numpy()
    nst, stamps, conf = r.n_stamps.cpu().numpy(), r.stamps.cpu().numpy(), r.conf.cpu().numpy()
    bad = 0
    for u in sample:
        lp_u = lp_of(u).cpu().numpy(); tg_u = tgt_of(u).cpu().numpy().astype(np.int32)
        T, N = int(Ts[u]), int(Ns[u])
        o = orc.align_batch(po, lp_u.reshape(1, T, Cc), np.zeros(1, np.int64), np.asarray([T], np.int32), Cc, tg_u,
                            np.asarray([0, N], np.int64), max_stamps=plan.max_stamps, n_threads=1)
        ok = (int(o["status"][0]) & 15) == (int(st[u]) & 15)
        if (int(st[u]) & 7) != 2:   # TOO_SHORT has no frames
            ok = ok and np.array_equal(fph[fo[u]:fo[u + 1]], o["frame_ph"]) and np.array_equal(fix[fo[u]:fo[u + 1]], o["frame_idx"])
            n = int(o["n_stamps"][0]); ok = ok and n == int(nst[u])
            if ok and n:
                want = np.stack([o["stamps"][0][f][:n] for f in ("phoneme", "start", "end", "target_idx")], 1)
                ok = ok and np.array_equal(stamps[u, :n], want) and np.allclose(conf[u, :n], o["conf"][0, :n], rtol=1e-4, atol=1e-6)
        bad += 0 if ok else 1
    ic = (C.c_int32 * 4)(); lib.bfa_debug_item_counts(ic)
    out = {"config": name, "items_exact_kernel": int(ic[0]), "items_banded_24_40_64": [int(ic[1]), int(ic[2]), int(ic[3])], "B": B, "frames": frames, "input_MB": round(lp_flat.numel() * 4 / 1e6, 1), "ms_per_step": round(ms, 4),
           "frames_per_s": frames / (ms / 1e3), "launches_per_step": (lib.bfa_launch_count() - l0) / STEPS,
           "status_counts": {str(k): int(v) for k, v in zip(*np.unique(st & 15, return_counts=True))},
           "oracle_sample": len(sample), "oracle_mismatches": bad}
    print(json.dumps(out), flush=True)
    return out


# ---- config 2 ----
B, T, N = 1024, 600, 40
lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=21, device=dev)
run("2: B=1024 T=600 N=40", lp, torch.arange(B, dtype=torch.int64, device=dev) * T * Cc, [T] * B, tgt.to(torch.int32).reshape(-1).contiguous(),
    [N] * B, list(range(0, B, 128)), lambda u: lp[u], lambda u: tgt[u])
del lp
# ---- config 3 ----
B, T, N = 256, 3600, 200
lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=22, peak=12.0, sil_every=40, sil_frames=18, device=dev)
run("3: B=256 T=3600 N=200 SIL anchors", lp, torch.arange(B, dtype=torch.int64, device=dev) * T * Cc, [T] * B,
    tgt.to(torch.int32).reshape(-1).contiguous(), [N] * B, list(range(0, B, 43)), lambda u: lp[u], lambda u: tgt[u])
del lp
# ---- config 4 ----
B = int(os.environ.get("RAGGED_B", "8192"))
t0 = time.time()
utts = synth.ragged_batch(B, C=Cc, seed=23, device=dev)
if os.environ.get("RAGGED_SORT", "1") == "1":      # what a length-bucketing loader does: the four utterances of a task are alike
    utts.sort(key=lambda u: -int(u[0].shape[0]))
Ts = [int(l.shape[0]) for l, _ in utts]; Ns = [int(t.shape[0]) for _, t in utts]
# rows packed back to back; every utterance starts on a 16-byte boundary (T*C*4 is a multiple of 8 at C=66: pad odd T by one row)
offs, cur = [], 0
for t in Ts:
    offs.append(cur); cur += (t * Cc + 3) // 4 * 4
flat = torch.empty(cur, dtype=torch.float32, device=dev)
for (l, _), o, t in zip(utts, offs, Ts):
    flat[o:o + t * Cc] = l.reshape(-1)
tg = torch.cat([t for _, t in utts]).to(torch.int32).contiguous()
row_off = torch.tensor(offs, dtype=torch.int64, device=dev)
print(f"# ragged corpus built in {time.time() - t0:.1f} s: {sum(Ts)} frames, {cur * 4 / 1e9:.2f} GB", flush=True)
run(f"4: ragged B={B} T in [60,1800] N in [4,120] packed" + (", utterances ordered by length" if os.environ.get("RAGGED_SORT", "1") == "1" else ""), flat, row_off, Ts, tg, Ns, list(range(0, B, max(1, B // 24))),
    lambda u: utts[u][0], lambda u: utts[u][1])
