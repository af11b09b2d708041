"""Development: one configuration of test_random_configurations_vs_oracle with and without the direct kernel."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bfa_b200
from bfa_b200 import synth, _cabi
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 3
rng = np.random.default_rng(9000 + seed)
Cc = int(rng.choice([9, 17, 33, 48, 66, 67, 72]))
t_hi = int(rng.choice([40, 200, 700, 1300]))
n_hi = int(rng.choice([3, 20, 60, 130]))
utts = synth.ragged_batch(36, C=Cc, t_range=(max(2, t_hi // 12), t_hi), n_range=(1, max(1, min(n_hi, Cc * 4))), seed=7000 + seed, peak=float(rng.choice([7.0, 10.0, 13.0])))
gaps = rng.integers(0, 4, len(utts)).tolist()
boost, floor = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
mode = int(rng.integers(0, 2))
anchors = int(rng.choice([0, 3, 10])) if mode == 0 else 0
base_shift = int(rng.integers(0, 4))
dev = torch.device("cuda:0")
B = len(utts)
Ts = [int(l.shape[0]) for l, _ in utts]; Ns = [int(t.shape[0]) for _, t in utts]
offs, cur = [], base_shift
for t, g in zip(Ts, gaps):
    offs.append(cur); cur += t * Cc + g
flat = torch.zeros(cur + 8, dtype=torch.float32)
for (l, _), o, t in zip(utts, offs, Ts):
    flat[o:o + t * Cc] = l.reshape(-1)
au = bfa_b200.AlignmentUtils(Cc - 1, 0, silence_anchors=anchors)
dec = au.viterbi_decoder
tg = torch.cat([t for _, t in utts]).to(torch.int32).contiguous()
for flag in (0, _cabi.FLAG_NO_DIRECT, _cabi.FLAG_DIRECT_ONLY):
    p = dec._params(boost, floor, anchors > 0, mode=mode)
    p.reserved |= flag
    r = dec.align_batch(flat.to(dev), torch.tensor(offs, dtype=torch.int64, device=dev), Ts, Cc, tg.to(dev), Ns, params=p)
    torch.cuda.synchronize()
    ic = (C.c_int32 * 4)(); _cabi.lib().bfa_debug_item_counts(ic)
    print("flag", flag, "counts", list(ic), "max_stamps", r.max_stamps)
    print(" status ", r.status[:B].cpu().tolist())
    print(" nstamps", r.n_stamps[:B].cpu().tolist())
    fo = np.zeros(B + 1, np.int64); np.cumsum(np.asarray(Ts, np.int64), out=fo[1:])
    for u in (4, 15):
        fp = r.frame_ph[fo[u]:fo[u + 1]].cpu().numpy(); fi = r.frame_idx[fo[u]:fo[u + 1]].cpu().numpy()
        print(f"  utt {u}: T {Ts[u]} N {Ns[u]} dp_final {float(r.dp_final[u]):.3f} frames: uniq ph {np.unique(fp)[:12]} idx range {fi.min()}..{fi.max()} first {fp[:6]} stamps0 {r.stamps[u, 0].cpu().tolist()}")
# utterance 4 alone
for flag in (0, _cabi.FLAG_NO_DIRECT):
    u = 4
    p = dec._params(boost, floor, anchors > 0, mode=mode); p.reserved |= flag
    l, t = utts[u]
    r = dec.align_batch(l.reshape(-1).contiguous().to(dev), torch.zeros(1, dtype=torch.int64, device=dev), [Ts[u]], Cc, t.to(torch.int32).to(dev), [Ns[u]], params=p)
    torch.cuda.synchronize()
    ic = (C.c_int32 * 4)(); _cabi.lib().bfa_debug_item_counts(ic)
    print("alone flag", flag, "counts", list(ic), "status", r.status[:1].cpu().tolist(), "n", r.n_stamps[:1].cpu().tolist())
