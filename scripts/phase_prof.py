"""Development: run the metric batch through the -DBFA_PHASE_PROF build and print warp-clock cycles per phase of the
banded kernel (BFA_B200_LIB must point at libbfa_b200_prof.so)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bfa_b200
from bfa_b200 import synth, _cabi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
CONF = (sys.argv[2] != 'noconf') if len(sys.argv) > 2 else True
T, N, Cc = 600, 40, 66
dev = torch.device("cuda:0")
lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=1, peak=8.0, device=dev)
au = bfa_b200.AlignmentUtils(blank_id=Cc - 1, silence_id=0, silence_anchors=10, ignore_noise=True, truly_forced=True)
lens = torch.full((B,), T, dtype=torch.long); nl = torch.full((B,), N, dtype=torch.long)
for it in range(3):
    au.decode_alignments(lp, true_seqs=tgt.to(dev), pred_lens=lens, true_seqs_lens=nl, with_confidence=CONF)
    torch.cuda.synchronize()
    out = (C.c_ulonglong * 32)()
    _cabi.lib().bfa_debug_phases(out, 1)
    wout = (C.c_ulonglong * 32)()
    _cabi.lib().bfa_debug_warps(wout, 1)
    cout = (C.c_ulonglong * 320)()
    _cabi.lib().bfa_debug_ctas(cout, 1)
    fout = (C.c_ulonglong * 32)()
    _cabi.lib().bfa_debug_fin(fout, 1)
names = ["setup", "slide", "barrier wait", "wait for the plan", "frames", "loop end", "pre-walk", "rec wait", "walk", "pair sync", "finish", "flush+sync", "end sync"]
tot = sum(out[:14]); nw = (B + 3) // 4
print(f"warps {nw}; cycles per warp-task {tot / nw:.0f} = {tot / nw / 1.965e3:.1f} us")
for n, v in zip(names, out):
    print(f"{n:14s} {v / nw:10.0f} cyc/task  {100 * v / tot:5.1f}%")
hn = ["wait for walk", "finish", "end sync", "-", "-", "fill: full wait", "fill: row stats", "fill: free wait+issue", "plan", "task+tables", "plan: metadata+rules", "plan: weight table", "plan: targets", "plan: items"]
print("helper warp:")
for n, v in zip(hn, out[16:30]):
    print(f"{n:14s} {v / nw:10.0f} cyc/task")
fn = ["-", "events+legality", "-", "-", "pdl wait", "frame sweep", "miss wait", "miss exp", "stamps+final"]
for w in range(2):
    print("finish, " + ("DP warp: " if w == 0 else "helper:  ") + "  ".join(f"{n} {fout[w * 16 + i] / nw:.0f}" for i, n in enumerate(fn)))
print(f"DP task cycles: max {out[14]} min {out[15]}  ({out[14]/1.965e3:.1f} / {out[15]/1.965e3:.1f} us);  helper: max {out[30]} min {out[31]} ({out[30]/1.965e3:.1f} / {out[31]/1.965e3:.1f} us)")
print("mean task us by warp id (scheduler = id % 4):", " ".join(f"w{i}:{wout[i] / max(wout[16 + i], 1) / 1.965e3:.0f}" for i in range(14)))
ct = sorted((cout[i] / 1.965e3, int(cout[160 + i]), i) for i in range(148))
print("slowest CTAs (us, smid, cta):", [(round(a), b, c) for a, b, c in ct[-12:]])
print("fastest CTAs (us, smid, cta):", [(round(a), b, c) for a, b, c in ct[:12]])
import collections
gp = collections.defaultdict(list)
for a, b, c in ct: gp[b // 2 // 10].append(a)   # rough grouping by smid / 20
print("mean of per-CTA max by smid//20:", {k: round(sum(v) / len(v), 1) for k, v in sorted(gp.items())})
