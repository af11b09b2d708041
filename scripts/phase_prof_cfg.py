"""Development: phase timers (-DBFA_PHASE_PROF build, BFA_B200_LIB must point at it) of the banded kernels on a BASELINE config."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bfa_b200
from bfa_b200 import synth, _cabi
dev = torch.device("cuda:0")
if len(sys.argv) > 1 and sys.argv[1] == "sil":      # the metric shape with silence_id at every 10th target (bench.py's variant)
    lpv, tgv, _ = synth.planted_batch(4096, 600, 40, 66, seed=6001, peak=10.0, sil_every=10, sil_frames=15, device=dev)
    w = dict(lp=lpv, row_off=torch.arange(4096, dtype=torch.int64, device=dev) * 600 * 66, Ts=[600] * 4096, Ns=[40] * 4096,
             tgt=tgv.to(torch.int32).reshape(-1).contiguous())
else:
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    w = synth.baseline_config(n, C=66, device=dev)
dec = bfa_b200.AlignmentUtils(65, 0).viterbi_decoder
p = dec._params(True, True, True)
plan = dec.plan_batch(w["Ts"], w["Ns"], 66, params=p, device=dev)
res = None
lib = _cabi.lib()
for it in range(3):
    res = dec.align_batch(w["lp"], w["row_off"], w["Ts"], 66, w["tgt"], w["Ns"], params=p, plan=plan, out=res)
    torch.cuda.synchronize()
    out = (C.c_ulonglong * 32)(); lib.bfa_debug_phases(out, 1)
    wout = (C.c_ulonglong * 32)(); lib.bfa_debug_warps(wout, 1)
names = ["setup", "slide", "barrier wait", "wait for the plan", "frames", "loop end", "pre-walk", "rec wait", "walk", "pair sync", "finish", "flush+sync", "end sync"]
tot = sum(out[:14]); ntask = sum(wout[16 + i] for i in range(14))
print(f"DP warp-tasks {ntask}; DP cycles in total {tot / 1e6:.1f} M = {tot / 1.965e3 / 1e3:.1f} ms-warp; per task {tot / max(ntask, 1):.0f}")
for nm, v in zip(names, out):
    print(f"{nm:18s} {v / 1e6:10.2f} M  {100 * v / tot:5.1f}%")
hn = ["wait for walk", "finish", "end sync", "-", "-", "fill: full wait", "fill: row stats", "fill: free wait+issue", "plan", "task+tables"]
ht = sum(out[16:30])
print("helper warps:")
for nm, v in zip(hn, out[16:26]):
    print(f"{nm:22s} {v / 1e6:10.2f} M  {100 * v / max(ht, 1):5.1f}%")
print(f"DP task cycles: max {out[14]} ({out[14] / 1.965e3:.0f} us) min {out[15]} ({out[15] / 1.965e3:.0f} us)")
print("summed task us by warp id:", " ".join(f"w{i}:{wout[i] / 1.965e3:.0f}/{wout[16 + i]}" for i in range(14)))
