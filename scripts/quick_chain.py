"""Development: device-timed steps of the planner-chain workloads (SIL-bearing metric shape, BASELINE configs 3 and 4) for the
library BFA_B200_LIB points at; per-kernel times come from scripts/variant_launches.py under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bfa_b200
from bfa_b200 import _cabi, synth

dev = torch.device("cuda:0")
dec = bfa_b200.AlignmentUtils(65, 0).viterbi_decoder
which = sys.argv[1:] or ["sil", "3", "4"]
out = []
for c in which:
    if c == "sil":
        lpv, tgv, _ = synth.planted_batch(4096, 600, 40, 66, seed=6001, peak=10.0, sil_every=10, sil_frames=15, device=dev)
        w = dict(lp=lpv, row_off=torch.arange(4096, dtype=torch.int64, device=dev) * 600 * 66, Ts=[600] * 4096, Ns=[40] * 4096,
                 tgt=tgv.to(torch.int32).reshape(-1).contiguous())
    else:
        w = synth.baseline_config(int(c), C=66, device=dev)
    p = dec._params(True, True, True)
    plan = dec.plan_batch(w["Ts"], w["Ns"], 66, params=p, device=dev)
    res = None
    for _ in range(5):
        res = dec.align_batch(w["lp"], w["row_off"], w["Ts"], 66, w["tgt"], w["Ns"], params=p, plan=plan, out=res)
    torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            res = dec.align_batch(w["lp"], w["row_off"], w["Ts"], 66, w["tgt"], w["Ns"], params=p, plan=plan, out=res)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 10)
    B = len(w["Ts"])
    ok = int((res.status[:B] & 7 != 2).sum())
    out.append(f"{c}: " + "/".join(f"{t:.4f}" for t in ts) + f" ms (checksum {int(res.frame_ph.sum())})")
    del w, res
    torch.cuda.empty_cache()
print(os.path.basename(os.environ.get("BFA_B200_LIB", "default")), " | ".join(out))
