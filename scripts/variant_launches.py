"""Development: run the SIL-bearing metric shape, BASELINE config 3 and config 4 a few steps each (to be run under
ncu --metrics gpu__time_duration.sum for a per-kernel launch list), printing markers between the workloads."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bfa_b200
from bfa_b200 import synth, _cabi
dev = torch.device("cuda:0")
Cc = 66
which = sys.argv[1:] or ["sil", "3", "4"]
dec = bfa_b200.AlignmentUtils(Cc - 1, 0).viterbi_decoder
STEPS = int(os.environ.get("STEPS", "3"))

def run(name, w):
    p = dec._params(True, True, True)
    plan = dec.plan_batch(w["Ts"], w["Ns"], Cc, params=p, device=dev)
    res = None
    for _ in range(STEPS):
        res = dec.align_batch(w["lp"], w["row_off"], w["Ts"], Cc, w["tgt"], w["Ns"], params=p, plan=plan, out=res)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.cudart().cudaProfilerStart()      # ncu --profile-from-start off: only these steps are in the launch list
    e0.record()
    for _ in range(STEPS):
        res = dec.align_batch(w["lp"], w["row_off"], w["Ts"], Cc, w["tgt"], w["Ns"], params=p, plan=plan, out=res)
    e1.record(); torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(f"{name}: {e0.elapsed_time(e1) / STEPS:.4f} ms/step", flush=True)

if "sil" in which:
    B, T, N = 4096, 600, 40
    lp, tg, _ = synth.planted_batch(B, T, N, Cc, seed=6001, peak=10.0, sil_every=10, sil_frames=15, device=dev)
    run("sil", dict(lp=lp, row_off=torch.arange(B, dtype=torch.int64, device=dev) * T * Cc, Ts=[T] * B, Ns=[N] * B, tgt=tg.to(torch.int32).reshape(-1).contiguous()))
    del lp
for n in (3, 4):
    if str(n) in which:
        torch.cuda.empty_cache()
        run(f"config{n}", synth.baseline_config(n, C=Cc, device=dev))
