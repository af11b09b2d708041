"""Headline step (B=4096, T=600, N=40, C=66) launched directly vs replayed from a CUDA graph captured once: device time per step
(CUDA events around 50 steps) and host time per launch call.  Prints one JSON line."""
import json, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import importlib
bfa_b200 = importlib.import_module("bfa_b200")
from bfa_b200 import synth, _cabi
B, T, N, C = 4096, 600, 40, 66
dev = torch.device("cuda:0")
lp, tgt, _ = synth.planted_batch(B, T, N, C, seed=4242, device=dev)
dec = bfa_b200.AlignmentUtils(blank_id=C - 1, silence_id=0, silence_anchors=10, ignore_noise=True, truly_forced=True).viterbi_decoder
params = dec._params(True, True, True)
params.reserved |= _cabi.HINT_NO_SIL
row_off = torch.arange(B, dtype=torch.int64, device=dev) * (T * C)
tgt32 = tgt.to(torch.int32).reshape(-1).contiguous()
Ts, Ns = [T] * B, [N] * B
plan = dec.plan_batch(Ts, Ns, C, params=params, device=dev)
s = torch.cuda.Stream(device=dev)
res = {}
with torch.cuda.stream(s):
    r = dec.align_batch(lp, row_off, Ts, C, tgt32, Ns, params=params, plan=plan)
    direct = lambda: dec.align_batch(lp, row_off, Ts, C, tgt32, Ns, params=params, plan=plan, out=r)
    for _ in range(5): direct()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        direct()
    for name, fn in (("direct", direct), ("graph", g.replay), ("direct2", direct), ("graph2", g.replay)):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 50
        e0.record(s); t0 = time.perf_counter()
        for _ in range(K): fn()
        th = time.perf_counter() - t0
        e1.record(s); torch.cuda.synchronize()
        res[name + "_ms_per_step"] = e0.elapsed_time(e1) / K
        res[name + "_host_us_per_call"] = th / K * 1e6
print(json.dumps(res))
