"""Development: in-situ CUDA-event times of the planner, the dominant banded kernel and the stamp kernel on the metric batch."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bfa_b200
from bfa_b200 import synth, _cabi
B, T, N, Cc = 4096, 600, 40, 66
dev = torch.device("cuda:0")
lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=4242, device=dev)
dec = bfa_b200.AlignmentUtils(Cc - 1, 0).viterbi_decoder
p = dec._params(True, True, True); p.reserved |= _cabi.HINT_NO_SIL
row_off = torch.arange(B, dtype=torch.int64, device=dev) * (T * Cc); tg = tgt.to(torch.int32).reshape(-1).contiguous()
plan = dec.plan_batch([T] * B, [N] * B, Cc, params=p, device=dev); res = None
lib = _cabi.lib()
for it in range(25):
    if it == 5: lib.bfa_profile_enable(2); lib.bfa_profile_read(None, None); lib.bfa_profile_read_aux(None)
    res = dec.align_batch(lp, row_off, [T] * B, Cc, tg, [N] * B, params=p, plan=plan, out=res)
torch.cuda.synchronize()
d, n, aux = C.c_float(), C.c_int32(), (C.c_float * 2)()
lib.bfa_profile_read(C.byref(d), C.byref(n)); lib.bfa_profile_read_aux(aux); lib.bfa_profile_enable(0)
print(f"planner {aux[0]*1e3:.1f} us   band<3> {d.value/max(n.value,1)*1e3:.1f} us   stamps+confidence {aux[1]*1e3:.1f} us")
