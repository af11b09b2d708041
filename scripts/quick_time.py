"""Development: device-timed steps of the metric batch (pipelined launches, isolated launches, fill phase only) for the library
BFA_B200_LIB points at.  No result checks: usable with debug builds whose results are wrong on purpose."""
import copy, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bfa_b200
from bfa_b200 import _cabi, synth

B, T, N, Cc = int(os.environ.get("QT_B", "4096")), 600, 40, int(sys.argv[1]) if len(sys.argv) > 1 else 66
dev = torch.device("cuda:0")
lib = _cabi.lib()
lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=4242, device=dev)
au = bfa_b200.AlignmentUtils(blank_id=Cc - 1, silence_id=0, silence_anchors=10, ignore_noise=True, truly_forced=True)
dec = au.viterbi_decoder
params = dec._params(True, True, True)
params.reserved |= _cabi.HINT_NO_SIL | _cabi.FLAG_DIRECT_ONLY | _cabi.FLAG_PIPELINED
row_off = torch.arange(B, dtype=torch.int64, device=dev) * (T * Cc)
tgt32 = tgt.to(torch.int32).reshape(-1).contiguous()
Ts, Ns = [T] * B, [N] * B


def run(p, n, profile):
    plan = dec.plan_batch(Ts, Ns, Cc, params=p, device=dev)
    out = None
    for _ in range(int(os.environ.get("QT_WARM", "200"))):
        out = dec.align_batch(lp, row_off, Ts, Cc, tgt32, Ns, params=p, want_stamps=True, want_conf=True, plan=plan, out=out)
    torch.cuda.synchronize()
    if profile:
        lib.bfa_profile_enable(1); lib.bfa_profile_read(None, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = dec.align_batch(lp, row_off, Ts, Cc, tgt32, Ns, params=p, want_stamps=True, want_conf=True, plan=plan, out=out)
    e1.record()
    torch.cuda.synchronize()
    if profile:
        ms, k = C.c_float(), C.c_int32()
        lib.bfa_profile_read(C.byref(ms), C.byref(k)); lib.bfa_profile_enable(0)
        return ms.value / max(k.value, 1), out
    return e0.elapsed_time(e1) / n, out


res = []
for rep in range(2):
    pip, out = run(params, 40, False)
    iso, _ = run(params, 10, True)
    pf = copy.copy(params); pf.reserved |= 16
    fill, _ = run(pf, 10, True)
    res.append((pip, iso, fill))
ok = int((out.status[:B] & 7 == 0).sum())
print(os.path.basename(os.environ.get("BFA_B200_LIB", "default")), " ".join(f"pipelined {a:.4f} isolated {b:.4f} fill {c:.4f} |" for a, b, c in res), f"ok {ok}/{B}")
