"""Development: the host-buffer entry (bfa_align_batch_host) on the metric batch with different chunk sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bfa_b200
from bfa_b200 import synth, _cabi
B, T, N, Cc = 4096, 600, 40, 66
dev = torch.device("cuda:0")
lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=4242, device=dev)
dec = bfa_b200.AlignmentUtils(Cc - 1, 0).viterbi_decoder
params = dec._params(True, True, True)
params.reserved |= _cabi.HINT_NO_SIL | _cabi.FLAG_DIRECT_ONLY | _cabi.FLAG_PIPELINED
lp_h = torch.empty((B, T, Cc), dtype=torch.float32, pin_memory=True); lp_h.copy_(lp)
tgt_h = torch.empty(B * N, dtype=torch.int32, pin_memory=True); tgt_h.copy_(tgt.to(torch.int32).reshape(-1))
ms = N + 8
out = {k: torch.empty(s, dtype=d, pin_memory=True).numpy() for k, s, d in (
    ("frame_ph", (B * T,), torch.int32), ("frame_idx", (B * T,), torch.int32), ("dp_final", (B,), torch.float32),
    ("status", (B,), torch.int32), ("stamps", (B, ms, 4), torch.int32), ("n_stamps", (B,), torch.int32), ("conf", (B, ms), torch.float32))}
ro_h = np.arange(B, dtype=np.int64) * T * Cc; T_h = np.full(B, T, np.int32); to_h = np.arange(B + 1, dtype=np.int64) * N
for chunk in [int(x) for x in (sys.argv[1:] or ["128", "256", "512", "1024", "2048"])]:
    run = lambda: bfa_b200.align_host(params, lp_h.numpy(), ro_h, T_h, Cc, tgt_h.numpy(), to_h, max_stamps=ms, device=0, chunk_utts=chunk, out=out)
    for _ in range(3): run()
    t0 = time.perf_counter()
    for _ in range(8): run()
    torch.cuda.synchronize()
    print(f"chunk {chunk}: {(time.perf_counter() - t0) / 8 * 1e3:.3f} ms/step", flush=True)
    _cabi.lib().bfa_host_release()
d_in = torch.empty_like(lp)
for _ in range(2): d_in.copy_(lp_h, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(8): d_in.copy_(lp_h, non_blocking=True)
torch.cuda.synchronize(); print(f"plain H2D only: {(time.perf_counter() - t0) / 8 * 1e3:.3f} ms")
