"""What bounds the host-buffer entry (bfa_align_batch_host): pinned host->device copy rate of the headline batch, next to the
e2e step time at several chunk sizes.  Prints one JSON line.  The posteriors here are plain noise, the worst case for the
aligner (with no peak to follow, paths leave the band and utterances are redone by the exact kernel), so the e2e times are an
upper bound; bench.py's e2e leg uses the headline's planted-peaky batch.  Measured (B200 box, PCIe Gen5 x16): 49.8-50.1 GB/s
pinned H2D => the headline batch (650 MB) cannot arrive in less than 13.0 ms; bench.py's e2e step is 12.85-12.98 ms."""
import json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import importlib
bfa_b200 = importlib.import_module("bfa_b200")
B, T, N, C = 4096, 600, 40, 66
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
lp_h = torch.empty((B, T, C), dtype=torch.float32, pin_memory=True)
lp_h.normal_(generator=g); lp_h -= 4.7   # roughly normalised log-posteriors; the values do not matter for the copy rate
d = torch.empty((B, T, C), dtype=torch.float32, device=dev)
res = {}
for name, parts in (("one_copy", 1), ("8_chunks", 8)):
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(parts):
            s = slice(i * B // parts, (i + 1) * B // parts)
            d[s].copy_(lp_h[s], non_blocking=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    res["h2d_GBs_" + name] = lp_h.numel() * 4 / min(ts) / 1e9
tgt = np.tile(np.arange(1, N + 1, dtype=np.int32), B)
tgt_h = torch.from_numpy(tgt).pin_memory()
ms = N + 8
out = {k: torch.empty(s, dtype=dt, pin_memory=True).numpy() for k, s, dt in (
    ("frame_ph", (B * T,), torch.int32), ("frame_idx", (B * T,), torch.int32), ("dp_final", (B,), torch.float32),
    ("status", (B,), torch.int32), ("stamps", (B, ms, 4), torch.int32), ("n_stamps", (B,), torch.int32), ("conf", (B, ms), torch.float32))}
ro = np.arange(B, dtype=np.int64) * T * C; Th = np.full(B, T, np.int32); to = np.arange(B + 1, dtype=np.int64) * N
au = bfa_b200.AlignmentUtils(blank_id=C - 1, silence_id=0, silence_anchors=10, ignore_noise=True, truly_forced=True)
params = au.viterbi_decoder._params(True, True, True)
for cu in (256, 512, 1024, 2048):
    run = lambda: bfa_b200.align_host(params, lp_h.numpy(), ro, Th, C, tgt_h.numpy(), to, max_stamps=ms, device=0, chunk_utts=cu, out=out)
    for _ in range(2): run()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): run()
    torch.cuda.synchronize()
    res[f"e2e_ms_chunk{cu}"] = (time.perf_counter() - t0) / 5 * 1e3
print(json.dumps(res))
