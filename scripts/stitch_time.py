"""Development: device time of stitch_log_softmax (window cross-fade + log-softmax in one pass) on a metric-sized batch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bfa_b200
B, W, fpw, C = 2048, 149, 10, 67          # 12 s of audio per utterance: 149 windows of 160 ms at a stride of 80 ms
x = torch.randn(B, W, fpw, C, device="cuda")
n = 16000 * 12
for _ in range(3): y = bfa_b200.stitch_log_softmax(x, n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y = bfa_b200.stitch_log_softmax(x, n)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"stitch_log_softmax [{B},{W},{fpw},{C}] -> {tuple(y.shape)}: {ms:.4f} ms  ({(x.numel() + y.numel()) * 4 / ms / 1e6:.0f} GB/s of input + output)")
