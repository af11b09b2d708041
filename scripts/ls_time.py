"""Development: device time of log_softmax_rows (the window-less case of bfa_stitch_log_softmax) against torch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bfa_b200
for C in (67, 17, 66):
    x = torch.randn(4096, 600, C, device="cuda")
    for _ in range(3): y = bfa_b200.log_softmax_rows(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): y = bfa_b200.log_softmax_rows(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ref = torch.log_softmax(x, dim=2)
    print(f"C={C}: {ms:.4f} ms  {2 * x.numel() * 4 / ms / 1e6:.0f} GB/s  max err {(y - ref).abs().max().item():.2e}")
