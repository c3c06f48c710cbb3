"""Throughput of the SelfAttention2d flash kernel (forward)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200 import lunar_generate as lg
dev = torch.device("cuda:0")
for C, HW, B in ((512, 64, 4), (256, 64, 8), (64, 128, 2)):
    att = lg.SelfAttention2d(C).to(dev)
    x = torch.randn(B, C, HW, HW, device=dev)
    with torch.no_grad():
        for _ in range(2): att(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): att(x)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    N = HW * HW
    fl = 2.0 * B * N * N * (C // 8 + C) + 2.0 * B * N * C * (C + C // 4)
    print(f"SelfAttention2d C={C} N={N} B={B}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (module, incl. q/k/v convs + layout)")
