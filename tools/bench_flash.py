"""Throughput of the SelfAttention2d flash kernels (forward and forward+backward)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200 import lunar_generate as lg
dev = torch.device("cuda:0")


def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for C, HW, B in ((512, 64, 4), (256, 64, 8), (64, 128, 2)):
    att = lg.SelfAttention2d(C).to(dev)
    with torch.no_grad():
        att.gamma.fill_(0.5)
    x = torch.randn(B, C, HW, HW, device=dev)
    N = HW * HW
    with torch.no_grad():
        ms = timeit(lambda: att(x))
    fl = 2.0 * B * N * N * (C // 8 + C) + 2.0 * B * N * C * (C + C // 4)
    print(f"SelfAttention2d fwd C={C} N={N} B={B}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (module, incl. q/k/v convs + layout)")
    xg = x.clone().requires_grad_(True)
    dy = torch.randn_like(x)

    def fb():
        y = att(xg)
        y.backward(dy)
    ms2 = timeit(fb)
    # autograd's attention backward: dV, dP (2C each per pair), dQ, dK (2 C/8 each) + conv dgrad/wgrad
    flb = 2.0 * B * N * N * (2 * C + 2 * (C // 8)) + 4.0 * B * N * C * (C + C // 4)
    print(f"SelfAttention2d fwd+bwd: {ms2:.3f} ms; bwd alone {ms2 - ms:.3f} ms = {flb / (ms2 - ms) / 1e9:.1f} algorithmic TFLOP/s")
