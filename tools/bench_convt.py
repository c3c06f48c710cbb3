"""Transposed-conv (k4 s2 p1) micro-benchmark at the decoder's thin stages: fused-phase halo kernel vs the four
per-phase launches of the generic tap-list kernel (LUN_CONVT_HALO=0), with and without the fused GroupNorm sums."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200 import ops

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for cin, cout, hw in ((64, 32, 64), (128, 64, 32)):
    x = torch.randn(B, hw, hw, cin, device=dev).to(torch.bfloat16)
    w = ops.pack_convT_weight(torch.randn(cin, cout, 4, 4, device=dev) * 0.05)
    bias = torch.zeros(cout, device=dev)
    out = torch.empty(B, 2 * hw, 2 * hw, cout, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(B, 2, cout, device=dev)
    fl = 2.0 * B * hw * hw * cin * cout * 16
    byt = (x.numel() + out.numel()) * 2
    for name, halo, stats in (("halo+stats", True, st), ("halo", True, None), ("4 phases+stats", False, st), ("4 phases", False, None)):
        ops._HALO_CONVT = halo
        us = timeit(lambda: ops.convT4x4s2_fprop(x, w, bias=bias, out=out, img_stats=stats))
        print(f"convT {cin}->{cout} @{hw}x{hw} B={B} {name:16s}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  {byt / us / 1e3:7.1f} GB/s (in+out)")
