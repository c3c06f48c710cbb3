"""Tiny driver for ncu: a few launches of the conv fprop (and optionally wgrad) kernel at one Teacher shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200 import ops

cin, cout, k, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
mode = sys.argv[5] if len(sys.argv) > 5 else "stats"
dev = torch.device("cuda:0")
x = torch.randn(B, 128, 128, cin, device=dev).to(torch.bfloat16)
wp = ops.pack_conv_weight(torch.randn(cout, cin, k, k, device=dev) * 0.02)
bias = torch.zeros(cout, device=dev)
stats = torch.zeros(2 * cout, device=dev) if "stats" in mode else None
y = torch.empty(B, 128, 128, cout, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.conv2d_fprop(x, wp, k, 1, k // 2, bias=bias, act_leaky=True, stats=stats, out=y)
    if "wgrad" in mode:
        ops.conv2d_wgrad(y, x, k, 1, k // 2)
torch.cuda.synchronize()
print("done")
