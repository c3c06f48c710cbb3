"""Micro-benchmark of the tcgen05 conv kernels at Teacher shapes (CUDA events, inputs > L2 at B>=16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200 import ops


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    dev = torch.device("cuda:0")
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    H = 128
    for (cin, cout, k) in ((512, 512, 3), (128, 512, 3), (512, 1536, 1), (512, 512, 1), (128, 512, 1)):
        x = torch.randn(B, H, H, cin, device=dev).to(torch.bfloat16)
        w = torch.randn(cout, cin, k, k, device=dev) * 0.02
        wp = ops.pack_conv_weight(w)
        bias = torch.zeros(cout, device=dev)
        stats = torch.zeros(2 * cout, device=dev)
        y = torch.empty(B, H, H, cout, device=dev, dtype=torch.bfloat16)
        fl = 2.0 * B * H * H * cout * cin * k * k
        ms = timeit(lambda: ops.conv2d_fprop(x, wp, k, 1, k // 2, bias=bias, act_leaky=True, stats=stats, out=y))
        print(f"fprop  B={B} {cin}->{cout} k{k}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        ms = timeit(lambda: ops.conv2d_fprop(x, wp, k, 1, k // 2, out=y))
        print(f"fprop(noepi) B={B} {cin}->{cout} k{k}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        if cout == 512:
            ms = timeit(lambda: ops.conv2d_wgrad(y, x, k, 1, k // 2))
            print(f"wgrad  B={B} {cin}->{cout} k{k}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        del x, y
    # cuDNN / cuBLAS comparison for context (library baseline, NHWC bf16)
    x = torch.randn(B, 512, H, H, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(512, 512, 3, 3, device=dev) * 0.02).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    torch.backends.cudnn.benchmark = True
    ms = timeit(lambda: torch.nn.functional.conv2d(x, w, padding=1))
    print(f"cudnn fprop B={B} 512->512 k3: {ms:.3f} ms  {2.0 * B * H * H * 512 * 512 * 9 / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
