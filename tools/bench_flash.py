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

# ---- kernels alone through the C ABI (no 1x1 convs, no layout conversion)
import ctypes
from lunaris_orion_b200 import _capi
lib = _capi.lib()
st = torch.cuda.current_stream().cuda_stream
print("kernel-only (C ABI):")
for C, N, B in ((512, 4096, 4), (256, 4096, 8), (512, 16384, 1), (64, 16384, 2)):
    qk = (torch.randn(B * N, 128, device=dev) * 0.3).to(torch.bfloat16)
    qk[:, C // 8:64] = 0; qk[:, 64 + C // 8:] = 0
    v = torch.randn(B, N, C, device=dev).to(torch.bfloat16)
    x = torch.randn(B, N, C, device=dev).to(torch.bfloat16)
    dy = torch.randn(B, N, C, device=dev).to(torch.bfloat16)
    y = torch.empty_like(x); o = torch.empty_like(x); dv = torch.empty_like(x)
    lse = torch.empty(B * N, device=dev); dsum = torch.randn(B * N, device=dev) * 0.01
    dqk = torch.empty_like(qk)
    gm = torch.full((1,), 0.5, device=dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    fwd = lambda: _capi.check(lib.lun_flash_attn2d_bf16(P(qk), P(v), P(x), P(y), P(gm), B, N, C, P(o), P(lse), st), "fwd")
    fdv = lambda: _capi.check(lib.lun_flash_attn2d_dv_bf16(P(qk), P(dy), P(lse), P(gm), P(dv), B, N, C, st), "dv")
    fdqk = lambda: _capi.check(lib.lun_flash_attn2d_dqk_bf16(P(qk), P(v), P(dy), P(lse), P(dsum), P(gm), P(dqk), B, N, C, st), "dqk")
    pairs = float(B) * N * N
    for name, fn, fl in (("forward (S twice + PV)", fwd, 2 * pairs * (C + 2 * 64)),
                         ("dV (S + P^T dY)", fdv, 2 * pairs * (C + 64)),
                         ("dQ + dK (2 x (S + dP + out))", fdqk, 2 * 2 * pairs * (C + 2 * 64))):
        ms = timeit(fn, 10)
        print(f"  C={C} N={N} B={B} {name:30s}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} executed TFLOP/s (64-wide padded q/k)")
