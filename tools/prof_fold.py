"""attn_fold (the as-executed Teacher attention without K / V) at C3 / C2 shapes, CUDA events."""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200 import _capi

def run(B, C, N=16384):
    dev = torch.device("cuda:0")
    lib = _capi.lib()
    nq = N // 32 + 31
    nq_pad = (nq + 7) // 8 * 8
    y = torch.randn(B, N, C, device=dev).to(torch.bfloat16)
    sc = torch.rand(C, device=dev) + 0.5; sh = torch.randn(C, device=dev)
    m2 = (torch.rand(B, C, device=dev) > 0.1).float() * 1.109375
    qt = torch.randn(B, nq_pad, 8 * C, device=dev).to(torch.bfloat16) * 0.05
    xbar = torch.empty(B, nq_pad, 8 * C, device=dev, dtype=torch.bfloat16)
    s = torch.cuda.current_stream().cuda_stream
    f = lambda: lib.lun_attn_fold_rows_bf16(y.data_ptr(), sc.data_ptr(), sh.data_ptr(), m2.data_ptr(), qt.data_ptr(),
                                             xbar.data_ptr(), B, N, C, 8, nq_pad, 12345, ctypes.c_float(0.1), s)
    for _ in range(3): assert f() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gb = (y.numel() + qt.numel() + xbar.numel()) * 2 / 1e9
    print(f"attn_fold B={B} C={C}: {ms*1e3:.1f} us  {gb/ms:.2f} TB/s")
    return float(xbar.float().abs().sum())

run(64, 512); run(16, 256)
