"""proj_expand (dropout(bias) stream + the surviving rows) at C3 / C2 shapes, CUDA events."""
import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200 import _capi
lib = _capi.lib(); dev = torch.device("cuda:0")
for (B, C) in ((64, 512), (16, 256)):
    HW = 16384; nq = HW // 32 + 31; nq_pad = (nq + 7) // 8 * 8
    small = torch.randn(B, nq_pad, C, device=dev).to(torch.bfloat16); bias = torch.randn(C, device=dev)
    y = torch.empty(B, HW, C, device=dev, dtype=torch.bfloat16)
    s = torch.cuda.current_stream().cuda_stream
    f = lambda: lib.lun_proj_expand_bf16(small.data_ptr(), bias.data_ptr(), y.data_ptr(), B, HW, C, nq, nq_pad, 12345, ctypes.c_float(0.1), s)
    for _ in range(3): assert f() == 0
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 20
    print(f"proj_expand B={B} C={C}: {ms*1e3:.1f} us  {y.numel()*2/1e9/ms:.2f} TB/s  checksum {float(y.float().abs().sum()):.6e} kept {float((y != 0).float().mean()):.5f}")
