"""Driver for ncu / timing of attn_fold_kernel at C3 shape."""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200 import _capi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
C, N, nq_pad = 512, 16384, 544
dev = torch.device("cuda:0")
lib = _capi.lib()
y = torch.randn(B, N, C, device=dev).to(torch.bfloat16)
qt = (torch.randn(B, nq_pad, 8 * C, device=dev) * 0.05).to(torch.bfloat16)
xbar = torch.zeros(B, nq_pad, 8 * C, device=dev, dtype=torch.bfloat16)
sc = torch.rand(C, device=dev) + 0.5; sh = torch.randn(C, device=dev)
m2 = (torch.rand(B, C, device=dev) > 0.1).float() * 1.109375
s = torch.cuda.current_stream().cuda_stream
def run():
    _capi.check(lib.lun_attn_fold_rows_bf16(y.data_ptr(), sc.data_ptr(), sh.data_ptr(), m2.data_ptr(), qt.data_ptr(),
                                            xbar.data_ptr(), B, N, C, 8, nq_pad, 7, ctypes.c_float(0.1), s), "fold")
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
print("attn_fold ms", e0.elapsed_time(e1) / 5, "GB", (y.numel() + qt.numel() + xbar.numel()) * 2 / 1e9)
