"""A/B: 3x3 512->512 data-gradient conv with / without the fused dropout-backward column sum, plus the separate
proj_bwd_gather pass it replaces (B=64, 128x128)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200 import ops, _capi
dev = torch.device("cuda:0")
B, H, C = 64, 128, 512
lib = _capi.lib()
dz = torch.randn(B, H, H, C, device=dev).to(torch.bfloat16)
wd = ops.pack_conv_weight_dgrad(torch.randn(C, C, 3, 3, device=dev) * 0.02)
out = torch.empty(B, H, H, C, device=dev, dtype=torch.bfloat16)
colsum = torch.zeros(2 * C, device=dev)
nq, nq_pad = 543, 544
dpo = torch.zeros(B, nq_pad, C, device=dev, dtype=torch.bfloat16)
db = torch.zeros(C, device=dev)
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for rep in range(2):
    a = timeit(lambda: ops.conv2d_dgrad(dz, wd, 3, 1, 1, (H, H), out=out))
    b = timeit(lambda: ops.conv2d_dgrad(dz, wd, 3, 1, 1, (H, H), out=out, drop_sum=(7, 0.1, colsum)))
    c = timeit(lambda: _capi.check(lib.lun_proj_bwd_gather_bf16(out.data_ptr(), dpo.data_ptr(), db.data_ptr(), B, H * H, C, nq, nq_pad, 7, ctypes.c_float(0.1), st), "g"))
    d = timeit(lambda: _capi.check(lib.lun_proj_bwd_gather_bf16(out.data_ptr(), dpo.data_ptr(), None, B, H * H, C, nq, nq_pad, 7, ctypes.c_float(0.1), st), "g"))
    print(f"dgrad plain {a:.3f} ms | dgrad + fused masked sum {b:.3f} ms | full gather pass {c:.3f} ms | rows-only gather {d:.3f} ms"
          f" | two-step {a + c:.3f} vs fused {b + d:.3f}")
