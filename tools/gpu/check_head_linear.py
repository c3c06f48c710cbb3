"""lun_head_linear_bf16 (heads independent GEMMs in one launch of the tcgen05 tap-list kernel) vs torch einsum, and
its time against the block-diagonal [C, heads*C] GEMM it replaces. Run with LUNARIS_B200_LIB pointing at a library
that exports the symbol."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lunaris_orion_b200 import _capi, ops

lib = _capi.lib()
fn = lib.lun_head_linear_bf16
fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
               ctypes.c_void_p, ctypes.c_void_p]
fn.restype = ctypes.c_int
dev = torch.device("cuda:0")
for (rows, heads, cin, cout) in ((2 * 160, 8, 64, 8), (64 * 544, 8, 512, 64), (16 * 544 + 24, 8, 256, 32)):
    if cout % 32:
        continue
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, heads * cin, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(heads, cout, cin, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    b = torch.randn(heads * cout, generator=g).to(dev)
    out = torch.empty(rows, heads * cout, device=dev, dtype=torch.bfloat16)
    s = torch.cuda.current_stream().cuda_stream
    rc = fn(x.data_ptr(), rows, heads, cin, w.data_ptr(), cout, b.data_ptr(), out.data_ptr(), s)
    assert rc == 0, rc
    ref = torch.einsum("rhk,hnk->rhn", x.float().view(rows, heads, cin), w.float()).reshape(rows, heads * cout) + b
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    print(f"head_linear rows={rows} heads={heads} cin={cin} cout={cout}: rel err {err:.2e}")
    assert err < 1e-2
    if cin == 512:
        wbd = torch.zeros(heads * cout, heads * cin, device=dev)
        for h in range(heads):
            wbd[h * cout:(h + 1) * cout, h * cin:(h + 1) * cin] = w[h].float()
        wbd = wbd.to(torch.bfloat16).contiguous()

        def t(f, n=20):
            for _ in range(3):
                f()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(n):
                f()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        t_new = t(lambda: fn(x.data_ptr(), rows, heads, cin, w.data_ptr(), cout, b.data_ptr(), out.data_ptr(), s))
        t_old = t(lambda: ops.linear_fprop(x, wbd, b, out_f32=False))
        print(f"  grouped {t_new * 1e3:.1f} us vs block-diagonal {t_old * 1e3:.1f} us")
print("head_linear ok")
