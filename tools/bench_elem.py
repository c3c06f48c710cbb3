"""Micro-benchmark of the bandwidth-bound Teacher kernels at C3 shapes (B x 16384 x 512 bf16), CUDA events."""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200 import _capi, lunar_evaluator as le

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    C, HW = 512, 16384
    dev = torch.device("cuda:0")
    lib = _capi.lib()
    x = torch.randn(B, HW, C, device=dev).to(torch.bfloat16)
    idn = torch.randn(B, HW, C, device=dev).to(torch.bfloat16)
    sc = torch.rand(C, device=dev) + 0.5; sh = torch.randn(C, device=dev)
    m2 = le._drop2d_mask(B, C, 0.1, dev, "bench"); ls = torch.full((C,), 0.1, device=dev)
    pool = torch.zeros(B, C, device=dev)
    gb = x.numel() * 2 / 1e9
    ms = timeit(lambda: le._affine(x, B, HW, C, sc, sh, mask2d=m2))
    print(f"affine bn+drop2d          : {ms:.3f} ms  {2*gb/ms:.0f} GB/s")
    ms = timeit(lambda: le._affine(x, B, HW, C, sc, sh, mask2d=m2, ls=ls, identity=idn, pool=pool))
    print(f"affine block epilogue+pool: {ms:.3f} ms  {3*gb/ms:.0f} GB/s")
    mean = torch.zeros(C, device=dev); rstd = torch.ones(C, device=dev); gamma = torch.ones(C, device=dev)
    x2 = torch.randn(B, HW, C, device=dev).to(torch.bfloat16)     # three distinct streams, as in the step
    def bwd():
        return le._block_tail_backward(B, HW, C, x, None, idn, x2, mean, rstd, gamma, ls, m2, 0.2, 0.2, want_dpre=True)
    ms = timeit(bwd)
    print(f"block tail backward (2 k) : {ms:.3f} ms  {7*gb/ms:.0f} GB/s (r3+w1, r2+w1)")
    qkv = torch.randn(B, HW, 3 * C, device=dev).to(torch.bfloat16)
    nq, nq_pad = 543, 544
    att = torch.zeros(B, nq_pad, C, device=dev, dtype=torch.bfloat16)
    s = torch.cuda.current_stream().cuda_stream
    ms = timeit(lambda: lib.lun_attn_ref_rows_bf16(qkv.data_ptr(), att.data_ptr(), B, HW, C, 8, nq_pad, 1, ctypes.c_float(0.1), s))
    print(f"attn_ref_rows             : {ms:.3f} ms  {2*gb/ms:.0f} GB/s (K,V read)")
    bias = torch.zeros(C, device=dev)
    ms = timeit(lambda: lib.lun_proj_expand_bf16(att.data_ptr(), bias.data_ptr(), x.data_ptr(), B, HW, C, nq, nq_pad, 1, ctypes.c_float(0.1), s))
    print(f"proj_expand               : {ms:.3f} ms  {gb/ms:.0f} GB/s (write)")
    db = torch.zeros(C, device=dev)
    ms = timeit(lambda: lib.lun_proj_bwd_gather_bf16(x.data_ptr(), att.data_ptr(), db.data_ptr(), B, HW, C, nq, nq_pad, 1, ctypes.c_float(0.1), s))
    print(f"proj_bwd_gather           : {ms:.3f} ms  {gb/ms:.0f} GB/s (read)")
    a = torch.empty(1 << 30, device=dev, dtype=torch.bfloat16); b2 = torch.empty_like(a)
    ms = timeit(lambda: b2.copy_(a))
    print(f"torch copy 2 GiB          : {ms:.3f} ms  {a.numel()*4/1e6/ms:.0f} GB/s")

main()
