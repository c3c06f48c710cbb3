"""Can an HBM-bound pass hide under the tensor-bound conv? Stream A: 3x3 512->512 convs (persistent tcgen05 kernel, one
CTA per SM); stream B: HBM-bound kernels without shared memory (proj_expand write pass, a plain device copy).
Prints alone / alone / concurrent times."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200 import ops, _capi

dev = torch.device("cuda:0")
B, H, C = 64, 128, 512
lib = _capi.lib()
x = torch.randn(B, H, H, C, device=dev).to(torch.bfloat16)
wp = ops.pack_conv_weight(torch.randn(C, C, 3, 3, device=dev) * 0.02)
bias = torch.zeros(C, device=dev)
stats = torch.zeros(2 * C, device=dev)
y = torch.empty(B, H, H, C, device=dev, dtype=torch.bfloat16)
src = torch.randn(B, H * H, C, device=dev).to(torch.bfloat16)
dst = torch.empty_like(src)
nq, nq_pad = 543, 544
ps = torch.randn(B, nq_pad, C, device=dev).to(torch.bfloat16)
h2 = torch.empty(B, H * H, C, device=dev, dtype=torch.bfloat16)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
NA, NB = 8, 24


def run_a():
    with torch.cuda.stream(sa):
        for _ in range(NA):
            ops.conv2d_fprop(x, wp, 3, 1, 1, bias=bias, act_leaky=True, stats=stats, out=y)


def run_b(kind):
    with torch.cuda.stream(sb):
        for _ in range(NB):
            if kind == "copy":
                dst.copy_(src)
            else:
                _capi.check(lib.lun_proj_expand_bf16(ps.data_ptr(), bias.data_ptr(), h2.data_ptr(), B, H * H, C, nq, nq_pad,
                                                     7, ctypes.c_float(0.1), sb.cuda_stream), "expand")


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    sa.wait_event(e0); sb.wait_event(e0)
    fn()
    main.wait_stream(sa); main.wait_stream(sb)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for kind in ("copy", "expand"):
    for _ in range(2):
        timed(lambda: (run_a(), run_b(kind)))
    ta = timed(run_a)
    tb = timed(lambda: run_b(kind))
    tab = timed(lambda: (run_a(), run_b(kind)))
    tba = timed(lambda: (run_b(kind), run_a()))
    print(f"{kind}: conv x{NA} alone {ta:.2f} ms | hbm x{NB} alone {tb:.2f} ms | concurrent (conv first) {tab:.2f} ms, "
          f"(hbm first) {tba:.2f} ms | serial sum {ta + tb:.2f} ms")
