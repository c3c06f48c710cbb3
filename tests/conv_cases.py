"""Parity cases for the tcgen05 implicit-GEMM kernels against torch fp32 (TF32 off) on bf16-rounded operands.

Each case returns (max_abs_err, tolerance). Used by tests/test_conv_gpu.py and tools/diag_conv.py.
Reference call sites: F.conv2d / F.conv_transpose2d / F.linear in lunar_generate.py, lunar_evaluator.py.
"""
import torch
import torch.nn.functional as F

from lunaris_orion_b200 import ops


def _rand(shape, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev)


def _bf(x):
    return x.to(torch.bfloat16).float()


def _nhwc(x):  # NCHW fp32 -> NHWC bf16
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _nchw(y):  # NHWC -> NCHW fp32
    return y.float().permute(0, 3, 1, 2)


def _err(got, ref, rel):
    tol = rel * ref.abs().max().item() + 1e-6
    return (got - ref).abs().max().item(), tol


def case_conv_fprop(dev, B, H, Cin, Cout, k, stride, pad, leaky=False, stats=False, seed=1):
    x = _bf(_rand((B, Cin, H, H), dev, seed=seed))
    w = _bf(_rand((Cout, Cin, k, k), dev, (2.0 / (Cin * k * k)) ** 0.5, seed + 1))
    b = _rand((Cout,), dev, 0.1, seed + 2)
    ref = F.conv2d(x, w, b, stride=stride, padding=pad)
    if leaky:
        ref = F.leaky_relu(ref, 0.2)
    st = torch.zeros(2 * Cout, device=dev) if stats else None
    y = ops.conv2d_fprop(_nhwc(x), ops.pack_conv_weight(w), k, stride, pad, bias=b, act_leaky=leaky, stats=st)
    torch.cuda.synchronize()
    e, tol = _err(_nchw(y), ref, 1.0 / 128)
    if stats:
        yb = _nchw(y)
        s_ref = torch.cat([yb.sum((0, 2, 3)), (yb * yb).sum((0, 2, 3))])
        e2, tol2 = _err(st, s_ref, 1e-4)
        if e2 > tol2:
            return e2, tol2
    return e, tol


def case_conv_dgrad(dev, B, H, Cin, Cout, k, stride, pad, seed=2):
    OH = (H + 2 * pad - k) // stride + 1
    w = _bf(_rand((Cout, Cin, k, k), dev, (2.0 / (Cout * k * k)) ** 0.5, seed))
    dy = _bf(_rand((B, Cout, OH, OH), dev, seed=seed + 1))
    ref = torch.nn.grad.conv2d_input((B, Cin, H, H), w, dy, stride=stride, padding=pad)
    dx = ops.conv2d_dgrad(_nhwc(dy), ops.pack_conv_weight_dgrad(w), k, stride, pad, (H, H))
    torch.cuda.synchronize()
    return _err(_nchw(dx), ref, 1.0 / 128)


def case_conv_wgrad(dev, B, H, Cin, Cout, k, stride, pad, seed=3):
    OH = (H + 2 * pad - k) // stride + 1
    x = _bf(_rand((B, Cin, H, H), dev, seed=seed))
    dy = _bf(_rand((B, Cout, OH, OH), dev, seed=seed + 1))
    ref = torch.nn.grad.conv2d_weight(x, (Cout, Cin, k, k), dy, stride=stride, padding=pad)
    dw = ops.conv2d_wgrad(_nhwc(dy), _nhwc(x), k, stride, pad)
    torch.cuda.synchronize()
    return _err(dw, ref, 2e-3)


def case_convT_fprop(dev, B, H, Cin, Cout, seed=4):
    x = _bf(_rand((B, Cin, H, H), dev, seed=seed))
    w = _bf(_rand((Cin, Cout, 4, 4), dev, (1.0 / (Cin * 4)) ** 0.5, seed + 1))
    b = _rand((Cout,), dev, 0.1, seed + 2)
    ref = F.conv_transpose2d(x, w, b, stride=2, padding=1)
    y = ops.convT4x4s2_fprop(_nhwc(x), ops.pack_convT_weight(w), bias=b)
    torch.cuda.synchronize()
    return _err(_nchw(y), ref, 1.0 / 128)


def case_convT_dgrad(dev, B, H, Cin, Cout, seed=5):
    w = _bf(_rand((Cin, Cout, 4, 4), dev, (1.0 / (Cout * 4)) ** 0.5, seed))
    dy = _bf(_rand((B, Cout, 2 * H, 2 * H), dev, seed=seed + 1))
    ref = F.conv2d(dy, w, None, stride=2, padding=1)   # adjoint of conv_transpose2d
    dx = ops.convT4x4s2_dgrad(_nhwc(dy), ops.pack_convT_weight_dgrad(w))
    torch.cuda.synchronize()
    return _err(_nchw(dx), ref, 1.0 / 128)


def case_convT_wgrad(dev, B, H, Cin, Cout, seed=6):
    x = _bf(_rand((B, Cin, H, H), dev, seed=seed)).requires_grad_(False)
    dy = _bf(_rand((B, Cout, 2 * H, 2 * H), dev, seed=seed + 1))
    w = torch.zeros(Cin, Cout, 4, 4, device=dev, requires_grad=True)
    F.conv_transpose2d(x, w, None, stride=2, padding=1).backward(dy)
    dw = ops.convT4x4s2_wgrad(_nhwc(dy), _nhwc(x))
    torch.cuda.synchronize()
    return _err(dw, w.grad, 2e-3)


def case_linear(dev, B, K, N, seed=7):
    x = _bf(_rand((B, K), dev, seed=seed))
    w = _bf(_rand((N, K), dev, K ** -0.5, seed + 1))
    b = _rand((N,), dev, 0.1, seed + 2)
    dy = _bf(_rand((B, N), dev, seed=seed + 3))
    y = ops.linear_fprop(x.to(torch.bfloat16), w.to(torch.bfloat16), b)
    dx = ops.linear_dgrad(dy.to(torch.bfloat16), w.t().contiguous().to(torch.bfloat16))
    dw = ops.linear_wgrad(dy.to(torch.bfloat16), x.to(torch.bfloat16))
    torch.cuda.synchronize()
    errs = [_err(y, x @ w.t() + b, 2e-3), _err(dx.float(), dy @ w, 1.0 / 128), _err(dw, dy.t() @ x, 2e-3)]
    return max(errs, key=lambda t: t[0] / t[1])


CASES = {
    # name: (fn, kwargs)
    "gemm1x1_64": (case_conv_fprop, dict(B=2, H=8, Cin=64, Cout=64, k=1, stride=1, pad=0)),
    "gemm1x1_qkv": (case_conv_fprop, dict(B=1, H=128, Cin=256, Cout=768, k=1, stride=1, pad=0)),
    "conv3x3_teacher": (case_conv_fprop, dict(B=2, H=128, Cin=128, Cout=256, k=3, stride=1, pad=1, leaky=True, stats=True)),
    "conv3x3_512": (case_conv_fprop, dict(B=1, H=128, Cin=512, Cout=512, k=3, stride=1, pad=1, leaky=True, stats=True)),
    "conv3x3_8x8": (case_conv_fprop, dict(B=6, H=8, Cin=512, Cout=512, k=3, stride=1, pad=1)),
    "conv3x3_odd_batch": (case_conv_fprop, dict(B=3, H=8, Cin=64, Cout=64, k=3, stride=1, pad=1)),
    "conv3x3_s2": (case_conv_fprop, dict(B=2, H=64, Cin=64, Cout=128, k=3, stride=2, pad=1)),
    "conv3x3_s2_small": (case_conv_fprop, dict(B=4, H=16, Cin=256, Cout=512, k=3, stride=2, pad=1)),
    "conv_n32": (case_conv_fprop, dict(B=2, H=32, Cin=64, Cout=32, k=3, stride=1, pad=1)),
    "dgrad3x3": (case_conv_dgrad, dict(B=2, H=128, Cin=256, Cout=256, k=3, stride=1, pad=1)),
    "dgrad3x3_s2": (case_conv_dgrad, dict(B=2, H=64, Cin=64, Cout=128, k=3, stride=2, pad=1)),
    "wgrad3x3": (case_conv_wgrad, dict(B=2, H=128, Cin=128, Cout=256, k=3, stride=1, pad=1)),
    "wgrad3x3_small": (case_conv_wgrad, dict(B=4, H=8, Cin=512, Cout=512, k=3, stride=1, pad=1)),
    "wgrad3x3_s2": (case_conv_wgrad, dict(B=2, H=64, Cin=64, Cout=128, k=3, stride=2, pad=1)),
    "wgrad1x1": (case_conv_wgrad, dict(B=2, H=128, Cin=128, Cout=512, k=1, stride=1, pad=0)),
    "convT_fprop": (case_convT_fprop, dict(B=2, H=8, Cin=512, Cout=256)),
    "convT_fprop_64": (case_convT_fprop, dict(B=2, H=64, Cin=64, Cout=32)),
    "convT_dgrad": (case_convT_dgrad, dict(B=2, H=16, Cin=256, Cout=128)),
    "convT_wgrad": (case_convT_wgrad, dict(B=2, H=16, Cin=256, Cout=128)),
    "convT_wgrad_64": (case_convT_wgrad, dict(B=2, H=64, Cin=64, Cout=32)),
    "linear_fc_mu": (case_linear, dict(B=8, K=32768, N=512)),
    "linear_dec_fc": (case_linear, dict(B=8, K=256, N=32768)),
}
