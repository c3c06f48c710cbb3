"""Host-side launchers: torch tensors in, C-ABI calls out (include/lunaris_b200.h).

Activations are NHWC bf16 contiguous tensors; packed weights are bf16 [slab][Cout][Cin]. Tap lists translate the
reference's conv2d / conv_transpose2d / linear call sites (lunar_generate.py, lunar_evaluator.py) and their
gradients into the implicit-GEMM form the kernels execute.
"""
import os

import torch

from . import _capi
from ._capi import check, int_array

_HALO_CONVT = os.environ.get("LUN_CONVT_HALO", "1") != "0"   # A/B switch for the fused-phase transposed conv
EPI_BIAS, EPI_LEAKY, EPI_STATS, EPI_OUT_F32, EPI_TANH, EPI_STATS_IMG = 1, 2, 4, 8, 16, 64


_stream = _capi.raw_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _nhwc(t):
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return t


# ------------------------------------------------------------------------------------------------ weight packing
def _pack(w, transpose):
    """fp32 [A,B,kh,kw] -> bf16 [kh*kw][A][B] (transpose False) or [kh*kw][B][A] (True): one kernel launch
    (csrc/optim.cu pack_weight_kernel). CUDA only, like every operand of the kernels."""
    A, B, kh, kw = w.shape
    if not w.is_cuda:
        raise _capi.LunarisB200Error("kernel operands are packed on the GPU: there is no CPU path")
    w = w.detach().float().contiguous()
    out = torch.empty(kh * kw, B if transpose else A, A if transpose else B, device=w.device, dtype=torch.bfloat16)
    check(_capi.lib().lun_pack_weight_bf16(w.data_ptr(), out.data_ptr(), A, B, kh * kw, 1 if transpose else 0, _stream()),
          "lun_pack_weight_bf16")
    return out


def pack_conv_weight(w):
    """[Cout,Cin,kh,kw] fp32 -> bf16 [kh*kw][Cout][Cin] (forward operand)."""
    return _pack(w, False)


def pack_conv_weight_dgrad(w):
    """[Cout,Cin,kh,kw] -> bf16 [kh*kw][Cin][Cout] (data-gradient operand)."""
    return _pack(w, True)


def pack_convT_weight(w):
    """ConvTranspose2d weight [Cin,Cout,kh,kw] -> bf16 [kh*kw][Cout][Cin] (forward operand)."""
    return _pack(w, True)


def pack_convT_weight_dgrad(w):
    """ConvTranspose2d weight [Cin,Cout,kh,kw] -> bf16 [kh*kw][Cin][Cout] (data-gradient operand)."""
    return _pack(w, False)


# ------------------------------------------------------------------------------------------------ raw launchers
def conv_taps(x, w_packed, cout, grid, in_mul, taps, out, out_hw, o_mul=1, o_ph=0, o_pw=0, o_coff=0, bias=None,
              flags=0, slope=0.2, stats=None, img_stats=None):
    """out[b, h*o_mul+o_ph, w*o_mul+o_pw, o_coff+n] = epi(sum_t x[b, h*in_mul+dy_t, w*in_mul+dx_t, :] . w[slab_t][n])."""
    _nhwc(x)
    XB, XH, XW, cin = x.shape
    GB, GH, GW = grid
    dy = int_array([t[0] for t in taps])
    dx = int_array([t[1] for t in taps])
    sl = int_array([t[2] for t in taps])
    if bias is not None:
        flags |= EPI_BIAS
        assert bias.dtype == torch.float32
    if stats is not None:
        flags |= EPI_STATS
        assert stats.dtype == torch.float32 and stats.numel() == 2 * cout
    if img_stats is not None:                    # per-image sums for GroupNorm, accumulated into [GB, 2, cout]
        assert stats is None and img_stats.dtype == torch.float32 and img_stats.numel() == 2 * cout * GB
        flags |= EPI_STATS | EPI_STATS_IMG
        stats = img_stats
    assert w_packed.dtype == torch.bfloat16 and w_packed.shape[1] == cout and w_packed.shape[2] == cin
    assert out.dtype == (torch.float32 if flags & EPI_OUT_F32 else torch.bfloat16)
    rc = _capi.lib().lun_conv_taps_bf16(
        x.data_ptr(), XB, XH, XW, cin, w_packed.data_ptr(), w_packed.shape[0], cout, GB, GH, GW, in_mul, len(taps),
        dy, dx, sl, _ptr(bias), out.data_ptr(), out_hw[0], out_hw[1], o_mul, o_ph, o_pw, out.shape[-1], o_coff,
        flags, slope, _ptr(stats), _stream())
    check(rc, "lun_conv_taps_bf16")
    return out


def wgrad_taps(dy, x, grid, taps, dw, dy_mul=1, dy_ph=0, dy_pw=0, in_mul=1):
    """dw[slab_t][m][n] += sum_grid dy[b, h*dy_mul+dy_ph, w*dy_mul+dy_pw, m] * x[b, h*in_mul+dy_t, w*in_mul+dx_t, n]."""
    _nhwc(dy), _nhwc(x)
    YB, YH, YW, cm = dy.shape
    XB, XH, XW, cn = x.shape
    GB, GH, GW = grid
    assert dw.dtype == torch.float32 and dw.shape[1] == cm and dw.shape[2] == cn and dw.is_contiguous()
    tdy = int_array([t[0] for t in taps])
    tdx = int_array([t[1] for t in taps])
    sl = int_array([t[2] for t in taps])
    rc = _capi.lib().lun_wgrad_taps_bf16(dy.data_ptr(), YB, YH, YW, cm, dy_mul, dy_ph, dy_pw, x.data_ptr(), XB, XH, XW,
                                         cn, in_mul, GB, GH, GW, len(taps), tdy, tdx, sl, dw.data_ptr(), _stream())
    check(rc, "lun_wgrad_taps_bf16")
    return dw


# ------------------------------------------------------------------------------------------------ conv2d
def _conv_taps_list(k, pad):
    return [(kh - pad, kw - pad, kh * k + kw) for kh in range(k) for kw in range(k)]


def img_stats_ok(gh, gw):
    """The fused per-image statistics need one image per 128-pixel tile of the output grid."""
    return gh * gw >= 128


def conv2d_fprop(x, w_packed, k, stride, pad, bias=None, act_leaky=False, stats=None, out=None, out_f32=False,
                 slope=0.2, img_stats=None):
    """F.conv2d on NHWC bf16 (lunar_evaluator.py:242; lunar_generate.py:36,95). Optional fused bias, LeakyReLU and
    per-channel batch statistics (sum, sum of squares) for the BatchNorm that follows in the reference."""
    B, H, W, _ = x.shape
    cout = w_packed.shape[1]
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    if out is None:
        out = torch.empty(B, OH, OW, cout, device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    flags = (EPI_LEAKY if act_leaky else 0) | (EPI_OUT_F32 if out_f32 else 0)
    return conv_taps(x, w_packed, cout, (B, OH, OW), stride, _conv_taps_list(k, pad), out, (OH, OW), bias=bias,
                     flags=flags, slope=slope, stats=stats, img_stats=img_stats)


def drop_sum_ok(B, H, W, C):
    """The fused dropout-backward column sum of conv2d_dgrad works on 128-pixel row tiles with 32-bit element indices."""
    return W % 128 == 0 and B * H * W * C < 2 ** 32


def conv2d_dgrad(dy, w_packed_dgrad, k, stride, pad, in_hw, out=None, drop_sum=None):
    """Gradient of F.conv2d w.r.t. its input. dy: [B,OH,OW,Cout] bf16; returns [B,H,W,Cin] bf16.
    drop_sum = (seed, p, colsum fp32 [2*Cin]): the epilogue also accumulates colsum[:Cin] += sum over pixels of
    dropout_mask(seed) * bf16(dx / (1-p)) - the backward of an elementwise Dropout that consumed dx's forward twin
    (proj_drop, lunar_evaluator.py:225) - so no separate pass reads dx for the bias gradient."""
    B, OH, OW, _ = dy.shape
    H, W = in_hw
    cin = w_packed_dgrad.shape[1]
    if out is None:
        out = torch.empty(B, H, W, cin, device=dy.device, dtype=torch.bfloat16)
    if stride == 1:
        taps = [(pad - kh, pad - kw, kh * k + kw) for kh in range(k) for kw in range(k)]
        if drop_sum is not None:
            seed, p, colsum = drop_sum
            _nhwc(dy)
            assert colsum.dtype == torch.float32 and colsum.numel() == 2 * cin and out.is_contiguous()
            check(_capi.lib().lun_conv_taps_dropsum_bf16(
                dy.data_ptr(), B, OH, OW, dy.shape[-1], w_packed_dgrad.data_ptr(), w_packed_dgrad.shape[0], cin,
                B, H, W, 1, len(taps), int_array([t[0] for t in taps]), int_array([t[1] for t in taps]),
                int_array([t[2] for t in taps]), out.data_ptr(), H, W, cin, colsum.data_ptr(), seed, float(p),
                _stream()), "lun_conv_taps_dropsum_bf16")
            return out
        return conv_taps(dy, w_packed_dgrad, cin, (B, H, W), 1, taps, out, (H, W))
    assert stride == 2 and H % 2 == 0 and W % 2 == 0
    # input pixel ih = 2*j + ph receives dy[oh] * w[kh] with ih = 2*oh - pad + kh  ->  oh = j + (ph + pad - kh) / 2
    for ph in range(2):
        for pw in range(2):
            taps = []
            for kh in range(k):
                if (ph + pad - kh) % 2:
                    continue
                for kw in range(k):
                    if (pw + pad - kw) % 2:
                        continue
                    taps.append(((ph + pad - kh) // 2, (pw + pad - kw) // 2, kh * k + kw))
            conv_taps(dy, w_packed_dgrad, cin, (B, H // 2, W // 2), 1, taps, out, (H, W), o_mul=2, o_ph=ph, o_pw=pw)
    return out


def conv2d_wgrad(dy, x, k, stride, pad):
    """Gradient of F.conv2d w.r.t. its weight, returned in the reference layout [Cout,Cin,kh,kw] fp32."""
    B, OH, OW, cout = dy.shape
    cin = x.shape[-1]
    dw = torch.zeros(k * k, cout, cin, device=dy.device, dtype=torch.float32)
    wgrad_taps(dy, x, (B, OH, OW), _conv_taps_list(k, pad), dw, in_mul=stride)
    return dw.view(k, k, cout, cin).permute(2, 3, 0, 1)


# ------------------------------------------------------------------------------------------------ conv_transpose2d (k4 s2 p1)
def _convT_phase_taps(ph, pw):
    # output row oh = 2*j + ph gathers input rows ih with oh = 2*ih - 1 + kh
    rows = [(0, 1), (-1, 3)] if ph == 0 else [(1, 0), (0, 2)]   # (ih - j, kh)
    cols = [(0, 1), (-1, 3)] if pw == 0 else [(1, 0), (0, 2)]
    return [(dh, dw, kh * 4 + kw) for dh, kh in rows for dw, kw in cols]


def convT4x4s2_fprop(x, w_packed, bias=None, out=None, img_stats=None):
    """F.conv_transpose2d(k=4, s=2, p=1) as four output-phase 2x2-tap sub-convolutions (lunar_generate.py:169-187)."""
    B, H, W, cin = x.shape
    cout = w_packed.shape[1]
    if out is None:
        out = torch.empty(B, 2 * H, 2 * W, cout, device=x.device, dtype=torch.bfloat16)
    if _HALO_CONVT and out.is_contiguous() and w_packed.shape[0] == 16 and (H * W >= 128 or img_stats is None):
        # one launch for all four phases: the fused-phase halo kernel for the thin stages (csrc/convt_halo_sm100.cu),
        # the tap-list kernel with the phases batched as a work-item dimension otherwise
        _nhwc(x)
        assert w_packed.shape[2] == cin
        assert img_stats is None or img_stats.numel() == 2 * cout * B
        check(_capi.lib().lun_convT4x4s2_bf16(x.data_ptr(), B, H, W, cin, w_packed.data_ptr(), cout, _ptr(bias),
                                              out.data_ptr(), _ptr(img_stats), _stream()), "lun_convT4x4s2_bf16")
        return out
    for ph in range(2):
        for pw in range(2):
            conv_taps(x, w_packed, cout, (B, H, W), 1, _convT_phase_taps(ph, pw), out, (2 * H, 2 * W), o_mul=2,
                      o_ph=ph, o_pw=pw, bias=bias, img_stats=img_stats)
    return out


def convT4x4s2_dgrad(dy, w_packed_dgrad, out=None):
    """Gradient of conv_transpose2d(k4,s2,p1) w.r.t. its input = stride-2 4x4 convolution of dy."""
    B, OH, OW, _ = dy.shape
    cin = w_packed_dgrad.shape[1]
    H, W = OH // 2, OW // 2
    if out is None:
        out = torch.empty(B, H, W, cin, device=dy.device, dtype=torch.bfloat16)
    taps = [(kh - 1, kw - 1, kh * 4 + kw) for kh in range(4) for kw in range(4)]
    return conv_taps(dy, w_packed_dgrad, cin, (B, H, W), 2, taps, out, (H, W))


def convT4x4s2_wgrad(dy, x):
    """Gradient w.r.t. the ConvTranspose2d weight, reference layout [Cin,Cout,4,4] fp32."""
    B, H, W, cin = x.shape
    cout = dy.shape[-1]
    dw = torch.zeros(16, cin, cout, device=dy.device, dtype=torch.float32)
    taps = [(kh - 1, kw - 1, kh * 4 + kw) for kh in range(4) for kw in range(4)]
    wgrad_taps(x, dy, (B, H, W), taps, dw, in_mul=2)
    return dw.view(4, 4, cin, cout).permute(2, 3, 0, 1)


# ------------------------------------------------------------------------------------------------ linear
def linear_fprop(x, w_bf16, bias=None, out_f32=True):
    """nn.Linear on [B,K] bf16 with weight [N,K] bf16 (lunar_generate.py:124-125,165)."""
    B, K = x.shape
    N = w_bf16.shape[0]
    out = torch.empty(B, N, device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    conv_taps(x.view(B, 1, 1, K), w_bf16.view(1, N, K), N, (B, 1, 1), 1, [(0, 0, 0)], out.view(B, 1, 1, N), (1, 1),
              bias=bias, flags=EPI_OUT_F32 if out_f32 else 0)
    return out


def head_linear(x, w_heads, bias=None):
    """`heads` independent Linears in one launch: x [rows, heads*cin] bf16, w_heads [heads, cout, cin] bf16,
    bias [heads*cout] fp32 -> [rows, heads*cout] bf16 (the per-head value projection of the as-executed attention)."""
    rows = x.shape[0]
    heads, cout, cin = w_heads.shape
    assert x.shape[1] == heads * cin and x.is_contiguous() and w_heads.is_contiguous()
    out = torch.empty(rows, heads * cout, device=x.device, dtype=torch.bfloat16)
    check(_capi.lib().lun_head_linear_bf16(x.data_ptr(), rows, heads, cin, w_heads.data_ptr(), cout, _ptr(bias),
                                           out.data_ptr(), _stream()), "lun_head_linear_bf16")
    return out


def linear_dgrad(dy, w_t_bf16):
    """dx = dy @ W with W^T packed as [K,N] bf16; dy [B,N] bf16 -> [B,K] bf16."""
    B, N = dy.shape
    K = w_t_bf16.shape[0]
    out = torch.empty(B, K, device=dy.device, dtype=torch.bfloat16)
    conv_taps(dy.view(B, 1, 1, N), w_t_bf16.view(1, K, N), K, (B, 1, 1), 1, [(0, 0, 0)], out.view(B, 1, 1, K), (1, 1))
    return out


def linear_wgrad(dy, x):
    """dW[N,K] = dy^T @ x, fp32."""
    B, N = dy.shape
    K = x.shape[1]
    dw = torch.zeros(1, N, K, device=dy.device, dtype=torch.float32)
    wgrad_taps(dy.view(B, 1, 1, N), x.view(B, 1, 1, K), (B, 1, 1), [(0, 0, 0)], dw)
    return dw.view(N, K)
