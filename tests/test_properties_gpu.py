"""Size-independent properties at the FULL Teacher tile shapes (C = 512, 128x128, CTA-pair + dx-tap-reuse paths), where
an fp32 reference of the whole tensor would be slow: exact identities that any correct implicit GEMM must satisfy."""
import ctypes

import pytest
import torch

from lunaris_orion_b200 import _capi, ops

C, H, B = 512, 128, 8


def _x(dev, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(B, H, H, C, generator=g).to(torch.bfloat16).to(dev)


@pytest.mark.gpu
@pytest.mark.parametrize("tap", [(1, 1), (0, 0), (2, 1), (1, 2)])
def test_delta_filter_is_an_exact_shift(cuda_dev, tap):
    """A 3x3 filter that is the identity at one tap and zero elsewhere must reproduce the input shifted by that tap,
    BIT EXACTLY (bf16 products and fp32 sums of a single term are exact), zero outside the image (TMA halo)."""
    kh, kw = tap
    x = _x(cuda_dev)
    w = torch.zeros(C, C, 3, 3, device=cuda_dev)
    w[torch.arange(C), torch.arange(C), kh, kw] = 1.0
    y = ops.conv2d_fprop(x, ops.pack_conv_weight(w), 3, 1, 1)
    ref = torch.zeros_like(x)
    dy, dx = kh - 1, kw - 1                       # y[h, w] = x[h + dy, w + dx]
    hs, ws = slice(max(0, -dy), H - max(0, dy)), slice(max(0, -dx), H - max(0, dx))
    hd, wd = slice(max(0, dy), H - max(0, -dy)), slice(max(0, dx), H - max(0, -dx))
    ref[:, hs, ws] = x[:, hd, wd]
    torch.cuda.synchronize()
    assert torch.equal(y, ref)
    # the data-gradient of the same filter is the opposite shift
    dxg = ops.conv2d_dgrad(x, ops.pack_conv_weight_dgrad(w), 3, 1, 1, (H, H))
    ref2 = torch.zeros_like(x)
    ref2[:, hd, wd] = x[:, hs, ws]
    assert torch.equal(dxg, ref2)


@pytest.mark.gpu
def test_conv_is_linear_and_stats_match_output(cuda_dev):
    """conv(2x) == 2 conv(x) bit exactly (power-of-two scaling commutes with every rounding), and the fused
    BatchNorm statistics equal the sums of the stored output."""
    x = _x(cuda_dev, 1)
    g = torch.Generator(device="cpu").manual_seed(2)
    w = (torch.randn(C, C, 3, 3, generator=g) * 0.02).to(cuda_dev)
    wp = ops.pack_conv_weight(w)
    st = torch.zeros(2 * C, device=cuda_dev)
    y1 = ops.conv2d_fprop(x, wp, 3, 1, 1, stats=st)
    y2 = ops.conv2d_fprop((x.float() * 2).to(torch.bfloat16), wp, 3, 1, 1)
    torch.cuda.synchronize()
    assert torch.equal((y1.float() * 2).to(torch.bfloat16), y2)
    yf = y1.float().view(-1, C)
    assert torch.allclose(st[:C], yf.sum(0), rtol=1e-4, atol=1e-1)
    assert torch.allclose(st[C:], (yf * yf).sum(0), rtol=1e-4, atol=1e-1)


@pytest.mark.gpu
def test_wgrad_gram_matrix_is_symmetric_with_exact_trace(cuda_dev):
    """1x1 weight gradient with dy == x is the Gram matrix of the activations: symmetric to fp32 reduction noise and
    its trace equals sum(x^2). 3x3: the center-tap slab of dW equals that Gram matrix too."""
    x = _x(cuda_dev, 3)
    dw1 = ops.conv2d_wgrad(x, x, 1, 1, 0).reshape(C, C)
    dw3 = ops.conv2d_wgrad(x, x, 3, 1, 1)
    torch.cuda.synchronize()
    scale = dw1.diag().mean().item()
    assert (dw1 - dw1.t()).abs().max().item() <= 2e-4 * scale
    assert abs(dw1.diag().sum().item() - (x.float() ** 2).sum().item()) <= 1e-4 * dw1.diag().sum().item()
    assert (dw3[:, :, 1, 1] - dw1).abs().max().item() <= 2e-4 * scale
    # opposite taps are transposes of each other: dW[:, :, 0, 0] == dW[:, :, 2, 2]^T
    assert (dw3[:, :, 0, 0] - dw3[:, :, 2, 2].t()).abs().max().item() <= 2e-4 * scale


@pytest.mark.gpu
def test_attention_fold_of_constant_chunks_returns_the_rows(cuda_dev):
    """If the 32 tokens of every chunk are identical the softmax-weighted average is that row, whatever the
    queries are: xbar[b,i,h,:] == affine(y[b, first token of chunk(i)]) for all heads."""
    Bs, N = 2, H * H
    nq, nq_pad = N // 32 + 31, 544
    g = torch.Generator(device="cpu").manual_seed(4)
    rows = torch.randn(Bs, N // 32, 1, C, generator=g).to(torch.bfloat16)
    y = rows.expand(Bs, N // 32, 32, C).reshape(Bs, N, C).contiguous().to(cuda_dev)
    qt = torch.randn(Bs, nq_pad, 8 * C, generator=g).to(torch.bfloat16).to(cuda_dev)
    sc = (torch.rand(C, generator=g) + 0.5).to(cuda_dev)
    sh = torch.randn(C, generator=g).to(cuda_dev)
    xbar = torch.zeros(Bs, nq_pad, 8 * C, device=cuda_dev, dtype=torch.bfloat16)
    _capi.check(_capi.lib().lun_attn_fold_rows_bf16(y.data_ptr(), sc.data_ptr(), sh.data_ptr(), None, qt.data_ptr(),
                                                    xbar.data_ptr(), Bs, N, C, 8, nq_pad, 0, ctypes.c_float(0.0),
                                                    torch.cuda.current_stream().cuda_stream), "fold")
    torch.cuda.synchronize()
    chunk = torch.arange(nq).clamp(max=N // 32 - 1)
    want = (rows[:, chunk, 0].float().to(cuda_dev) * sc + sh)                      # [Bs, nq, C]
    got = xbar[:, :nq].float().view(Bs, nq, 8, C)
    err = (got - want.unsqueeze(2)).abs().max().item()
    assert err <= 2e-2 * want.abs().max().item()
