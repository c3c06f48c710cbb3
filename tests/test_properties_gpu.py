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
@pytest.mark.parametrize("cin,cout,hw,stride", [(64, 128, 32, 2), (256, 256, 16, 1), (128, 64, 32, 1), (64, 32, 64, 1),
                                                (64, 64, 32, 1)])
def test_per_image_statistics_of_conv_and_transposed_conv_match_output(cuda_dev, cin, cout, hw, stride):
    """The GroupNorm statistics the conv epilogue accumulates per image (LUN_EPI_STATS_IMG) equal the per-image channel
    sums / sums of squares of the stored bf16 output - for a strided 3x3 conv and for all four phases of a transposed
    conv - and a grid with fewer than 128 pixels per image is refused (error code, no silent mixing of images)."""
    Bn = 5
    g = torch.Generator(device="cpu").manual_seed(cin + hw)
    x = torch.randn(Bn, hw, hw, cin, generator=g).to(torch.bfloat16).to(cuda_dev)
    w = (torch.randn(cout, cin, 3, 3, generator=g) * 0.05).to(cuda_dev)
    ohw = hw // stride
    st = torch.zeros(Bn, 2, cout, device=cuda_dev)
    y = ops.conv2d_fprop(x, ops.pack_conv_weight(w), 3, stride, 1, img_stats=st).float().view(Bn, ohw * ohw, cout)
    assert torch.allclose(st[:, 0], y.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(st[:, 1], (y * y).sum(1), rtol=1e-4, atol=1e-2)
    wt = (torch.randn(cin, cout, 4, 4, generator=g) * 0.05).to(cuda_dev)
    st = torch.zeros(Bn, 2, cout, device=cuda_dev)
    yt = ops.convT4x4s2_fprop(x, ops.pack_convT_weight(wt), img_stats=st).float().view(Bn, 4 * hw * hw, cout)
    assert torch.allclose(st[:, 0], yt.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(st[:, 1], (yt * yt).sum(1), rtol=1e-4, atol=1e-2)
    ref = torch.nn.functional.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), stride=2,
                                               padding=1).permute(0, 2, 3, 1).reshape(Bn, 4 * hw * hw, cout)
    assert (yt - ref).abs().max().item() < 2e-2 * ref.abs().max().item()
    small = torch.randn(Bn, 8, 8, cin, generator=g).to(torch.bfloat16).to(cuda_dev)
    with pytest.raises(_capi.LunarisB200Error):
        ops.conv2d_fprop(small, ops.pack_conv_weight(w), 3, 1, 1, img_stats=torch.zeros(Bn, 2, cout, device=cuda_dev))


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


@pytest.mark.gpu
def test_weight_pack_kernel_matches_the_permutes_it_replaces(cuda_dev):
    """lun_pack_weight_bf16 (one launch per weight after every optimizer step) reproduces, bit for bit, the four layout
    conversions of ops.py: conv forward / data-gradient and transposed-conv forward / data-gradient operands."""
    g = torch.Generator(device="cpu").manual_seed(9)
    w = torch.randn(96, 40, 3, 3, generator=g).to(cuda_dev)
    wt = torch.randn(40, 96, 4, 4, generator=g).to(cuda_dev)
    bf = lambda t: t.to(torch.bfloat16).contiguous()
    assert torch.equal(ops.pack_conv_weight(w), bf(w.permute(2, 3, 0, 1).reshape(9, 96, 40)))
    assert torch.equal(ops.pack_conv_weight_dgrad(w), bf(w.permute(2, 3, 1, 0).reshape(9, 40, 96)))
    assert torch.equal(ops.pack_convT_weight(wt), bf(wt.permute(2, 3, 1, 0).reshape(16, 96, 40)))
    assert torch.equal(ops.pack_convT_weight_dgrad(wt), bf(wt.permute(2, 3, 0, 1).reshape(16, 40, 96)))
    w1 = torch.randn(512, 128, 1, 1, generator=g).to(cuda_dev)
    assert torch.equal(ops.pack_conv_weight(w1), bf(w1.view(1, 512, 128)))


@pytest.mark.gpu
@pytest.mark.parametrize("cin,hw,tap", [(64, 64, (1, 2)), (64, 64, (3, 0)), (128, 32, (0, 0)), (128, 32, (2, 3))])
def test_transposed_conv_delta_filter_is_an_exact_phase_shift(cuda_dev, cin, hw, tap):
    """ConvTranspose2d(k4,s2,p1) with a filter that is the identity at ONE tap (kh, kw) and zero elsewhere scatters the
    input to one output phase, BIT EXACTLY (single-term fp32 sums of exact bf16 products), zero elsewhere - at the thin
    decoder shapes that run on the halo kernel (shifts read through UMMA descriptor offsets into the halo tile), and
    identically through the four per-phase launches of the tap-list kernel."""
    kh, kw = tap
    cout = cin // 2
    Bn = 3
    g = torch.Generator(device="cpu").manual_seed(hw + kh * 4 + kw)
    x = torch.randn(Bn, hw, hw, cin, generator=g).to(torch.bfloat16).to(cuda_dev)
    w = torch.zeros(cin, cout, 4, 4, device=cuda_dev)
    w[torch.arange(cout), torch.arange(cout), kh, kw] = 1.0          # output channel c copies input channel c
    ref = torch.nn.functional.conv_transpose2d(x.float().permute(0, 3, 1, 2), w, stride=2, padding=1)
    ref = ref.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    wp = ops.pack_convT_weight(w)
    y = ops.convT4x4s2_fprop(x, wp)
    torch.cuda.synchronize()
    assert torch.equal(y, ref)
    saved = ops._HALO_CONVT
    try:
        ops._HALO_CONVT = False                                       # four launches of the generic kernel
        y4 = ops.convT4x4s2_fprop(x, wp)
    finally:
        ops._HALO_CONVT = saved
    assert torch.equal(y4, ref)
    # random filter: the two implementations agree to one bf16 rounding of fp32 sums taken in different orders
    wr = (torch.randn(cin, cout, 4, 4, generator=g) * 0.05).to(cuda_dev)
    wpr = ops.pack_convT_weight(wr)
    a = ops.convT4x4s2_fprop(x, wpr).float()
    try:
        ops._HALO_CONVT = False
        b = ops.convT4x4s2_fprop(x, wpr).float()
    finally:
        ops._HALO_CONVT = saved
    assert (a - b).abs().max().item() <= 2.0 ** -7 * b.abs().max().item()


@pytest.mark.gpu
@pytest.mark.parametrize("P,C", [(1000, 64), (4099, 192), (65536 + 17, 512), (7, 8)])
def test_channel_stats_match_fp64_sums_at_ragged_sizes(cuda_dev, P, C):
    """lun_channel_stats_bf16 (BatchNorm batch statistics of a bf16 [P, C] tensor; reference lunar_evaluator.py:244,251
    through nn.BatchNorm2d): sum and sum of squares per channel against fp64, with P ragged against every block, lane
    and unroll width of the kernel (the four loads of an iteration are predicated one by one)."""
    from lunaris_orion_b200 import _capi
    torch.manual_seed(P + C)
    x = (torch.randn(P, C, device=cuda_dev) * 2 + 0.5).to(torch.bfloat16)
    st = torch.zeros(2 * C, device=cuda_dev)
    _capi.check(_capi.lib().lun_channel_stats_bf16(x.data_ptr(), P, C, st.data_ptr(), _capi.raw_stream()),
                "lun_channel_stats_bf16")
    xd = x.double()
    s1, s2 = xd.sum(0), (xd * xd).sum(0)
    assert (st[:C].double() - s1).abs().max().item() <= 1e-5 * xd.abs().sum(0).max().item() + 1e-4
    assert (st[C:].double() - s2).abs().max().item() <= 1e-5 * s2.max().item() + 1e-4
