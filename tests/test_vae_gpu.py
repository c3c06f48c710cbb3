"""GPU parity of the VAE path (lunaris_orion_b200.lunar_generate) vs the CPU oracle, calibrated against a bf16-autocast
execution of the same op sequence (see test_teacher_gpu.py for the tolerance rationale)."""
import pytest

import vae_cases as vc


@pytest.mark.gpu
def test_vae_forward_backward_and_sampling(cuda_dev):
    rep = vc.vae_report(cuda_dev)
    assert rep["all_grads_present"]                                   # 72/72 tensors, as in the reference
    assert rep["recon"] <= 3 * rep["cal_recon"] + 5e-3
    assert rep["mu"] <= 3 * rep["cal_mu"] + 5e-3
    assert rep["logvar"] <= 3 * rep["cal_mu"] + 5e-3
    assert rep["grad_rel_max"] <= 3 * rep["cal_grad_rel_max"] + 0.01, rep["grad_worst"]
    assert rep["ratio_worst"][0][1] < 4.0, rep["ratio_worst"]
    assert rep["sample"] < 0.05


@pytest.mark.gpu
def test_fused_losses_and_sprite_loader_match_torch(cuda_dev):
    """lun_vae_loss_fwd/bwd vs F.mse_loss + the KL expression of train_hybrid.py:859-862 (fp32: rtol 1e-5), and the
    uint8 loader kernel vs x/127.5 - 1 (bit exact)."""
    import torch
    import torch.nn.functional as F
    from lunaris_orion_b200.lunar_generate import sprites_to_tensor, vae_losses
    g = torch.Generator().manual_seed(1)
    B, L = 4, 64
    recon = torch.tanh(torch.randn(B, 3, 128, 128, generator=g)).to(cuda_dev).requires_grad_(True)
    x = (torch.rand(B, 3, 128, 128, generator=g) * 2 - 1).to(cuda_dev)
    mulv = torch.randn(B, 2 * L, generator=g).to(cuda_dev).requires_grad_(True)
    mu, lv = mulv[:, :L], mulv[:, L:]
    r1, k1 = vae_losses(recon, x, mu, lv)
    (0.7 * r1 + 0.1 * k1).backward()
    g_recon, g_mulv = recon.grad.clone(), mulv.grad.clone()
    recon.grad = mulv.grad = None
    r2 = F.mse_loss(recon, x)
    k2 = -0.5 * torch.mean(1 + lv - mu.pow(2) - lv.exp())
    (0.7 * r2 + 0.1 * k2).backward()
    assert torch.allclose(r1, r2, rtol=1e-5) and torch.allclose(k1, k2, rtol=1e-5)
    assert torch.allclose(g_recon, recon.grad, rtol=1e-5, atol=1e-10)
    assert torch.allclose(g_mulv, mulv.grad, rtol=1e-5, atol=1e-9)
    u8 = torch.randint(0, 256, (3, 128, 128, 3), generator=g, dtype=torch.uint8).to(cuda_dev)
    # the reference normalises on the CPU inside PixelArtDataset (true division); compare with that, bit for bit
    assert torch.equal(sprites_to_tensor(u8).cpu(), u8.cpu().permute(0, 3, 1, 2).float() / 127.5 - 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("C,HW", [(64, 16), (256, 32), (512, 16), (128, 48), (512, 48)])   # 2 / 8 / 2 / 18 / 18 key tiles
def test_self_attention2d_flash_kernel_matches_reference_math(cuda_dev, C, HW):
    """SelfAttention2d.forward (flash-style tcgen05 kernel) vs the reference formula of lunar_generate.py:66-78 in
    fp32 on bf16-rounded operands. Tolerance: 2 % of max |ref| (bf16 q/k/v/P, fp32 accumulation)."""
    import torch
    import torch.nn.functional as F
    from lunaris_orion_b200 import lunar_generate as lg
    torch.manual_seed(C + HW)
    att = lg.SelfAttention2d(C).to(cuda_dev)
    with torch.no_grad():
        att.gamma.fill_(0.7)
        for m in (att.query_conv, att.key_conv):
            m.weight.mul_(0.5)
    x = torch.randn(2, C, HW, HW, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        mine = att(x.to(cuda_dev)).cpu()
        w = {k: v.detach().cpu().float() for k, v in att.state_dict().items()}
        bf = lambda t: t.to(torch.bfloat16).float()
        xb = bf(x)
        B, N = 2, HW * HW
        q = bf(F.conv2d(xb, bf(w["query_conv.weight"]), w["query_conv.bias"])).view(B, -1, N)
        k = bf(F.conv2d(xb, bf(w["key_conv.weight"]), w["key_conv.bias"])).view(B, -1, N)
        v = bf(F.conv2d(xb, bf(w["value_conv.weight"]), w["value_conv.bias"])).view(B, -1, N)
        attn = torch.softmax(torch.bmm(q.permute(0, 2, 1), k), dim=-1)
        out = torch.bmm(v, attn.permute(0, 2, 1)).view(B, C, HW, HW)
        ref = w["gamma"] * out + xb
    err = (mine - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-2, err


@pytest.mark.gpu
@pytest.mark.parametrize("C,HW,B", [(64, 16, 2), (256, 32, 1), (512, 16, 3), (128, 16, 1), (256, 48, 1)])
def test_self_attention2d_backward_matches_autograd_of_reference_math(cuda_dev, C, HW, B):
    """Gradients of SelfAttention2d (flash backward kernels: dV in the forward kernel's key-owner mode, dQ/dK in
    flash_attn2d_bwd_kernel, D/dgamma prep, 1x1-conv dgrad/wgrad) vs torch autograd of the reference formula
    (lunar_generate.py:66-78) in fp32 on the same bf16-rounded parameters and input. Tolerance: 3 % of max |ref| per
    tensor (bf16 q/k/v/P/dS/dY operands, fp32 accumulation)."""
    import torch
    import torch.nn.functional as F
    from lunaris_orion_b200 import lunar_generate as lg
    torch.manual_seed(7 * C + HW)
    att = lg.SelfAttention2d(C).to(cuda_dev)
    with torch.no_grad():
        att.gamma.fill_(0.6)
        for m in (att.query_conv, att.key_conv):
            m.weight.mul_(0.5)
        for p in att.parameters():                       # bf16-representable parameters: both sides see the same values
            p.copy_(p.to(torch.bfloat16).float())
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, C, HW, HW, generator=g).to(torch.bfloat16).float()
    dy = torch.randn(B, C, HW, HW, generator=g).to(torch.bfloat16).float()

    xg = x.to(cuda_dev).requires_grad_(True)
    y = att(xg)
    y.backward(dy.to(cuda_dev))
    mine = {"x": xg.grad.cpu()}
    mine.update({k: p.grad.cpu() for k, p in att.named_parameters()})

    from oracle import restatement as R
    w = {k: v.detach().cpu().float().requires_grad_(True) for k, v in att.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    yr, out = R.self_attention2d(xr, w)              # pinned against the reference class in test_oracle_vs_reference.py
    assert (y.detach().cpu() - yr.detach()).abs().max().item() / yr.abs().max().item() < 2e-2
    yr.backward(dy)
    ref = {"x": xr.grad}
    ref.update({k: t.grad for k, t in w.items()})
    errs = {}
    for name, r in ref.items():
        scale = r.abs().max().item()
        if name == "key_conv.bias":
            # a key bias shifts every score of a row by the same amount, so its true gradient is zero (softmax
            # invariance); compare on the scale of the query-bias gradient instead of its own rounding noise
            scale = ref["query_conv.bias"].abs().max().item()
        if name == "gamma":
            # d gamma = sum(dy * out) is a signed sum of B*N*C terms; measure its error on the scale of that sum's
            # random-walk magnitude when the sum itself happens to cancel
            scale = max(scale, (dy * out.detach()).pow(2).sum().sqrt().item())
        errs[name] = (mine[name] - r).abs().max().item() / (scale + 1e-12)
    assert max(errs.values()) < 3e-2, errs


@pytest.mark.gpu
def test_self_attention2d_gamma_zero_init_gives_identity_and_only_gamma_gradient(cuda_dev):
    """At the reference's init (gamma = 0, lunar_generate.py:67) the module is the identity, every q/k/v gradient is
    exactly zero, dx == dy and d gamma = sum(dy * attention output)."""
    import torch
    from lunaris_orion_b200 import lunar_generate as lg
    torch.manual_seed(0)
    att = lg.SelfAttention2d(64).to(cuda_dev)
    x = torch.randn(1, 64, 16, 16, device=cuda_dev).to(torch.bfloat16).float().requires_grad_(True)
    y = att(x)
    assert torch.equal(y, x.detach())
    dy = torch.randn_like(y).to(torch.bfloat16).float()
    y.backward(dy)
    assert torch.equal(x.grad, dy)
    for name, p in att.named_parameters():
        if name != "gamma":
            assert p.grad is not None and p.grad.abs().max().item() == 0.0, name
    assert att.gamma.grad.abs().item() > 0


@pytest.mark.gpu
def test_generate_and_rank_keeps_only_images_above_the_threshold(cuda_dev):
    """SURVEY 8 f3 scoring path (lunaris_orion_b200/scoring.py; examples/simple_generation.py:71-134): decode ==
    LunarisCoreVAE.sample for the same latents, scores == the Teacher's eval quality mean, the loop returns at most
    num_samples images, all at or above the threshold, best first, and an unreachable threshold returns nothing."""
    import torch
    from lunaris_orion_b200 import lunar_evaluator as le, lunar_generate as lg, scoring
    torch.manual_seed(5)
    vae = lg.LunarisCoreVAE(64).to(cuda_dev).eval()
    teacher = le.LunarMoETeacher(feature_dim=64, embedding_dim=32).to(cuda_dev).eval()
    torch.manual_seed(9)
    a = vae.sample(4)
    torch.manual_seed(9)
    z = torch.randn(4, 64, device=cuda_dev)
    b = scoring.decode(vae, z)
    assert (a - b).abs().max().item() < 2e-2                      # same kernels; GroupNorm sums are atomically ordered
    s = scoring.assess_quality(teacher, b)
    with torch.no_grad():
        ref = teacher(b)["quality_scores"].mean(1, keepdim=True)
    assert s.shape == (4, 1) and (s - ref).abs().max().item() < 2e-2 and not teacher.training
    thr = float(s.median())
    imgs, sc = scoring.generate_and_rank(vae, teacher, num_samples=6, quality_threshold=thr - 0.2, max_attempts=2, seed=3)
    assert 0 < imgs.shape[0] <= 6 and imgs.shape[1:] == (3, 128, 128) and sc.shape == (imgs.shape[0], 1)
    assert bool((sc >= thr - 0.2).all()) and bool((sc[:-1] >= sc[1:]).all())
    none, _ = scoring.generate_and_rank(vae, teacher, num_samples=2, quality_threshold=1.5, max_attempts=1, seed=3)
    assert none.shape[0] == 0


@pytest.mark.gpu
def test_self_attention2d_at_4096_tokens_matches_fp32_attention(cuda_dev):
    """The flash kernels at a size where the N x N matrix is large (N = 4096, C = 256, 32 key tiles, both S buffers and
    both P atoms cycling many times): forward and all gradients vs an fp32 evaluation of lunar_generate.py:66-78 on the
    GPU (TF32 off). Tolerance 3 % of max |ref| per tensor."""
    import torch
    import torch.nn.functional as F
    from lunaris_orion_b200 import lunar_generate as lg
    torch.manual_seed(11)
    C, HW, B = 256, 64, 1
    att = lg.SelfAttention2d(C).to(cuda_dev)
    with torch.no_grad():
        att.gamma.fill_(0.8)
        for m in (att.query_conv, att.key_conv):
            m.weight.mul_(0.5)
        for p in att.parameters():
            p.copy_(p.to(torch.bfloat16).float())
    x = torch.randn(B, C, HW, HW, device=cuda_dev).to(torch.bfloat16).float()
    dy = torch.randn(B, C, HW, HW, device=cuda_dev).to(torch.bfloat16).float()
    xg = x.clone().requires_grad_(True)
    y = att(xg)
    y.backward(dy)
    mine = {"x": xg.grad.clone(), **{k: p.grad.clone() for k, p in att.named_parameters()}}
    from oracle import restatement as R
    w = {k: v.detach().clone().requires_grad_(True) for k, v in att.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    yr, out = R.self_attention2d(xr, w)              # fp32 on the GPU (TF32 off, tests/conftest.py)
    yr.backward(dy)
    assert (y.detach() - yr.detach()).abs().max().item() < 2e-2 * yr.abs().max().item()
    ref = {"x": xr.grad, **{k: t.grad for k, t in w.items()}}
    errs = {}
    for name, r in ref.items():
        scale = r.abs().max().item()
        if name == "key_conv.bias":
            scale = ref["query_conv.bias"].abs().max().item()
        if name == "gamma":
            scale = max(scale, (dy * out.detach()).pow(2).sum().sqrt().item())
        errs[name] = (mine[name] - r).abs().max().item() / (scale + 1e-12)
    assert max(errs.values()) < 3e-2, errs


@pytest.mark.gpu
def test_fused_decoder_tail_equals_the_two_kernels_it_replaces(cuda_dev):
    """lun_gn_mish_final_conv_tanh_fwd (GroupNorm + Mish applied while a halo tile is staged, 32->3 conv on mma.sync,
    tanh) vs lun_gn_mish_fwd_bf16 followed by lun_final_conv_tanh_fwd on the normalised tensor: the saved activation h
    is bit-identical, the images agree to the SFU tanh's 2^-11."""
    import ctypes
    import torch
    from lunaris_orion_b200 import _capi
    lib = _capi.lib()
    B, H = 3, 128
    g = torch.Generator().manual_seed(4)
    t = (torch.randn(B, H * H, 32, generator=g) * 1.5).to(torch.bfloat16).to(cuda_dev)
    gamma = (torch.rand(32, generator=g) + 0.5).to(cuda_dev)
    beta = (torch.randn(32, generator=g) * 0.1).to(cuda_dev)
    w = (torch.randn(3, 32, 3, 3, generator=g) * 0.1).to(cuda_dev)
    bias = (torch.randn(3, generator=g) * 0.1).to(cuda_dev)
    st = torch.zeros(B, 2, 32, device=cuda_dev)
    s = torch.cuda.current_stream().cuda_stream
    P = lambda x: ctypes.c_void_p(x.data_ptr())
    _capi.check(lib.lun_image_channel_stats_bf16(P(t), P(st), B, H * H, 32, s), "stats")
    h_fused = torch.empty_like(t)
    r_fused = torch.empty(B, 3, H, H, device=cuda_dev)
    _capi.check(lib.lun_gn_mish_final_conv_tanh_fwd(P(t), P(st), P(gamma), P(beta), P(w), P(bias), P(h_fused), P(r_fused),
                                                    B, H, H, 8, ctypes.c_float(1e-5), s), "fused tail")
    h_two = torch.empty_like(t)
    r_two = torch.empty_like(r_fused)
    _capi.check(lib.lun_gn_mish_fwd_bf16(P(t), P(st), P(gamma), P(beta), None, None, P(h_two), B, H * H, 32, 8,
                                         ctypes.c_float(1e-5), s), "gn_mish")
    _capi.check(lib.lun_final_conv_tanh_fwd(P(h_two), P(w), P(bias), P(r_two), B, H, H, s), "final conv")
    torch.cuda.synchronize()
    assert torch.equal(h_fused, h_two)
    assert (r_fused - r_two).abs().max().item() < 2e-3
    # and against fp32 torch on the same bf16-rounded activation
    ref = torch.tanh(torch.nn.functional.conv2d(h_two.float().view(B, H, H, 32).permute(0, 3, 1, 2),
                                                w.to(torch.bfloat16).float(), bias, padding=1))
    assert (r_fused - ref).abs().max().item() < 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("C,groups,HW", [(32, 8, 1024), (64, 8, 300), (256, 8, 64)])
def test_gn_mish_forward_matches_torch_fp32_to_a_bf16_ulp(cuda_dev, C, groups, HW):
    """GroupNorm(8) + Mish (lunar_generate.py:37-38, 96-97, 170-189) through lun_image_channel_stats_bf16 +
    lun_gn_mish_fwd_bf16 against torch's fp32 group_norm + mish on the same bf16 input, over a wide input range
    (|x| up to 40: the log2-domain Mish has no clamp - exp overflow must give the linear branch, large negative inputs
    zero). Bound: two bf16 ulps of the fp32 result."""
    import ctypes
    import torch
    import torch.nn.functional as F
    from lunaris_orion_b200 import _capi
    lib = _capi.lib()
    torch.manual_seed(C + HW)
    B = 3
    x = torch.randn(B, HW, C, device=cuda_dev) * 3
    x[0, :7, :5] = 40.0
    x[1, 3:9, 1:4] = -40.0
    x = x.to(torch.bfloat16)
    gamma = torch.rand(C, device=cuda_dev) * 4 + 0.5          # large gains push gn(x) far into both tails
    beta = torch.randn(C, device=cuda_dev)
    st = torch.zeros(B, 2, C, device=cuda_dev)
    s = _capi.raw_stream()
    _capi.check(lib.lun_image_channel_stats_bf16(x.data_ptr(), st.data_ptr(), B, HW, C, s), "stats")
    y = torch.empty_like(x)
    _capi.check(lib.lun_gn_mish_fwd_bf16(x.data_ptr(), st.data_ptr(), gamma.data_ptr(), beta.data_ptr(), None, None,
                                         y.data_ptr(), B, HW, C, groups, ctypes.c_float(1e-5), s), "gn_mish")
    xr = x.float().permute(0, 2, 1).reshape(B, C, HW, 1)
    ref = F.mish(F.group_norm(xr, groups, gamma, beta, 1e-5)).reshape(B, C, HW).permute(0, 2, 1)
    err = (y.float() - ref).abs()
    assert torch.isfinite(y.float()).all()
    assert bool((err <= 2.0 ** -7 * ref.abs() + 2e-6).all()), (err.max().item(), ref.abs().max().item())


@pytest.mark.gpu
@pytest.mark.parametrize("H", [32, 64, 128])
def test_decoder_tail_matches_torch_fp32_restatement(cuda_dev, H):
    """up4's GroupNorm + Mish, final_conv and tanh (lunar_generate.py:187-189, 226-228) in the one fused launch against
    torch: group_norm + mish in fp32, rounded to bf16 (what the kernel stages), conv2d with bf16-rounded weights in fp32,
    bias, bf16 rounding of the pre-activation, tanh. Whole images, so every border / corner tile (zero padding) and
    every interior tile boundary (halo exchange) is covered; the normalised tensor written for the backward must be
    the bf16 rounding of the fp32 result to two ulps."""
    import ctypes
    import torch
    import torch.nn.functional as F
    from lunaris_orion_b200 import _capi
    lib = _capi.lib()
    torch.manual_seed(H)
    B, C = 2, 32
    t = (torch.randn(B, H * H, C, device=cuda_dev) * 1.5).to(torch.bfloat16)
    gamma = torch.rand(C, device=cuda_dev) + 0.5
    beta = torch.randn(C, device=cuda_dev) * 0.1
    w = torch.randn(3, C, 3, 3, device=cuda_dev) * 0.1
    bias = torch.randn(3, device=cuda_dev) * 0.1
    st = torch.zeros(B, 2, C, device=cuda_dev)
    s = _capi.raw_stream()
    _capi.check(lib.lun_image_channel_stats_bf16(t.data_ptr(), st.data_ptr(), B, H * H, C, s), "stats")
    h = torch.empty_like(t)
    recon = torch.empty(B, 3, H, H, device=cuda_dev)
    _capi.check(lib.lun_gn_mish_final_conv_tanh_fwd(t.data_ptr(), st.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                                    w.data_ptr(), bias.data_ptr(), h.data_ptr(), recon.data_ptr(), B, H, H,
                                                    8, ctypes.c_float(1e-5), s), "fused tail")
    x = t.float().view(B, H, H, C).permute(0, 3, 1, 2)
    h_ref = F.mish(F.group_norm(x, 8, gamma, beta, 1e-5))
    h_got = h.float().view(B, H, H, C).permute(0, 3, 1, 2)
    assert bool(((h_got - h_ref).abs() <= 2.0 ** -7 * h_ref.abs() + 2e-6).all())
    pre = F.conv2d(h_got, w.to(torch.bfloat16).float(), bias, padding=1)       # from what the kernel staged: isolates the conv
    ref = torch.tanh(pre.to(torch.bfloat16).float())
    assert (recon - ref).abs().max().item() < 2.0 ** -7, (recon - ref).abs().max().item()
