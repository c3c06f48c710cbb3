"""GPU parity of the Teacher path (lunaris_orion_b200.lunar_evaluator through the C ABI) against the CPU oracle.

Tolerances: the kernels compute in bf16 with fp32 accumulation, like the reference under bf16 autocast. Every check
is therefore calibrated against how far a bf16-autocast execution of the reference's own op sequence (the oracle
run under torch.autocast on the same GPU) drifts from the fp32 oracle: ours must stay within 3x that drift plus a
small floor (stated per assertion)."""
import os

import pytest
import torch

import teacher_cases as tc

gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_small.pt"),
                  weights_only=False)


@pytest.mark.gpu
def test_teacher_eval_matches_oracle(cuda_dev):
    rep = tc.eval_report(cuda_dev)
    assert rep["expert_weights"] < 5e-3            # relative to max |ref|
    assert rep["style_embedding"] < 2e-2 and rep["prompt_embedding"] < 2e-2
    assert rep["feature_maps"] < 3e-2
    assert rep["quality_logit_abs"] < 0.03 * rep["quality_logit_scale"] + 0.02


@pytest.mark.gpu
def test_teacher_trunk_forward_backward_within_bf16_calibration(cuda_dev):
    rep = tc.trunk_report(cuda_dev)
    assert rep["grad_keys_equal"]
    assert rep["pool"] <= 3 * rep["cal_pool"] + 2e-3
    assert rep["grad_rel_max"] <= 3 * rep["cal_grad_rel_max"] + 0.02, rep["grad_worst"]


@pytest.mark.gpu
def test_teacher_train_gradient_set_and_bn_counters(cuda_dev):
    rep = tc.train_report(cuda_dev)
    assert rep["none_set_equal"] and rep["n_none"] == 168
    assert rep["quality_logit_abs"] < 0.15 * rep["quality_logit_scale"]     # ill-conditioned heads at B=2


@pytest.mark.gpu
def test_teacher_eval_matches_reference_golden(cuda_dev):
    """Against outputs of the REAL reference (tests/golden, made by oracle/make_golden.py)."""
    cfg = gold["cfg"]
    from lunaris_orion_b200 import lunar_evaluator as le
    torch.manual_seed(cfg["seed"])
    from lunaris_orion_b200 import lunar_generate as lg
    lg.LunarisCoreVAE(latent_dim=cfg["latent"])            # same RNG consumption order as the golden script
    t = le.LunarMoETeacher(feature_dim=cfg["feat"], embedding_dim=cfg["emb"], dropout_rate=0.0).to(cuda_dev).eval()
    with torch.no_grad():
        out = t(tc.images(cfg["B"], cfg["img_seed"]).to(cuda_dev))
    ref = gold["teacher_eval"]
    assert tc.rel_err(out["expert_weights"], ref["expert_weights"]) < 5e-3
    assert tc.rel_err(out["style_embedding"], ref["style_embedding"]) < 2e-2
    assert (tc.logit(out["quality_scores"]) - tc.logit(ref["quality_scores"])).abs().max().item() < 0.1
    assert len(t.state_dict()) == len(gold["teacher_keys_after_forward"]) == 391


@pytest.mark.gpu
def test_dropout_kernels_statistics_and_replay(cuda_dev):
    """Elementwise dropout: keep-rate ~ 1-p, survivors scaled by bf16(1/(1-p)), and the backward replays the mask."""
    import ctypes
    from lunaris_orion_b200 import _capi
    lib = _capi.lib()
    B, HW, C, nq, nq_pad, p, seed = 2, 4096, 64, 159, 160, 0.1, 12345
    small = torch.ones(B, nq_pad, C, device=cuda_dev, dtype=torch.bfloat16)
    bias = torch.ones(C, device=cuda_dev)
    y = torch.empty(B, HW, C, device=cuda_dev, dtype=torch.bfloat16)
    s = torch.cuda.current_stream().cuda_stream
    _capi.check(lib.lun_proj_expand_bf16(small.data_ptr(), bias.data_ptr(), y.data_ptr(), B, HW, C, nq, nq_pad, seed,
                                         ctypes.c_float(p), s), "expand")
    yf = y.float()
    keep = (yf != 0).float().mean().item()
    assert abs(keep - (1 - p)) < 5e-3
    assert torch.allclose(yf[yf != 0], torch.tensor(1.109375, device=cuda_dev))
    dpo = torch.zeros(B, nq_pad, C, device=cuda_dev, dtype=torch.bfloat16)
    db = torch.zeros(C, device=cuda_dev)
    ones = torch.ones(B, HW, C, device=cuda_dev, dtype=torch.bfloat16)
    _capi.check(lib.lun_proj_bwd_gather_bf16(ones.data_ptr(), dpo.data_ptr(), db.data_ptr(), B, HW, C, nq, nq_pad, seed,
                                             ctypes.c_float(p), s), "gather")
    torch.cuda.synchronize()
    assert torch.equal(dpo[:, :nq].float(), yf[:, :nq])                       # same mask as the forward
    assert torch.allclose(db, yf.sum((0, 1)), rtol=1e-3)


@pytest.mark.gpu
def test_attention_rows_split_equals_full_qkv_and_oracle(cuda_dev):
    """The pruned attention (Q rows gathered, K|V tensor) equals the kernel on the full qkv tensor bit-for-bit and
    matches the oracle's as-executed scatter (local_attention, lunar_evaluator.py:203-216)."""
    import ctypes
    from lunaris_orion_b200 import _capi
    from oracle import restatement as R
    lib = _capi.lib()
    B, H, C, heads = 2, 64, 128, 8
    N = H * H
    nq = N // 32 + 31
    nq_pad = (nq + 7) // 8 * 8
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(B, N, 3 * C, generator=g).to(torch.bfloat16).to(cuda_dev)
    s = torch.cuda.current_stream().cuda_stream
    full = torch.zeros(B, nq_pad, C, device=cuda_dev, dtype=torch.bfloat16)
    _capi.check(lib.lun_attn_ref_rows_bf16(qkv.data_ptr(), full.data_ptr(), B, N, C, heads, nq_pad, 0,
                                           ctypes.c_float(0.0), s), "attn")
    qg = torch.zeros(B, nq_pad, 3 * C, device=cuda_dev, dtype=torch.bfloat16)
    _capi.check(lib.lun_gather_query_rows_bf16(qkv.data_ptr(), qg.data_ptr(), B, N, 3 * C, nq_pad, s), "gather")
    q_small = qg[:, :, :C].contiguous()
    kv = qkv[:, :, C:].contiguous()
    split = torch.zeros(B, nq_pad, C, device=cuda_dev, dtype=torch.bfloat16)
    _capi.check(lib.lun_attn_ref_rows_split_bf16(q_small.data_ptr(), kv.data_ptr(), split.data_ptr(), B, N, C, heads,
                                                 nq_pad, 0, ctypes.c_float(0.0), s), "attn split")
    torch.cuda.synchronize()
    assert torch.equal(full, split)
    ref = R.local_attention(qkv.float().cpu().view(B, H, H, 3 * C).permute(0, 3, 1, 2), "reference")
    ref = ref.permute(0, 2, 3, 1).reshape(B, N, C)
    assert ref[:, nq:].abs().max().item() == 0.0                       # everything past row nq stays zero
    assert tc.rel_err(full[:, :nq], ref[:, :nq]) < 2e-2


@pytest.mark.gpu
def test_folded_attention_module_matches_oracle(cuda_dev):
    """PixelArtAttention.forward (K/V-free folded path, csrc/attn_fold.cu) vs the oracle's qkv conv + as-executed
    local attention + proj (lunar_evaluator.py:146-227), eval mode. Tolerance 2 % of max |ref| (bf16 path)."""
    import torch.nn.functional as F
    from lunaris_orion_b200 import lunar_evaluator as le
    from oracle import restatement as R
    torch.manual_seed(4)
    att = le.PixelArtAttention(128, dropout=0.0).to(cuda_dev).eval()
    with torch.no_grad():
        att.qkv.bias.normal_(0, 0.2)
        att.proj.bias.normal_(0, 0.2)
    x = torch.randn(2, 128, 64, 64, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        mine = att(x.to(cuda_dev)).cpu()
        w = {k: v.detach().cpu().float() for k, v in att.state_dict().items() if v is not None}
        qkv = F.conv2d(x, w["qkv.weight"], w["qkv.bias"])
        ref = F.conv2d(R.local_attention(qkv, "reference"), w["proj.weight"], w["proj.bias"])
    assert tc.rel_err(mine, ref) < 2e-2
    nq = 64 * 64 // 32 + 31
    flat = mine.permute(0, 2, 3, 1).reshape(2, -1, 128)
    assert torch.allclose(flat[:, nq:], w["proj.bias"].to(torch.bfloat16).float().expand_as(flat[:, nq:]))


@pytest.mark.gpu
def test_teacher_trunk_feat256_config2_shapes(cuda_dev):
    """Same trunk parity at feature_dim 256 (BASELINE configs[1], head_dim 32, single 256-wide N block)."""
    rep = tc.trunk_report(cuda_dev, B=2, feat=256, calibrate=True)
    assert rep["grad_keys_equal"]
    assert rep["pool"] <= 3 * rep["cal_pool"] + 2e-3
    assert rep["grad_rel_max"] <= 3 * rep["cal_grad_rel_max"] + 0.02, rep["grad_worst"]


@pytest.mark.gpu
@pytest.mark.parametrize("C,hw,p", [(64, 128, 0.1), (512, 128, 0.1), (128, 128, 0.0)])
def test_conv_epilogue_replays_dropout_for_the_bias_gradient(cuda_dev, C, hw, p):
    """lun_conv_taps_dropsum_bf16: the 3x3 data-gradient conv whose epilogue also sums mask * bf16(dx / (1-p)) per
    channel (proj_drop backward + proj bias gradient, lunar_evaluator.py:224-225) equals the two-step path it replaces
    (plain conv, then lun_proj_bwd_gather_bf16 over the whole tensor): same dx bit for bit, same column sums to fp32
    summation order, same mask as the forward lun_proj_expand_bf16 (seeded counter RNG)."""
    import ctypes
    from lunaris_orion_b200 import _capi, ops
    lib = _capi.lib()
    B, seed = 2, 987654321
    nq = hw * hw // 32 + 31
    nq_pad = (nq + 7) // 8 * 8
    g = torch.Generator().manual_seed(C + hw)
    dz = torch.randn(B, hw, hw, C, generator=g).to(torch.bfloat16).to(cuda_dev)
    w = (torch.randn(C, C, 3, 3, generator=g) * 0.05).to(cuda_dev)
    wd = ops.pack_conv_weight_dgrad(w)
    colsum = torch.zeros(2 * C, device=cuda_dev)
    assert ops.drop_sum_ok(B, hw, hw, C) and not ops.drop_sum_ok(B, 32, 32, C)
    with pytest.raises(_capi.LunarisB200Error):          # narrower rows are refused, never silently mis-indexed
        ops.conv2d_dgrad(dz[:, :32, :32].contiguous(), wd, 3, 1, 1, (32, 32), drop_sum=(seed, p, colsum))
    dx_fused = ops.conv2d_dgrad(dz, wd, 3, 1, 1, (hw, hw), drop_sum=(seed, p, colsum))
    dx_plain = ops.conv2d_dgrad(dz, wd, 3, 1, 1, (hw, hw))
    assert torch.equal(dx_fused, dx_plain)
    dpo = torch.zeros(B, nq_pad, C, device=cuda_dev, dtype=torch.bfloat16)
    db = torch.zeros(C, device=cuda_dev)
    s = torch.cuda.current_stream().cuda_stream
    _capi.check(lib.lun_proj_bwd_gather_bf16(dx_plain.data_ptr(), dpo.data_ptr(), db.data_ptr(), B, hw * hw, C, nq,
                                             nq_pad, seed, ctypes.c_float(p), s), "gather")
    dpo2 = torch.zeros_like(dpo)
    _capi.check(lib.lun_proj_bwd_gather_bf16(dx_plain.data_ptr(), dpo2.data_ptr(), None, B, hw * hw, C, nq, nq_pad, seed,
                                             ctypes.c_float(p), s), "gather rows only")
    torch.cuda.synchronize()
    assert torch.equal(dpo, dpo2)
    scale = db.abs().max().item() + 1e-6
    assert (colsum[:C] - db).abs().max().item() < 2e-3 * scale, ((colsum[:C] - db).abs().max().item(), scale)
    if p == 0.0:
        assert (colsum[:C] - dx_plain.float().sum((0, 1, 2))).abs().max().item() < 2e-3 * scale
