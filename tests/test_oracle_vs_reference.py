"""Pins oracle/restatement.py against the reference itself (runs only where /root/reference exists)."""

import pytest
import torch

from oracle import reference_loader, restatement as R

pytestmark = pytest.mark.skipif(not reference_loader.available(), reason="reference checkout not present")

FEAT, EMB, LAT, B = 64, 32, 64, 2


def _sd_leaf(module):
    sd = {k: v.detach().clone() for k, v in module.state_dict().items()}
    names = {n for n, _ in module.named_parameters()}
    for k in names:
        sd[k].requires_grad_(True)
    return sd


@pytest.fixture(scope="module")
def ref():
    return reference_loader.load()


def test_attention_scatter_matches_reference(ref):
    _, le = ref
    torch.manual_seed(0)
    att = le.PixelArtAttention(64, dropout=0.0).eval()
    x = torch.randn(2, 64, 128, 128)
    captured = {}
    att.proj.register_forward_pre_hook(lambda m, inp: captured.setdefault("pre", inp[0].detach()))
    with torch.no_grad():
        att(x)
        mine = R.local_attention(att.qkv(x), "reference")
    assert torch.allclose(mine, captured["pre"], atol=1e-5)
    nz = (captured["pre"].abs().sum(1).flatten(1) > 0).sum(1)
    assert int(nz.max()) <= 543


def test_teacher_eval_matches_reference(ref):
    _, le = ref
    torch.manual_seed(1)
    t = le.LunarMoETeacher(feature_dim=FEAT, embedding_dim=EMB).eval()
    x = torch.rand(B, 3, 128, 128) * 2 - 1
    with torch.no_grad():
        out_ref = t(x)
        out = R.teacher_forward(x, _sd_leaf(t), training=False)
    for k in ("quality_scores", "expert_weights", "style_embedding", "prompt_embedding", "semantic_score"):
        assert torch.allclose(out[k], out_ref[k], atol=2e-5, rtol=1e-4), k
    for a, b in zip(out["feature_maps"], out_ref["feature_maps"]):
        assert torch.allclose(a, b, atol=1e-4, rtol=1e-4)


def test_teacher_train_grads_and_bn_counters_match_reference(ref):
    _, le = ref
    torch.manual_seed(2)
    t = le.LunarMoETeacher(feature_dim=FEAT, embedding_dim=EMB, dropout_rate=0.0).train()
    sd = _sd_leaf(t)
    x = torch.rand(B, 3, 128, 128) * 2 - 1
    out_ref = t(x)
    (-out_ref["quality_scores"].mean() * 0.5).backward()
    out = R.teacher_forward(x, sd, training=True)
    (-out["quality_scores"].mean() * 0.5).backward()
    assert torch.allclose(out["quality_scores"], out_ref["quality_scores"], atol=1e-4)
    none_ref = {n for n, p in t.named_parameters() if p.grad is None}
    none_mine = {n for n, _ in t.named_parameters() if sd[n].grad is None}
    assert none_ref == none_mine
    assert len(none_ref) == 168 and len(dict(t.named_parameters())) == 268
    for n, p in t.named_parameters():
        if p.grad is None:
            continue
        scale = p.grad.abs().max().item()
        if n.endswith("shortcut.0.bias"):          # mathematically zero (bias straight into train-mode BN)
            w_scale = dict(t.named_parameters())[n.replace("bias", "weight")].grad.abs().max().item()
            assert (sd[n].grad - p.grad).abs().max().item() <= 1e-3 * w_scale + 1e-7, n
            continue
        assert (sd[n].grad - p.grad).abs().max().item() <= 5e-3 * scale + 1e-7, n
    # BN buffers: running stats and num_batches_tracked (+1 / +2 in a single grad-enabled train pass)
    ref_sd = t.state_dict()
    for k, v in ref_sd.items():
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(sd[k]), k
        elif "running_" in k:
            assert torch.allclose(v, sd[k], atol=1e-4, rtol=1e-3), k


def test_vae_matches_reference(ref):
    lg, _ = ref
    torch.manual_seed(3)
    vae = lg.LunarisCoreVAE(latent_dim=LAT).train()
    sd = _sd_leaf(vae)
    x = torch.rand(B, 3, 128, 128) * 2 - 1
    torch.manual_seed(7)
    recon_ref, mu_ref, lv_ref = vae(x)
    torch.manual_seed(7)
    eps = torch.randn(B, LAT)
    recon, mu, lv = R.vae_forward(x, sd, eps)
    assert torch.allclose(recon, recon_ref, atol=1e-5)
    assert torch.allclose(mu, mu_ref, atol=1e-5) and torch.allclose(lv, lv_ref, atol=1e-5)
    torch.nn.functional.mse_loss(recon_ref, x).backward()
    torch.nn.functional.mse_loss(recon, x).backward()
    for n, p in vae.named_parameters():
        assert torch.allclose(sd[n].grad, p.grad, atol=1e-6 + 1e-4 * p.grad.abs().max().item()), n
    # sampling path (no skips)
    z = torch.randn(3, LAT)
    with torch.no_grad():
        assert torch.allclose(R.decoder_forward(z, [], sd), vae.decoder(z, []), atol=1e-5)


def test_self_attention2d_restatement_matches_reference_class(ref):
    """oracle.restatement.self_attention2d (the formula the GPU parity tests of the flash kernels compare against) ==
    the reference's own SelfAttention2d module (lunar_generate.py:56-78), outputs and all gradients, fp32 CPU."""
    lg, _ = ref
    torch.manual_seed(5)
    att = lg.SelfAttention2d(32)
    with torch.no_grad():
        att.gamma.fill_(0.7)
    sd = _sd_leaf(att)
    x = torch.randn(2, 32, 16, 16)
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y_ref = att(xr)
    y, _ = R.self_attention2d(xo, sd)
    assert torch.allclose(y, y_ref, atol=1e-5)
    g = torch.randn_like(y)
    y_ref.backward(g)
    y.backward(g)
    assert torch.allclose(xo.grad, xr.grad, atol=1e-5)
    for n, p in att.named_parameters():
        assert torch.allclose(sd[n].grad, p.grad, atol=1e-6 + 1e-4 * p.grad.abs().max().item()), n
