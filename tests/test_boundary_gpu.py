"""Drop-in boundary behaviour on the GPU: `ResBlock.forward` as a standalone module including the
in != out 1x1 shortcut (reference lunar_generate.py:28-53), and the autograd guards - entry points that cannot deliver a
gradient refuse inputs that require one instead of silently cutting the graph."""
import pytest
import torch
import torch.nn.functional as F

import teacher_cases as tc
from oracle import restatement as R


def _ref_resblock(x, sd):
    """lunar_generate.py:47-53 with the optional 1x1 shortcut conv."""
    idn = F.conv2d(x, sd["shortcut.weight"], sd["shortcut.bias"]) if "shortcut.weight" in sd else x
    h = R._gn_mish(F.conv2d(x, sd["conv1.0.weight"], sd["conv1.0.bias"], padding=1), sd, "conv1.1")
    h = R._gn_mish(F.conv2d(h, sd["conv2.0.weight"], sd["conv2.0.bias"], padding=1), sd, "conv2.1")
    return R._mish(h + idn)


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout,hw", [(64, 64, 32), (64, 128, 32), (128, 256, 16)])
def test_resblock_forward_matches_reference_formula(cuda_dev, cin, cout, hw):
    from lunaris_orion_b200 import lunar_generate as lg
    torch.manual_seed(cin + cout)
    rb = lg.ResBlock(cin, cout).to(cuda_dev)
    x = torch.randn(2, cin, hw, hw, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        mine = rb(x.to(cuda_dev)).cpu()
        sd = {k: v.detach().cpu().float() for k, v in rb.state_dict().items()}
        ref = _ref_resblock(x, sd)
    assert mine.shape == ref.shape == (2, cout, hw, hw)
    assert tc.rel_err(mine, ref) < 3e-2                     # bf16 operands, three rounded intermediates


@pytest.mark.gpu
def test_inference_only_entry_points_refuse_inputs_that_require_grad(cuda_dev):
    from lunaris_orion_b200 import _capi, lunar_evaluator as le, lunar_generate as lg
    vae = lg.LunarisCoreVAE(64).to(cuda_dev)
    teacher = le.LunarMoETeacher(feature_dim=64, embedding_dim=32).to(cuda_dev)
    x = (torch.rand(2, 3, 128, 128, device=cuda_dev) * 2 - 1).requires_grad_(True)
    for call in (lambda: vae.encoder(x), lambda: teacher(x), lambda: teacher.feature_extractor(x),
                 lambda: vae.decoder(torch.randn(2, 64, device=cuda_dev, requires_grad=True), []),
                 lambda: vae.encoder.down2[3](torch.randn(2, 128, 32, 32, device=cuda_dev, requires_grad=True))):
        with pytest.raises(_capi.LunarisB200Error):
            call()
    with torch.no_grad():                                   # the same calls are fine without autograd
        mu, lv, skips = vae.encoder(x)
        assert mu.shape == (2, 64) and len(skips) == 3
        assert vae.decoder(mu, skips).shape == (2, 3, 128, 128)
        assert teacher(x)["quality_scores"].shape == (2, 4)
    # the differentiable entry points still differentiate: VAE w.r.t. all 72 tensors, Teacher (train) w.r.t. its live set
    recon, mu, lv = vae(x.detach())
    (recon.mean() + mu.mean() + lv.mean()).backward()
    assert all(p.grad is not None for p in vae.parameters())
    teacher.train()
    (-teacher(x.detach())["quality_scores"].mean()).backward()
    assert sum(p.grad is not None for p in teacher.parameters()) == 100
    # eval mode under autograd: no graph (warned once), never a partial one
    teacher.eval()
    le._warned.clear()
    with pytest.warns(UserWarning):
        out = teacher(x.detach())
    assert all(not v.requires_grad for v in out.values() if torch.is_tensor(v))
    # use_checkpointing=False changes the reference's gradient set: refused in training
    t2 = le.LunarMoETeacher(feature_dim=64, embedding_dim=32, use_checkpointing=False).to(cuda_dev).train()
    with pytest.raises(_capi.LunarisB200Error):
        t2(x.detach())
    with torch.no_grad():
        assert t2(x.detach())["quality_scores"].shape == (2, 4)
