"""Fused Teacher heads (csrc/teacher_heads.cu: lun_heads_fwd / lun_heads_bwd through lunar_evaluator._HeadsFn) against
the oracle's head arithmetic (oracle/restatement.py `_mlp_head` + the mixing of teacher_forward, reference
lunar_evaluator.py:353-397, 417-456) on the same pooled features: the five outputs, the gradient of the pooled expert
sums and all 46 head parameter gradients, dropout ON with the kernel's counter-RNG masks injected into the oracle.
fp32 against fp32: 2e-4 of each tensor's scale. Also: the reference's gradient set (a loss on quality_scores alone
leaves semantic / style / prompt parameters without a gradient) and eval mode."""
import pytest
import torch
import torch.nn.functional as F

import dropout_rng as dr
import teacher_cases as tc
from oracle import restatement as R

P = 0.1


def _oracle_heads(sd, fe_mean, means, masks):
    E = len(means)
    weights = R._mlp_head(fe_mean, sd, "gate", False, masks, "gate_drop").softmax(1)
    quals = [R._mlp_head(means[e], sd, f"quality_heads.{e}", True, masks, f"quality_drop.{e}") for e in range(E)]
    wq = (torch.stack(quals, 1) * weights.unsqueeze(-1)).sum(1)
    comb = (torch.stack(means, 1) * weights.unsqueeze(-1)).sum(1)
    style = R._mlp_head(comb, sd, "style_net", True, masks, "style_drop")
    prompt = R._mlp_head(comb, sd, "prompt_net", True, masks, "prompt_drop")
    sem = torch.sigmoid(R._mlp_head(means[0], sd, "semantic_head", True, masks, "semantic_drop"))
    sem = sem * F.cosine_similarity(prompt, prompt.detach(), dim=1).unsqueeze(1)
    return torch.sigmoid(wq), weights, style, prompt, sem


def _setup(dev, feat, emb, B, dropout):
    from lunaris_orion_b200 import lunar_evaluator as le
    t = tc.make_teacher(dev, feat=feat, emb=emb, dropout=dropout, seed=5)
    with torch.no_grad():                                   # zero biases / unit LayerNorms would hide index mix-ups
        for p in le._head_params(t):
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.1)
    g = torch.Generator().manual_seed(B + feat)
    hw = 128 * 128
    fe = (torch.randn(B, 128, generator=g) * 0.5 * hw)
    pooled = (torch.randn(4, B, feat, generator=g) * hw)
    return t, fe, pooled, hw, g


@pytest.mark.gpu
@pytest.mark.parametrize("feat,emb,B", [(64, 32, 3), (512, 256, 8)])
def test_fused_heads_forward_backward_match_the_oracle_with_dropout_on(cuda_dev, feat, emb, B):
    from lunaris_orion_b200 import _host, lunar_evaluator as le
    t, fe, pooled, hw, g = _setup(cuda_dev, feat, emb, B, P)
    t.train()
    params = le._head_params(t)
    pooled_d = pooled.to(cuda_dev).requires_grad_(True)
    trace = []
    _host.set_dropout_trace(trace)
    try:
        outs = le._HeadsFn.apply(t, 1.0 / hw, True, fe.to(cuda_dev), pooled_d, *params)
    finally:
        _host.set_dropout_trace(None)
    cot = [torch.randn(o.shape, generator=g) for o in outs]
    sum((o * c.to(cuda_dev)).sum() for o, c in zip(outs, cot)).backward()
    assert len(trace) == 8
    masks = dr.oracle_masks(trace, B, 128, 128, feat, P)

    sd = tc.oracle_sd(t)
    means = [(pooled[e] / hw).clone().requires_grad_(True) for e in range(4)]
    ref = _oracle_heads(sd, fe / hw, means, masks)
    sum((o * c).sum() for o, c in zip(ref, cot)).backward()
    names = ("quality_scores", "expert_weights", "style_embedding", "prompt_embedding", "semantic_score")
    for n, a, b in zip(names, outs, ref):
        assert tc.rel_err(a, b) < 2e-4, (n, tc.rel_err(a, b))
    for e in range(4):                                       # d(pooled SUM) = d(mean) / (H W)
        assert tc.rel_err(pooled_d.grad[e] * hw, means[e].grad) < 2e-4, e
    checked = 0
    for name, p in t.named_parameters():
        if not name.startswith(("gate", "quality_heads", "semantic_head", "style_net", "prompt_net")):
            continue
        assert p.grad is not None and sd[name].grad is not None, name
        assert tc.rel_err(p.grad, sd[name].grad) < 2e-4, (name, tc.rel_err(p.grad, sd[name].grad))
        checked += 1
    assert checked == 46


@pytest.mark.gpu
def test_fused_heads_keep_the_reference_gradient_set_and_eval_mode(cuda_dev):
    from lunaris_orion_b200 import lunar_evaluator as le
    t, fe, pooled, hw, _ = _setup(cuda_dev, 64, 32, 4, 0.0)
    t.train()
    pooled_d = pooled.to(cuda_dev).requires_grad_(True)
    outs = le._HeadsFn.apply(t, 1.0 / hw, True, fe.to(cuda_dev), pooled_d, *le._head_params(t))
    (-outs[0].mean() * 0.5).backward()                       # teacher_loss of train_hybrid.py:890-896
    for name, p in t.named_parameters():
        if name.startswith(("gate", "quality_heads")):
            assert p.grad is not None, name
        elif name.startswith(("semantic_head", "style_net", "prompt_net")):
            assert p.grad is None, name                      # reference: these heads never reach the loss
    assert pooled_d.grad is not None and pooled_d.grad.abs().max() > 0
    # eval: no dropout, no saved state, same numbers as the oracle without masks
    t.eval()
    with torch.no_grad():
        ev = le._HeadsFn.apply(t, 1.0 / hw, False, fe.to(cuda_dev), pooled.to(cuda_dev))
        sd = tc.oracle_sd(t)
        ref = _oracle_heads(sd, fe / hw, [pooled[e] / hw for e in range(4)], None)
    for a, b in zip(ev, ref):
        assert tc.rel_err(a, b) < 2e-4
