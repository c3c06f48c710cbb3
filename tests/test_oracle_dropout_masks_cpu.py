"""Pins the oracle's mask-injection plumbing (oracle/restatement.py `masks=`) against the UNMODIFIED reference with
dropout ON: every nn.Dropout / nn.Dropout2d of the reference Teacher (lunar_evaluator.py:97,212,225,245,252,359,371)
draws an explicit keep-mask that is recorded per module, the same masks are injected into the oracle, and outputs,
the executed gradient set and all gradients must agree. Build container only (needs /root/reference)."""
import pytest
import torch
import torch.nn as nn

from oracle import reference_loader
from oracle import restatement as R

pytestmark = pytest.mark.skipif(not reference_loader.available(), reason="reference not present (GPU box)")


def _images(B, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (B, 3, 128, 128), generator=g, dtype=torch.uint8).float() / 127.5 - 1.0


def test_oracle_mask_sites_match_reference_dropout_sites():
    _, le = reference_loader.load()
    torch.manual_seed(3)
    teacher = le.LunarMoETeacher(feature_dim=64, embedding_dim=32)          # dropout_rate 0.1 (reference default)
    with torch.no_grad():                                                   # zero biases would hide mask mismatches
        for n, p in teacher.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.05)
    sd = {k: v.detach().clone() for k, v in teacher.state_dict().items()}
    for n, _ in teacher.named_parameters():
        sd[n].requires_grad_(True)
    names = {id(m): n for n, m in teacher.named_modules()}
    rec, recording = {}, [True]

    def make(feature):
        def forward(self, x):
            if not self.training or self.p == 0:
                return x
            shape = (x.shape[0], x.shape[1]) + (1,) * (x.dim() - 2) if feature else x.shape
            keep = torch.empty(shape).bernoulli_(1 - self.p) / (1 - self.p)
            if recording[0]:                       # the checkpoint recompute replays the RNG: same masks, not recorded
                rec.setdefault(names[id(self)], []).append(keep)
            return x * keep
        return forward
    saved = nn.Dropout.forward, nn.Dropout2d.forward
    nn.Dropout.forward, nn.Dropout2d.forward = make(False), make(True)
    try:
        teacher.train()
        x = _images(2, 6)
        torch.manual_seed(17)
        out = teacher(x)
        recording[0] = False
        (-out["quality_scores"].mean() * 0.5).backward()
    finally:
        nn.Dropout.forward, nn.Dropout2d.forward = saved

    masks = {"fe_drop": rec["feature_extractor.dropout"][0], "gate_drop": rec["gate.4"][0],
             "style_drop": rec["style_net.5"][0], "prompt_drop": rec["prompt_net.5"][0],
             "semantic_drop": rec["semantic_head.5"][0]}
    for e in range(4):
        masks[f"quality_drop.{e}"] = rec[f"quality_heads.{e}.5"][0]
        for b in range(3):
            p = f"experts.{e}.{b}"
            masks[p + ".drop2d_1"] = rec[p + ".conv1.3"][0]
            masks[p + ".drop2d_2"] = rec[p + ".conv2.3"][0]
            masks[p + ".proj_drop"] = rec[p + ".attention.proj_drop"][0]
            chunks = rec[p + ".attention.attn_drop"]
            assert len(chunks) == 512                                      # one nn.Dropout call per 32-token chunk
            masks[p + ".attn_drop"] = torch.stack(chunks, 2)               # [B,h,nc,32,32]
    assert len(masks) == 1 + 12 * 4 + 8

    ref = R.teacher_forward(x, sd, training=True, masks=masks)
    (-ref["quality_scores"].mean() * 0.5).backward()
    for k in ("quality_scores", "expert_weights", "style_embedding", "prompt_embedding", "semantic_score"):
        d = (out[k].detach() - ref[k].detach()).abs().max().item()
        assert d <= 2e-4 * (1 + ref[k].detach().abs().max().item()), (k, d)
    none_ref = {n for n, p in teacher.named_parameters() if p.grad is None}
    none_oracle = {n for n, _ in teacher.named_parameters() if sd[n].grad is None}
    assert none_ref == none_oracle and len(none_ref) == 168
    worst = 0.0
    for n, p in teacher.named_parameters():
        if p.grad is None:
            continue
        if n.endswith("shortcut.0.bias"):      # exact gradient is zero (bias feeding a train-mode BN): noise only
            scale = dict(teacher.named_parameters())[n.replace("bias", "weight")].grad.abs().max().item()
        else:
            scale = p.grad.abs().max().item() + 1e-12
        worst = max(worst, (p.grad - sd[n].grad).abs().max().item() / scale)
    assert worst < 5e-3, worst
