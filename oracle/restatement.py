"""CPU oracle: a loop-free fp32 restatement of the reference's hybrid training step AS EXECUTED.

TEST INFRASTRUCTURE ONLY. Nothing in the product path (lunaris_orion_b200/) may import this module; it is used by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker.

Parity status: PINNED by execution. The reference ships no tests or golden vectors, so this
restatement is pinned against the reference itself: tests/test_oracle_vs_reference.py imports /root/reference in the
build container and compares module outputs, gradients, BN counters and the 12 step metrics; oracle/make_golden.py
writes the reference's outputs to tests/golden/ so the pin travels to the GPU box where /root/reference is absent.

Functions take a *state_dict* (reference key names) and plain tensors, so they restate arithmetic only.
Reference lines restated (MeryylleA/Lunaris-Orion):
  * PixelArtFeatureExtractor.forward      lunar_evaluator.py:105-112
  * PixelArtAttention.forward             lunar_evaluator.py:146-227  (chunk-index scatter of :203-216)
  * ExpertBlock._forward_path / forward   lunar_evaluator.py:260-275
  * LunarMoETeacher.forward               lunar_evaluator.py:409-462
  * reentrant-checkpoint gradient cuts    lunar_evaluator.py:195,271,412 (three detach rules)
  * ResBlock / Encoder / Decoder / VAE    lunar_generate.py:28-53,127-153,194-229,248-276
  * step losses                           train_hybrid.py:845-896
"""
import math

import torch
import torch.nn.functional as F

CHUNK = 32      # min(self.chunk_size, 32)  lunar_evaluator.py:148
HEADS = 8       # PixelArtAttention num_heads default, lunar_evaluator.py:126
BN_EPS = 1e-5
BN_MOM = 0.1


# ------------------------------------------------------------------------------------------------ helpers
def _bn(x, sd, prefix, training, bn_updates=1):
    """nn.BatchNorm2d. In training mode the running statistics in `sd` are updated `bn_updates` times with the same
    batch statistics (the reference's pass-B recompute re-runs the block: SURVEY.md §0.4)."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if not training:
        return F.batch_norm(x, rm, rv, w, b, False, 0.0, BN_EPS)
    y = F.batch_norm(x, None, None, w, b, True, 0.0, BN_EPS)
    with torch.no_grad():
        n = x.numel() // x.shape[1]
        mean = x.mean((0, 2, 3))
        var_u = x.var((0, 2, 3), unbiased=True) if n > 1 else torch.zeros_like(mean)
        for _ in range(bn_updates):
            rm.mul_(1 - BN_MOM).add_(BN_MOM * mean)
            rv.mul_(1 - BN_MOM).add_(BN_MOM * var_u)
            sd[prefix + ".num_batches_tracked"] += 1
    return y


def _mask(masks, key, x, p):
    """Dropout with an injected keep-mask (already scaled by 1/(1-p)); masks=None means p == 0 or eval."""
    if masks is None or key not in masks:
        return x
    return x * masks[key]


def _conv(x, sd, prefix, **kw):
    return F.conv2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), **kw)


def _mlp_head(pooled, sd, prefix, ln, masks, key):
    """[LayerNorm] -> Linear -> LeakyReLU(0.2) -> Dropout -> Linear on pooled [B,C] features
    (lunar_evaluator.py:353-397; indices follow the nn.Sequential: LN=2, Linear=3, Linear=6 / gate: 2, 5)."""
    if ln:
        h = F.layer_norm(pooled, pooled.shape[-1:], sd[prefix + ".2.weight"], sd[prefix + ".2.bias"])
        l1, l2 = prefix + ".3", prefix + ".6"
    else:
        h = pooled
        l1, l2 = prefix + ".2", prefix + ".5"
    h = F.leaky_relu(F.linear(h, sd[l1 + ".weight"], sd[l1 + ".bias"]), 0.2)
    h = _mask(masks, key, h, None)
    return F.linear(h, sd[l2 + ".weight"], sd[l2 + ".bias"])


# ------------------------------------------------------------------------------------------------ Teacher
def feature_extractor(x, sd, training, masks=None, bn_updates=1, p="feature_extractor"):
    """lunar_evaluator.py:105-112."""
    def cbn(h, conv, bn, **kw):
        return _bn(F.leaky_relu(_conv(h, sd, conv, **kw), 0.2), sd, bn, training, bn_updates)
    h = cbn(x, p + ".conv1.0", p + ".conv1.2", padding=1)
    br = []
    for name, pad in (("edge_branch", 1), ("color_branch", 2), ("detail_branch", 1)):
        d = _conv(h, sd, f"{p}.{name}.0", padding=pad, groups=32)
        br.append(cbn(d, f"{p}.{name}.1", f"{p}.{name}.3"))
    cat = _mask(masks, "fe_drop", torch.cat(br, 1), None)
    return cbn(cat, p + ".fusion.0", p + ".fusion.2")


def local_attention(qkv, semantics="reference", attn_mask=None):
    """Block-local attention over 32-token chunks of the raster-ordered tokens (lunar_evaluator.py:146-220).
    qkv: [B, 3C, H, W] (channel order [3][heads][hd], :154). Returns the pre-proj tensor [B, C, H, W].
    semantics='reference' reproduces the chunk-INDEX scatter of :216: position p < nc holds row 0 of
    chunk p, positions nc-1 .. nc+30 hold chunk nc-1 entirely, everything else stays zero.
    semantics='intended' writes chunk c to tokens 32c..32c+31.
    The rel-pos term (:209-210) is constant along the softmax axis and is dropped (no effect on outputs)."""
    B, C3, H, W = qkv.shape
    C = C3 // 3
    hd = C // HEADS
    N = H * W
    nc = N // CHUNK
    t = qkv.reshape(B, 3, HEADS, hd, N).permute(0, 1, 2, 4, 3)              # [B,3,h,N,hd]
    q, k, v = (t[:, i].reshape(B, HEADS, nc, CHUNK, hd) for i in range(3))
    a = ((q @ k.transpose(-1, -2)) * hd ** -0.5).softmax(-1)                 # [B,h,nc,32,32]
    if attn_mask is not None:
        a = a * attn_mask
    o = a @ v                                                                # [B,h,nc,32,hd]
    if semantics == "intended":
        out = o.reshape(B, HEADS, N, hd)
    else:
        out = torch.zeros(B, HEADS, N, hd, dtype=o.dtype, device=qkv.device)
        out[:, :, :nc] = o[:, :, :, 0]
        out[:, :, nc - 1: nc - 1 + CHUNK] = o[:, :, nc - 1]
    return out.permute(0, 1, 3, 2).reshape(B, C, H, W)                       # :223


def expert_block(x, sd, p, training, semantics, masks, mkey, bn_updates, grad_mode):
    """lunar_evaluator.py:260-275. grad_mode='reference' applies detach rules (ii) and (iii) of SURVEY App. A.5."""
    has_sc = (p + ".shortcut.0.weight") in sd
    if has_sc:
        identity = _bn(_conv(x, sd, p + ".shortcut.0"), sd, p + ".shortcut.1", training, bn_updates)
    else:
        identity = x
    cut_main = grad_mode == "reference" and training and not x.requires_grad     # rule (ii)
    xin = x
    h = F.leaky_relu(_conv(xin, sd, p + ".conv1.0", padding=1), 0.2)
    h = _bn(h, sd, p + ".conv1.2", training, bn_updates)
    h = _mask(masks, mkey + ".drop2d_1", h, None)
    qkv = _conv(h, sd, p + ".attention.qkv")
    att = local_attention(qkv, semantics, None if masks is None else masks.get(mkey + ".attn_drop"))
    if grad_mode == "reference" and training:
        att = att.detach()                                                       # rule (iii)
    h = _conv(att, sd, p + ".attention.proj")
    h = _mask(masks, mkey + ".proj_drop", h, None)
    h = F.leaky_relu(_conv(h, sd, p + ".conv2.0", padding=1), 0.2)
    h = _bn(h, sd, p + ".conv2.2", training, bn_updates)
    h = _mask(masks, mkey + ".drop2d_2", h, None)
    h = h * sd[p + ".layer_scale"]
    if cut_main:
        h = h.detach()
    return F.leaky_relu(h + identity, 0.2)


def teacher_trunk(x, sd, training=True, semantics="reference", grad_mode="reference", masks=None,
                  no_grad_pass=False, num_experts=4, expert_layers=3):
    """Feature extractor + the expert stacks (lunar_evaluator.py:411-424): returns (features, [expert outputs])."""
    fe = feature_extractor(x, sd, training, masks, 1)
    if grad_mode == "reference" and training:
        fe = fe.detach()                                                         # rule (i)
    outs = []
    for e in range(num_experts):
        h = fe
        for b in range(expert_layers):
            p = f"experts.{e}.{b}"
            recomputed = training and not no_grad_pass and grad_mode == "reference" and h.requires_grad
            h = expert_block(h, sd, p, training, semantics, masks, p, 2 if recomputed else 1, grad_mode)
        outs.append(h)
    return fe, outs


def teacher_forward(x, sd, training=True, semantics="reference", grad_mode="reference", masks=None,
                    no_grad_pass=False, num_experts=4, expert_layers=3):
    """LunarMoETeacher.forward (lunar_evaluator.py:409-462) on state_dict `sd` (BN buffers updated in place).

    training=True, no_grad_pass=True : the pass-A call under torch.no_grad (train_hybrid.py:853-855): every BN
                                        running stat is updated once.
    training=True, no_grad_pass=False: pass B; BNs of blocks whose checkpoint segment is re-run in backward
                                        (blocks 1,2 in reference grad mode) are updated twice.
    Returns the reference's output dict plus 'quality_logits' / 'semantic_logit' (pre-sigmoid, for tolerances)."""
    fe, outs = teacher_trunk(x, sd, training, semantics, grad_mode, masks, no_grad_pass, num_experts, expert_layers)
    pooled_fe = fe.mean((2, 3))
    gate_logits = _mlp_head(pooled_fe, sd, "gate", False, masks, "gate_drop")
    weights = gate_logits.softmax(1)
    quals = [_mlp_head(outs[e].mean((2, 3)), sd, f"quality_heads.{e}", True, masks, f"quality_drop.{e}")
             for e in range(num_experts)]
    qt = torch.stack(quals, 1)
    wq = (qt * weights.unsqueeze(-1)).sum(1)
    comb = (torch.stack([o.mean((2, 3)) for o in outs], 1) * weights.unsqueeze(-1)).sum(1)
    style = _mlp_head(comb, sd, "style_net", True, masks, "style_drop")
    prompt = _mlp_head(comb, sd, "prompt_net", True, masks, "prompt_drop")
    sem_logit = _mlp_head(outs[0].mean((2, 3)), sd, "semantic_head", True, masks, "semantic_drop")
    sem = torch.sigmoid(sem_logit) * F.cosine_similarity(prompt, prompt.detach(), dim=1).unsqueeze(1)
    return {
        "quality_scores": torch.sigmoid(wq), "expert_weights": weights, "style_embedding": style,
        "prompt_embedding": prompt, "semantic_score": sem,
        "feature_maps": None if training else outs,
        "quality_logits": wq, "semantic_logit": sem_logit,
    }


# ------------------------------------------------------------------------------------------------ VAE
def _mish(x):
    return x * torch.tanh(F.softplus(x))


def _gn_mish(x, sd, p):
    return _mish(F.group_norm(x, 8, sd[p + ".weight"], sd[p + ".bias"], 1e-5))


def _resblock(x, sd, p):
    """lunar_generate.py:47-53 (shortcut is Identity wherever the model uses it)."""
    h = _gn_mish(_conv(x, sd, p + ".conv1.0", padding=1), sd, p + ".conv1.1")
    h = _gn_mish(_conv(h, sd, p + ".conv2.0", padding=1), sd, p + ".conv2.1")
    return _mish(h + x)


def encoder_forward(x, sd, p="encoder"):
    """lunar_generate.py:127-153."""
    skips = []
    h = x
    for i in range(1, 5):
        q = f"{p}.down{i}"
        h = _gn_mish(_conv(h, sd, q + ".0", stride=2, padding=1), sd, q + ".1")
        h = _resblock(h, sd, q + ".3")
        if i < 4:
            skips.append(h)
    flat = h.flatten(1)
    mu = F.linear(flat, sd[p + ".fc_mu.weight"], sd[p + ".fc_mu.bias"])
    logvar = F.linear(flat, sd[p + ".fc_logvar.weight"], sd[p + ".fc_logvar.bias"])
    return mu, logvar, skips


def decoder_forward(z, skips, sd, p="decoder"):
    """lunar_generate.py:194-229 (skips may be [] for sampling, :212-222)."""
    h = F.linear(z, sd[p + ".fc.weight"], sd[p + ".fc.bias"]).view(z.shape[0], 512, 8, 8)
    for i in range(1, 5):
        q = f"{p}.up{i}"
        h = F.conv_transpose2d(h, sd[q + ".0.weight"], sd[q + ".0.bias"], stride=2, padding=1)
        h = _gn_mish(h, sd, q + ".1")
        if i <= 3 and len(skips) >= 4 - i:
            h = h + skips[3 - i]
    return torch.tanh(_conv(h, sd, p + ".final_conv", padding=1))


def self_attention2d(x, sd, p=""):
    """SelfAttention2d.forward (lunar_generate.py:66-78): q, k = 1x1 convs to C/8, v = 1x1 conv to C; attention =
    softmax_j(q_i . k_j) over all N = H*W positions (no scaling); out = v attention^T; gamma * out + x.
    sd keys: query_conv / key_conv / value_conv .weight/.bias, gamma."""
    B, C, H, W = x.shape
    N = H * W
    q = F.conv2d(x, sd[p + "query_conv.weight"], sd[p + "query_conv.bias"]).view(B, -1, N)
    k = F.conv2d(x, sd[p + "key_conv.weight"], sd[p + "key_conv.bias"]).view(B, -1, N)
    v = F.conv2d(x, sd[p + "value_conv.weight"], sd[p + "value_conv.bias"]).view(B, -1, N)
    attn = torch.softmax(torch.bmm(q.permute(0, 2, 1), k), dim=-1)
    out = torch.bmm(v, attn.permute(0, 2, 1)).view(B, C, H, W)
    return sd[p + "gamma"] * out + x, out


def vae_forward(x, sd, eps):
    """LunarisCoreVAE.forward with the reparameterisation noise supplied (lunar_generate.py:259-276)."""
    mu, logvar, skips = encoder_forward(x, sd)
    z = mu + eps * torch.exp(0.5 * logvar)
    return decoder_forward(z, skips, sd), mu, logvar


# ------------------------------------------------------------------------------------------------ training step
def step_losses(recon, images, mu, logvar, quality_scores, semantic_score, baseline, *, recon_weight=1.0,
                kl_weight=0.1, quality_weight=0.5, semantic_weight=0.5, reward_scale=0.1, baseline_momentum=0.9,
                accum=1):
    """train_hybrid.py:859-896. Returns (vae_loss, teacher_loss, new_baseline, metrics dict of python floats)."""
    recon_loss = F.mse_loss(recon, images)
    kl_loss = -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())
    quality_reward = quality_scores.mean(1, keepdim=True)
    total_reward = quality_reward + semantic_weight * semantic_score
    tr = total_reward.mean().item()
    baseline = tr if baseline is None else baseline_momentum * baseline + (1 - baseline_momentum) * tr
    advantage = (total_reward - baseline).detach() * reward_scale
    pg_loss = -(advantage * recon_loss).mean()
    vae_loss = (recon_weight * recon_loss + kl_weight * kl_loss + pg_loss) / accum
    quality_loss = -quality_scores.mean()
    teacher_loss = quality_weight * quality_loss / accum
    metrics = {
        "recon_loss": recon_loss.item(), "kl_loss": kl_loss.item(), "quality_loss": quality_loss.item(),
        "pg_loss": pg_loss.item(), "semantic_reward": semantic_score.mean().item(),
        "quality_reward": quality_reward.mean().item(), "baseline": baseline, "advantage": advantage.mean().item(),
        "vae_loss": vae_loss.item(), "teacher_loss": teacher_loss.item(),
        "total_loss": (vae_loss + teacher_loss).item(), "quality_scores": quality_scores.mean().item(),
    }
    return vae_loss, teacher_loss, baseline, metrics


def train_step(images, vae_sd, teacher_sd, eps, baseline=None, semantics="reference", grad_mode="reference",
               masks_a=None, masks_b=None, **loss_kw):
    """One _process_batch (train_hybrid.py:838-904) up to and including both backward passes.
    `vae_sd` / `teacher_sd` hold leaf tensors with requires_grad=True for parameters; gradients land in .grad
    (parameters outside the reference's executed gradient set keep .grad None). Returns (metrics, recon, baseline)."""
    images = images.detach()
    recon, mu, logvar = vae_forward(images, vae_sd, eps)
    with torch.no_grad():
        teacher_forward(images, teacher_sd, True, semantics, grad_mode, masks_a, no_grad_pass=True)   # pass A
    out = teacher_forward(recon.detach(), teacher_sd, True, semantics, grad_mode, masks_b)             # pass B
    vae_loss, teacher_loss, baseline, metrics = step_losses(
        recon, images, mu, logvar, out["quality_scores"], out["semantic_score"], baseline, **loss_kw)
    vae_loss.backward()
    teacher_loss.backward()
    return metrics, recon.detach(), baseline


def cosine_warm_restart_lr(step, lr0, t0=10, t_mult=2, eta_min=1e-6):
    """CosineAnnealingWarmRestarts stepped once per optimizer step (train_hybrid.py:514-527, 925-926)."""
    t_i, t_cur = t0, step
    while t_cur >= t_i:
        t_cur -= t_i
        t_i *= t_mult
    return eta_min + (lr0 - eta_min) * (1 + math.cos(math.pi * t_cur / t_i)) / 2
