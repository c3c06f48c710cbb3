"""Drop-in for the reference's `lunar_evaluator` module (MeryylleA/Lunaris-Orion lunar_evaluator.py:57-462).

Same classes, constructor signatures, sub-module tree (=> identical state_dict keys, parameter order and init-RNG
consumption) and forward contract as the reference, but `forward` runs the B200-native kernels of
liblunaris_b200.so on NHWC bf16 tensors. Semantics are the reference's AS EXECUTED:

  * attention: 32-token block-local softmax attention with the chunk-INDEX scatter of lunar_evaluator.py:203-216,
  * gradients: the set the reentrant checkpoints leave alive (shortcut.*, layer_scale, attention.proj.*, conv2.* of
    blocks 1,2, gate.*, quality_heads.*) - everything else keeps grad None,
  * BatchNorm running statistics: one update per forward plus one per checkpoint recompute.

There is no CPU path: calling forward on a CPU tensor raises.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _capi, _host, ops
from ._capi import check

_HEADS = 8
_CHUNK = 32
_SLOPE = 0.2


def mish(x):
    """x * tanh(softplus(x)) (reference lunar_evaluator.py:48-50)."""
    return x * torch.tanh(F.softplus(x))


# ====================================================================================================== module tree
def _conv_act_bn(cin, cout, k, pad, groups=1):
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=k, padding=pad, groups=groups), nn.LeakyReLU(_SLOPE),
                         nn.BatchNorm2d(cout))


class PixelArtFeatureExtractor(nn.Module):
    """Parameter container mirroring lunar_evaluator.py:57-112 (conv1, three depthwise+pointwise branches, fusion)."""

    def __init__(self, in_channels=3, dropout_rate=0.1, feature_dim=128):
        super().__init__()
        self.conv1 = _conv_act_bn(in_channels, 32, 3, 1)

        def branch(k):
            return nn.Sequential(nn.Conv2d(32, 32, kernel_size=k, padding=k // 2, groups=32),
                                 nn.Conv2d(32, 64, kernel_size=1), nn.LeakyReLU(_SLOPE), nn.BatchNorm2d(64))
        self.edge_branch = branch(3)
        self.color_branch = branch(5)
        self.detail_branch = branch(3)
        self.dropout = nn.Dropout(dropout_rate)
        self.fusion = _conv_act_bn(64 * 3, feature_dim, 1, 0)

    def forward(self, x):
        _host.require_no_grad("PixelArtFeatureExtractor.forward", x)
        feats, _ = _fe_forward(self, x, 1)
        return _to_nchw(feats, x.shape[0], x.shape[2], x.shape[3])


class PixelArtAttention(nn.Module):
    """Parameter container mirroring lunar_evaluator.py:119-145; forward = as-executed block-local attention."""

    def __init__(self, in_channels, num_heads=8, rel_pos_size=8, dropout=0.1, chunk_size=64):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = in_channels // num_heads
        self.chunk_size = chunk_size
        assert self.head_dim * num_heads == in_channels, "in_channels must be divisible by num_heads"
        self.qkv = nn.Conv2d(in_channels, in_channels * 3, kernel_size=1)
        self.proj = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.rel_pos_h = nn.Parameter(torch.randn(1, num_heads, rel_pos_size, 1) * 0.02)
        self.rel_pos_w = nn.Parameter(torch.randn(1, num_heads, 1, rel_pos_size) * 0.02)
        self.attn_drop = nn.Dropout(dropout)
        self.proj_drop = nn.Dropout(dropout)
        self.register_buffer('rel_pos_cache', None)
        self.register_buffer('last_spatial_shapes', torch.zeros(2))

    def _touch_rel_pos(self, H, W, device):
        """Keeps the two state_dict buffers the reference creates on first use (lunar_evaluator.py:174-186). The
        bias itself is constant along the softmax axis and never changes an output, so no kernel consumes it."""
        if self.rel_pos_cache is not None and self.rel_pos_cache.shape[2] == H * W:
            return
        with torch.no_grad():
            rel_h = F.interpolate(self.rel_pos_h, size=(H, 1), mode='bilinear', align_corners=True)
            rel_w = F.interpolate(self.rel_pos_w, size=(1, W), mode='bilinear', align_corners=True)
            self.rel_pos_cache = (rel_h.expand(-1, -1, -1, W) + rel_w.expand(-1, -1, H, -1)).reshape(
                1, self.num_heads, H * W).unsqueeze(-1)
            self.last_spatial_shapes = torch.tensor([H, W], device=device)

    def forward(self, x):
        _host.require_no_grad("PixelArtAttention.forward", x)
        B, C, H, W = x.shape
        one, zero = torch.ones(C, device=x.device), torch.zeros(C, device=x.device)
        h2, _ = _attention_forward(self, _to_nhwc(x).view(B, H * W, C), one, zero, None, B, H, W, self.training,
                                   save=False, tag="attention")
        return _to_nchw(h2, B, H, W)


class ExpertBlock(nn.Module):
    """Parameter container mirroring lunar_evaluator.py:234-258."""

    def __init__(self, in_channels, out_channels, dropout_rate=0.1, rel_pos_size=8, layer_scale_init=0.1):
        super().__init__()
        def conv_stack(cin):
            return nn.Sequential(nn.Conv2d(cin, out_channels, kernel_size=3, padding=1), nn.LeakyReLU(_SLOPE),
                                 nn.BatchNorm2d(out_channels), nn.Dropout2d(dropout_rate))
        self.conv1 = conv_stack(in_channels)
        self.attention = PixelArtAttention(out_channels, rel_pos_size=rel_pos_size, dropout=dropout_rate)
        self.conv2 = conv_stack(out_channels)
        if in_channels != out_channels:
            self.shortcut = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1),
                                          nn.BatchNorm2d(out_channels))
        else:
            self.shortcut = nn.Identity()
        self.layer_scale = nn.Parameter(torch.ones(1, out_channels, 1, 1) * layer_scale_init)

    def forward(self, x):
        _host.require_no_grad("ExpertBlock.forward", x)
        B, C, H, W = x.shape
        out, _ = _block_forward(self, _to_nhwc(x), B, H, W, self.training, 1, save=False, pool=None, tag="block")
        return _to_nchw(out, B, H, W)


class LunarMoETeacher(nn.Module):
    """Mixture-of-experts quality teacher, drop-in for lunar_evaluator.py:278-462."""

    def __init__(self, num_experts=4, feature_dim=128, dropout_rate=0.1, rel_pos_size=8, use_checkpointing=True,
                 expert_layers=3, intermediate_dim=256, embedding_dim=64):
        super().__init__()
        self.num_experts = num_experts
        self.feature_dim = feature_dim
        self.dropout_rate = dropout_rate
        self.rel_pos_size = rel_pos_size
        self.use_checkpointing = use_checkpointing
        self.expert_layers = expert_layers
        self.intermediate_dim = intermediate_dim
        self.embedding_dim = embedding_dim

        self.feature_extractor = PixelArtFeatureExtractor(in_channels=3, dropout_rate=dropout_rate, feature_dim=128)

        def expert():
            dims = [128] + [feature_dim] * expert_layers
            return nn.Sequential(*[ExpertBlock(dims[i], dims[i + 1], dropout_rate=dropout_rate,
                                               rel_pos_size=rel_pos_size) for i in range(expert_layers)])
        self.experts = nn.ModuleList([expert() for _ in range(num_experts)])
        self.gate = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(128, intermediate_dim),
                                  nn.LeakyReLU(_SLOPE), nn.Dropout(dropout_rate),
                                  nn.Linear(intermediate_dim, num_experts), nn.Softmax(dim=1))

        def head(hidden, out, sigmoid=False):
            layers = [nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.LayerNorm(feature_dim), nn.Linear(feature_dim, hidden),
                      nn.LeakyReLU(_SLOPE), nn.Dropout(dropout_rate), nn.Linear(hidden, out)]
            return nn.Sequential(*(layers + ([nn.Sigmoid()] if sigmoid else [])))
        self.quality_heads = nn.ModuleList([head(intermediate_dim // 4, 4) for _ in range(num_experts)])
        self.semantic_head = head(intermediate_dim // 2, 1, sigmoid=True)
        self.style_net = head(intermediate_dim // 2, embedding_dim)
        self.prompt_net = head(intermediate_dim // 2, embedding_dim)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='leaky_relu')
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, (nn.BatchNorm2d, nn.LayerNorm)):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)

    def forward(self, x, prompt_embedding=None):
        """Same contract as lunar_evaluator.py:409-462. Differentiable in training mode (the reference's executed
        gradient set); in eval mode the outputs are returned without an autograd graph."""
        if not x.is_cuda:
            raise _capi.LunarisB200Error("lunaris_orion_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        _host.require_no_grad("LunarMoETeacher.forward w.r.t. its input (the reference's reentrant checkpoints would "
                              "then produce a different gradient set; pass x.detach())", x)
        grad_on = torch.is_grad_enabled() and self.training
        if grad_on and not self.use_checkpointing:
            raise _capi.LunarisB200Error(
                "LunarMoETeacher(use_checkpointing=False) in training mode under autograd is not implemented: without "
                "the reentrant checkpoints the reference back-propagates into the feature extractor, conv1 and qkv "
                "(a different gradient set from the one this path reproduces). Use use_checkpointing=True.")
        if not grad_on and torch.is_grad_enabled():
            _warn_once("LunarMoETeacher.forward in eval mode returns outputs without an autograd graph "
                       "(the reference would back-propagate through them); wrap the call in torch.no_grad().")
        B, _, H, W = x.shape
        params = _trunk_grad_params(self) if grad_on else []
        res = _TeacherTrunk.apply(self, x.detach(), grad_on, *params)
        pooled_fe, pooled = res[0], res[1]                    # channel sums: [B,128] and [E,B,C]
        fmaps = res[2:]
        hp = _head_params(self)
        quality, weights, style, prompt, sem = _HeadsFn.apply(self, 1.0 / float(H * W), grad_on, pooled_fe, pooled,
                                                              *(hp if grad_on else []))
        return {
            'quality_scores': quality,
            'expert_weights': weights,
            'style_embedding': style,
            'prompt_embedding': prompt,
            'semantic_score': sem,
            'feature_maps': None if self.training else [_to_nchw(f, B, H, W) for f in fmaps],
        }


_HEAD_TAGS = ("gate_drop", "quality_drop.{e}", "semantic_drop", "style_drop", "prompt_drop")


def _head_modules(teacher):
    """Heads in the order of the C ABI (include/lunaris_b200.h lun_heads_fwd): gate, quality[0..E), semantic, style,
    prompt. Each entry: (LayerNorm or None, first Linear, Dropout, second Linear, dropout-trace tag)."""
    g = teacher.gate
    out = [(None, g[2], g[4], g[5], "gate_drop")]
    for e, q in enumerate(teacher.quality_heads):
        out.append((q[2], q[3], q[5], q[6], f"quality_drop.{e}"))
    for seq, tag in ((teacher.semantic_head, "semantic_drop"), (teacher.style_net, "style_drop"),
                     (teacher.prompt_net, "prompt_drop")):
        out.append((seq[2], seq[3], seq[5], seq[6], tag))
    return out


def _head_params(teacher):
    """Flat parameter list of the heads, 6 slots per head {ln.weight, ln.bias, l1.weight, l1.bias, l2.weight, l2.bias}
    (the gate has no LayerNorm: 4 tensors)."""
    ps = []
    for ln, l1, _, l2, _ in _head_modules(teacher):
        if ln is not None:
            ps += [ln.weight, ln.bias]
        ps += [l1.weight, l1.bias, l2.weight, l2.bias]
    return ps


class _HeadsFn(torch.autograd.Function):
    """(pooled FE sums [B,128], pooled expert sums [E,B,C]) -> (quality_scores, expert_weights, style_embedding,
    prompt_embedding, semantic_score): lunar_evaluator.py:417-456 in one launch (csrc/teacher_heads.cu); backward in
    two. Gradients flow to the pooled expert sums and to every head parameter whose output reaches the loss; the
    feature-extractor sums take none (detach rule (i))."""

    @staticmethod
    def _tables(teacher, B, seeds):
        mods = _head_modules(teacher)
        E = teacher.num_experts
        ptrs = []
        for ln, l1, _, l2, _ in mods:
            for t in ((ln.weight, ln.bias) if ln is not None else (None, None)) + (l1.weight, l1.bias, l2.weight, l2.bias):
                if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
                    raise _capi.LunarisB200Error("Teacher head parameters must be contiguous fp32")
                ptrs.append(0 if t is None else t.data_ptr())
        params = (_capi.ctypes.c_void_p * len(ptrs))(*ptrs)
        dims = (_capi.ctypes.c_int * 10)(B, E, mods[0][1].in_features, teacher.feature_dim, teacher.embedding_dim,
                                          mods[0][1].out_features, mods[1][1].out_features,
                                          mods[E + 1][1].out_features, mods[E + 2][1].out_features,
                                          mods[E + 3][1].out_features)
        seed_arr = (_capi.ctypes.c_ulonglong * (E + 4))(*seeds)
        return params, dims, seed_arr

    @staticmethod
    def forward(ctx, teacher, inv_hw, grad_on, pooled_fe, pooled, *params):
        lib = _capi.lib()
        mods = _head_modules(teacher)
        E, C, emb = teacher.num_experts, teacher.feature_dim, teacher.embedding_dim
        B = pooled_fe.shape[0]
        dev = pooled_fe.device
        training = teacher.training
        ctx.set_materialize_grads(False)       # unused outputs must stay None: their heads then keep grad None
        p = float(mods[0][2].p) if training else 0.0
        if training and any(float(m[2].p) != p for m in mods):
            raise _capi.LunarisB200Error("the fused heads kernel takes one dropout probability for all heads")
        seeds = [_cpu_seed(m[4], B, m[1].out_features) if p > 0 else 0 for m in mods]
        tabs = _HeadsFn._tables(teacher, B, seeds)
        sizes = (_capi.ctypes.c_int * 3)()
        check(lib.lun_heads_buffer_sizes(tabs[1], sizes), "lun_heads_buffer_sizes")
        pooled_fe = pooled_fe.detach().float().contiguous()
        pooled = pooled.detach().float().contiguous()
        f32 = dict(device=dev, dtype=torch.float32)
        quality, weights = torch.empty(B, 4, **f32), torch.empty(B, E, **f32)
        style, prompt, sem = torch.empty(B, emb, **f32), torch.empty(B, emb, **f32), torch.empty(B, 1, **f32)
        save = torch.empty(B, sizes[0], **f32) if grad_on else None
        check(lib.lun_heads_fwd(tabs[0], tabs[1], tabs[2], inv_hw, _SLOPE, p, 1 if training else 0,
                                pooled_fe.data_ptr(), pooled.data_ptr(), quality.data_ptr(), weights.data_ptr(),
                                style.data_ptr(), prompt.data_ptr(), sem.data_ptr(), _p(save), _stream()),
              "lun_heads_fwd")
        if grad_on:
            ctx.teacher, ctx.inv_hw, ctx.p, ctx.seeds, ctx.sizes = teacher, inv_hw, p, seeds, tuple(sizes)
            ctx.save_for_backward(pooled_fe, pooled, weights, save)
        return quality, weights, style, prompt, sem

    @staticmethod
    def backward(ctx, g_quality, g_weights, g_style, g_prompt, g_sem):
        lib = _capi.lib()
        teacher = ctx.teacher
        pooled_fe, pooled, weights, save = ctx.saved_tensors
        E, C = teacher.num_experts, teacher.feature_dim
        B = pooled_fe.shape[0]
        dev = pooled_fe.device
        tabs = _HeadsFn._tables(teacher, B, ctx.seeds)
        f32 = dict(device=dev, dtype=torch.float32)
        gs = [None if g is None else g.detach().float().contiguous() for g in (g_quality, g_weights, g_style, g_prompt, g_sem)]
        # which heads' parameters take a gradient: those whose output reaches an incoming gradient
        wants = [gs[0] is not None or gs[1] is not None or gs[2] is not None or gs[3] is not None]      # gate
        wants += [gs[0] is not None] * E                                                               # quality heads
        wants += [gs[4] is not None, gs[2] is not None, gs[3] is not None]                             # semantic, style, prompt
        mods = _head_modules(teacher)
        grads, ptrs = [], []
        for (ln, l1, _, l2, _), want in zip(mods, wants):
            row = []
            for t in ((ln.weight, ln.bias) if ln is not None else (None, None)) + (l1.weight, l1.bias, l2.weight, l2.bias):
                g = torch.empty_like(t) if (want and t is not None and t.requires_grad) else None
                row.append(g)
                ptrs.append(0 if g is None else g.data_ptr())
            grads += row[2:] if ln is None else row
        gtab = (_capi.ctypes.c_void_p * len(ptrs))(*ptrs)
        dpre = torch.empty(B, ctx.sizes[1], **f32)
        dout2 = torch.empty(B, ctx.sizes[2], **f32)
        dxn = torch.empty(E + 3, B, C, **f32)
        d_pooled = torch.empty(E, B, C, **f32)
        check(lib.lun_heads_bwd(tabs[0], tabs[1], tabs[2], ctx.inv_hw, _SLOPE, ctx.p, pooled_fe.data_ptr(),
                                pooled.data_ptr(), weights.data_ptr(), save.data_ptr(), _p(gs[0]), _p(gs[1]), _p(gs[2]),
                                _p(gs[3]), _p(gs[4]), dpre.data_ptr(), dout2.data_ptr(), dxn.data_ptr(),
                                d_pooled.data_ptr(), gtab, _stream()), "lun_heads_bwd")
        return (None, None, None, None, d_pooled, *grads)


_warned = set()


def _warn_once(msg):
    if msg not in _warned:
        _warned.add(msg)
        import warnings
        warnings.warn(msg, stacklevel=3)


# ====================================================================================================== plumbing
def _to_nhwc(x):
    return x.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _to_nchw(y, B, H, W):
    return y.view(B, H, W, -1).permute(0, 3, 1, 2).float()


def _packed(param, kind):
    """bf16 kernel-layout shadow of an fp32 parameter (cached on the parameter; refreshed when the optimizer changes
    it in place or its storage moves, see _host.cached)."""
    def build():
        if kind == "fwd":
            return ops.pack_conv_weight(param)
        if kind == "dgrad":
            return ops.pack_conv_weight_dgrad(param)
        if kind == "f32":
            return param.detach().float().contiguous()
        if kind in ("q_fwd", "kv_fwd"):          # qkv conv weight [3C, C, 1, 1] split into its Q and K|V row blocks
            C = param.shape[1]
            return ops.pack_conv_weight(param.detach()[:C] if kind == "q_fwd" else param.detach()[C:])
        if kind in ("q_f32", "kv_f32"):          # qkv bias [3C]
            C = param.shape[0] // 3
            return (param.detach()[:C] if kind == "q_f32" else param.detach()[C:]).float().contiguous()
        raise KeyError(kind)
    return _host.cached(param, kind, (param,), build)


_stream = _capi.raw_stream


def _cpu_seed(tag, *shape):
    """Seed of one counter-RNG dropout launch, drawn from torch's CPU generator (rank r is seeded with seed + r).
    `shape` (head dropouts: [B, hidden]) goes to the dropout trace with the seed."""
    seed = int(torch.randint(0, 2 ** 62, (1,), device="cpu").item())
    _host.trace("seed", tag, (seed, *shape) if shape else seed)
    return seed


def _p(t):
    return 0 if t is None else t.data_ptr()


def _bn_train(bn, stats, n, n_updates):
    """Batch statistics -> (scale, shift, mean, rstd); running buffers updated n_updates times."""
    C = bn.num_features
    out = torch.empty(4, C, device=stats.device, dtype=torch.float32)
    check(_capi.lib().lun_bn_finalize(stats.data_ptr(), float(n), bn.weight.data_ptr(), bn.bias.data_ptr(),
                                      bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                                      bn.num_batches_tracked.data_ptr(), n_updates, bn.momentum, bn.eps,
                                      out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), C,
                                      _stream()), "lun_bn_finalize")
    return out[0], out[1], out[2], out[3]


def _bn_eval(bn):
    scale = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
    return scale.contiguous(), (bn.bias.detach() - bn.running_mean * scale).contiguous()


def _affine(x, B, HW, C, scale, shift, mask2d=None, ls=None, identity=None, id_scale=None, id_shift=None, pool=None,
            seed=0, drop_p=0.0):
    y = torch.empty_like(x)
    check(_capi.lib().lun_affine_fwd_bf16(x.data_ptr(), _p(scale), _p(shift), _p(mask2d), _p(ls), _p(identity),
                                          _p(id_scale), _p(id_shift), y.data_ptr(), _p(pool), seed, float(drop_p), B,
                                          HW, C, _SLOPE, _stream()), "lun_affine_fwd_bf16")
    return y


def _drop2d_mask(B, C, p, device, tag):
    """Dropout2d keep-mask [B,C]: 0 or bf16(1/(1-p)) (what the reference's bf16 feature_dropout multiplies by)."""
    keep = torch.empty(B, C, device=device, dtype=torch.float32).bernoulli_(1.0 - p)
    _host.trace("mask2d", tag, keep)
    return keep * float(torch.tensor(1.0 / (1.0 - p)).to(torch.bfloat16))


# ====================================================================================================== forward pieces
def _fe_forward(fe, x, n_updates):
    """PixelArtFeatureExtractor.forward (lunar_evaluator.py:105-112). x: NCHW fp32. Returns NHWC bf16 [B,HW,128]
    features and their per-image channel sums [B,128]."""
    lib = _capi.lib()
    B, _, H, W = x.shape
    HW, dev = H * W, x.device
    training = fe.training
    x = x.detach().float().contiguous()
    p_drop = fe.dropout.p if training else 0.0

    c1, bn1 = fe.conv1[0], fe.conv1[2]
    y0 = torch.empty(B, HW, 32, device=dev, dtype=torch.bfloat16)
    st = _host.zeros(64, device=dev)
    check(lib.lun_fe_conv1(x.data_ptr(), _packed(c1.weight, "f32").data_ptr(), _packed(c1.bias, "f32").data_ptr(),
                           y0.data_ptr(), st.data_ptr(), B, H, W, _SLOPE, _stream()), "lun_fe_conv1")
    sc0, sh0 = _bn_train(bn1, st, B * HW, n_updates)[:2] if training else _bn_eval(bn1)

    branches = (fe.edge_branch, fe.color_branch, fe.detail_branch)
    keep = [[_packed(getattr(br[i], a), "f32") for br in branches] for i, a in
            ((0, "weight"), (0, "bias"), (1, "weight"), (1, "bias"))]
    arrs = [(_capi.ctypes.c_void_p * 3)(*[t.data_ptr() for t in ts]) for ts in keep]
    cat = torch.empty(B, HW, 192, device=dev, dtype=torch.bfloat16)
    check(lib.lun_fe_branches(y0.data_ptr(), sc0.data_ptr(), sh0.data_ptr(), arrs[0], arrs[1], arrs[2], arrs[3],
                              cat.data_ptr(), B, H, W, _SLOPE, _stream()), "lun_fe_branches")
    if training:
        st = _host.zeros(2 * 192, device=dev)
        check(lib.lun_channel_stats_bf16(cat.data_ptr(), B * HW, 192, st.data_ptr(), _stream()),
              "lun_channel_stats_bf16")
        sums, sqs = st[:192], st[192:]
        aff = [_bn_train(br[3], torch.cat([sums[64 * i:64 * i + 64], sqs[64 * i:64 * i + 64]]).contiguous(), B * HW,
                         n_updates)[:2] for i, br in enumerate(branches)]
    else:
        aff = [_bn_eval(br[3]) for br in branches]
    scale = torch.cat([a[0] for a in aff]).contiguous()
    shift = torch.cat([a[1] for a in aff]).contiguous()
    cat_n = _affine(cat, B, HW, 192, scale, shift, seed=_cpu_seed("fe_drop") if p_drop > 0 else 0, drop_p=p_drop)

    cf, bnf = fe.fusion[0], fe.fusion[2]
    C = cf.out_channels
    st = _host.zeros(2 * C, device=dev) if training else None
    f_pre = ops.conv2d_fprop(cat_n.view(B, H, W, 192), _packed(cf.weight, "fwd"), 1, 1, 0,
                             bias=_packed(cf.bias, "f32"), act_leaky=True, stats=st, slope=_SLOPE)
    scf, shf = _bn_train(bnf, st, B * HW, n_updates)[:2] if training else _bn_eval(bnf)
    pooled = _host.zeros(B, C, device=dev)
    feats = _affine(f_pre.view(B, HW, C), B, HW, C, scf, shf, pool=pooled)
    return feats, pooled


def _folded_attention_weights(att):
    """Weight-only precompute for the K/V-free attention (refreshed when qkv changes; in reference mode it never
    does - qkv receives no gradient): Mq [8C, C] = Wk_h^T Wq_h / sqrt(hd) stacked over heads, cq [8C] = Wk_h^T bq_h /
    sqrt(hd), and the value projection per head Wv [h][hd][C] (lun_head_linear_bf16: eight [hd, C] GEMMs in one launch
    instead of a block-diagonal [C, 8C] one). bf16 operands, fp32 products."""
    w, bq = att.qkv.weight, att.qkv.bias

    def build():
        C, h = att.qkv.in_channels, att.num_heads
        hd = C // h
        wf = w.detach().reshape(3, h, hd, C).to(torch.bfloat16).float()          # [3][head][d][c]
        bf = bq.detach().reshape(3, h, hd).float()
        scale = float(torch.tensor(hd ** -0.5).to(torch.bfloat16))
        mq = torch.einsum("hdk,hdc->hkc", wf[1], wf[0]).mul_(scale).reshape(h * C, C)
        cq = torch.einsum("hdk,hd->hk", wf[1], bf[0]).mul_(scale).reshape(h * C)
        if hd % 32 == 0:
            wv = wf[2].to(torch.bfloat16).contiguous()                           # [head][d][c]: one GEMM per head
        else:           # head_dim below the kernel's 32-channel output block (feature_dim < 256): block-diagonal [C, 8C]
            wv = torch.zeros(C, h * C, device=w.device)
            for i in range(h):
                wv[i * hd:(i + 1) * hd, i * C:(i + 1) * C] = wf[2, i]
            wv = wv.to(torch.bfloat16).contiguous()
        return (mq.to(torch.bfloat16).contiguous(), cq.contiguous(), wv, bf[2].reshape(C).contiguous())
    return _host.cached(att, "fold", (w, bq), build)


def _attention_forward(att, y1, sc1, sh1, m1, B, H, W, training, save, tag):
    """PixelArtAttention.forward as executed (lunar_evaluator.py:189-227) on the pre-BatchNorm conv1 output y1
    [B,HW,C] (x1 = drop2d(bn(y1)) is applied on load, never written). Only N/32+31 rows survive the reference's
    chunk-index scatter; for those rows the K and V projections are folded into the query / output side (see
    csrc/attn_fold.cu), then proj runs on the surviving rows, bias elsewhere, then proj_drop.
    Returns conv2's input h2 [B,HW,C] and what proj's backward needs."""
    lib = _capi.lib()
    C = att.qkv.in_channels
    HW = H * W
    nq = HW // _CHUNK + _CHUNK - 1
    nq_pad = (nq + 7) // 8 * 8
    heads = att.num_heads
    p_attn = att.attn_drop.p if training else 0.0
    p_proj = att.proj_drop.p if training else 0.0
    att._touch_rel_pos(H, W, y1.device)
    mq, cq, wv, bv = _folded_attention_weights(att)
    # rows nq .. nq_pad-1 of xq / qt / xbar are padding that no kernel reads back; att_small's are zeroed below
    xq = torch.empty(B, nq_pad, C, device=y1.device, dtype=torch.bfloat16)
    check(lib.lun_gather_query_rows_affine_bf16(y1.data_ptr(), sc1.data_ptr(), sh1.data_ptr(), _p(m1), xq.data_ptr(),
                                                B, HW, C, nq_pad, _stream()), "lun_gather_query_rows_affine_bf16")
    qt = ops.linear_fprop(xq.view(B * nq_pad, C), mq, cq, out_f32=False)              # [B*nq_pad, heads*C]
    xbar = torch.empty(B * nq_pad, heads * C, device=y1.device, dtype=torch.bfloat16)
    seed_attn = _cpu_seed(tag + ".attn_drop") if p_attn > 0 else 0
    check(lib.lun_attn_fold_rows_bf16(y1.data_ptr(), sc1.data_ptr(), sh1.data_ptr(), _p(m1), qt.data_ptr(),
                                      xbar.data_ptr(), B, HW, C, heads, nq_pad, seed_attn, float(p_attn), _stream()),
          "lun_attn_fold_rows_bf16")
    att_small = (ops.head_linear(xbar, wv, bv) if wv.dim() == 3
                 else ops.linear_fprop(xbar, wv, bv, out_f32=False)).view(B, nq_pad, C)
    if nq_pad > nq:
        att_small[:, nq:].zero_()                # padding rows enter proj's weight gradient: they must be finite
    wp = _packed(att.proj.weight, "fwd")
    proj_small = ops.linear_fprop(att_small.view(B * nq_pad, C), wp.view(C, C), _packed(att.proj.bias, "f32"),
                                  out_f32=False)
    seed = _cpu_seed(tag + ".proj_drop") if p_proj > 0 else 0
    h2 = torch.empty(B, HW, C, device=y1.device, dtype=torch.bfloat16)
    check(lib.lun_proj_expand_bf16(proj_small.data_ptr(), _packed(att.proj.bias, "f32").data_ptr(), h2.data_ptr(), B,
                                   HW, C, nq, nq_pad, seed, float(p_proj), _stream()), "lun_proj_expand_bf16")
    saved = dict(att_small=att_small, seed=seed, p_proj=p_proj, nq=nq, nq_pad=nq_pad) if save else None
    return h2, saved


def _input_moments(x, B, HW, Cin):
    """First and second moments of a block input over all pixels: (sum x [Cin], sum x x^T [Cin,Cin]), fp32. The Gram
    matrix is one launch of the tcgen05 weight-gradient kernel (dY = X)."""
    st = _host.zeros(2 * Cin, device=x.device)
    check(_capi.lib().lun_channel_stats_bf16(x.data_ptr(), B * HW, Cin, st.data_ptr(), _stream()),
          "lun_channel_stats_bf16")
    x4 = x.view(B, 1, HW, Cin)
    gram = ops.conv2d_wgrad(x4, x4, 1, 1, 0).reshape(Cin, Cin)
    return st[:Cin], gram


def _shortcut_stats(conv, xmom, n):
    """[sum y | sum y^2] per output channel of the 1x1 shortcut conv y = W x + b (ExpertBlock.shortcut[0],
    lunar_evaluator.py:254-257) WITHOUT reading y: the conv is linear and feeds its BatchNorm directly, so
    sum y = W sum(x) + n b and sum y^2 = diag(W G W^T) + 2 b (W sum x) + n b^2 with G = sum x x^T. The four experts share
    one G, and the conv itself runs without the statistics epilogue (which cost it 2x: 0.43 -> 0.9 ms at C3)."""
    sx, gram = xmom
    w = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).to(torch.bfloat16).float()
    b = conv.bias.detach().float()
    wsx = w @ sx
    s1 = wsx + n * b
    s2 = ((w @ gram) * w).sum(1) + 2.0 * b * wsx + n * b * b
    return torch.cat([s1, s2]).contiguous()


def _block_forward(blk, x, B, H, W, training, n_updates, save, pool, tag, xmom=None):
    """ExpertBlock.forward (lunar_evaluator.py:260-275) on NHWC bf16 x [B,HW,Cin]. Returns (out [B,HW,C], saved).
    xmom: _input_moments(x) when the caller already has them (the experts' first blocks share one input)."""
    HW = H * W
    dev = x.device
    c1, bn1, d1 = blk.conv1[0], blk.conv1[2], blk.conv1[3]
    c2, bn2, d2 = blk.conv2[0], blk.conv2[2], blk.conv2[3]
    C, Cin = c1.out_channels, c1.in_channels
    p2d = d1.p if training else 0.0

    st = _host.zeros(2 * C, device=dev) if training else None
    # the statistics epilogue hides behind a K = 9 * 512 main loop; behind the thin first conv of an expert (K = 9 * 128)
    # it costs more than one separate pass over the 1 GB output, so that conv runs without it
    epi_stats = training and Cin * 9 >= 2304
    y1 = ops.conv2d_fprop(x.view(B, H, W, Cin), _packed(c1.weight, "fwd"), 3, 1, 1, bias=_packed(c1.bias, "f32"),
                          act_leaky=True, stats=st if epi_stats else None, slope=_SLOPE)
    if training and not epi_stats:
        check(_capi.lib().lun_channel_stats_bf16(y1.data_ptr(), B * HW, C, st.data_ptr(), _stream()),
              "lun_channel_stats_bf16")
    sc1, sh1 = _bn_train(bn1, st, B * HW, n_updates)[:2] if training else _bn_eval(bn1)
    m1 = _drop2d_mask(B, C, p2d, dev, tag + ".drop2d_1") if p2d > 0 else None
    h2, att_saved = _attention_forward(blk.attention, y1.view(B, HW, C), sc1, sh1, m1, B, H, W, training, save, tag)
    del y1

    st = _host.zeros(2 * C, device=dev) if training else None
    y2 = ops.conv2d_fprop(h2.view(B, H, W, C), _packed(c2.weight, "fwd"), 3, 1, 1, bias=_packed(c2.bias, "f32"),
                          act_leaky=True, stats=st, slope=_SLOPE).view(B, HW, C)
    if training:
        sc2, sh2, mean2, rstd2 = _bn_train(bn2, st, B * HW, n_updates)
    else:
        (sc2, sh2), mean2, rstd2 = _bn_eval(bn2), None, None
    m2 = _drop2d_mask(B, C, d2.p, dev, tag + ".drop2d_2") if training and d2.p > 0 else None

    has_sc = not isinstance(blk.shortcut, nn.Identity)
    if has_sc:
        cs, bns = blk.shortcut[0], blk.shortcut[1]
        identity = ops.conv2d_fprop(x.view(B, H, W, Cin), _packed(cs.weight, "fwd"), 1, 1, 0,
                                    bias=_packed(cs.bias, "f32")).view(B, HW, C)
        if training:
            st = _shortcut_stats(cs, xmom if xmom is not None else _input_moments(x, B, HW, Cin), float(B * HW))
            isc, ish, imean, irstd = _bn_train(bns, st, B * HW, n_updates)
        else:
            (isc, ish), imean, irstd = _bn_eval(bns), None, None
    else:
        identity, isc, ish, imean, irstd = x, None, None, None, None
    ls = blk.layer_scale.detach().reshape(-1).contiguous()
    out = _affine(y2, B, HW, C, sc2, sh2, mask2d=m2, ls=ls, identity=identity, id_scale=isc, id_shift=ish, pool=pool)
    saved = None
    if save:
        saved = dict(att=att_saved, h2=h2, a2=y2, out=out, mean2=mean2, rstd2=rstd2, m2=m2, has_sc=has_sc)
        if has_sc:
            saved.update(sc_pre=identity, imean=imean, irstd=irstd, x=x)
    return out, saved


# ====================================================================================================== autograd trunk
def _trunk_grad_params(teacher):
    """Parameters that receive gradients in the reference's executed graph, fixed order."""
    ps = []
    for expert in teacher.experts:
        for b, blk in enumerate(expert):
            if b == 0:
                if not isinstance(blk.shortcut, nn.Identity):
                    ps += [blk.shortcut[0].weight, blk.shortcut[0].bias, blk.shortcut[1].weight, blk.shortcut[1].bias]
            else:
                ps += [blk.layer_scale, blk.attention.proj.weight, blk.attention.proj.bias, blk.conv2[0].weight,
                       blk.conv2[0].bias, blk.conv2[2].weight, blk.conv2[2].bias]
    return ps


def _block_tail_backward(B, HW, C, dout, gpool, out, bn_in, mean, rstd, gamma, ls, m2, slope_out, slope_a, want_dpre):
    """Backward through leaky_relu(residual) -> layer_scale -> Dropout2d -> BatchNorm -> LeakyReLU.
    Returns (dpre, dz, dbias_conv, t1, t2)."""
    lib = _capi.lib()
    dev = bn_in.device
    t = _host.zeros(2, C, device=dev)
    dpre = torch.empty(B, HW, C, device=dev, dtype=torch.bfloat16) if want_dpre else None
    check(lib.lun_block_bwd_reduce_bf16(_p(dout), _p(gpool), _p(out), bn_in.data_ptr(), mean.data_ptr(),
                                        rstd.data_ptr(), _p(m2), _p(dpre), t[0].data_ptr(), t[1].data_ptr(), B, HW, C,
                                        slope_out, _stream()), "lun_block_bwd_reduce_bf16")
    dz = torch.empty(B, HW, C, device=dev, dtype=torch.bfloat16)
    dbias = _host.zeros(C, device=dev)
    check(lib.lun_block_bwd_apply_bf16(_p(dpre), None if dpre is not None else _p(gpool),
                                       None if dpre is not None else _p(out), bn_in.data_ptr(), mean.data_ptr(),
                                       rstd.data_ptr(), gamma.data_ptr(), _p(ls), _p(m2), t[0].data_ptr(),
                                       t[1].data_ptr(), dz.data_ptr(), dbias.data_ptr(), B, HW, C, slope_out, slope_a,
                                       _stream()), "lun_block_bwd_apply_bf16")
    return dpre, dz, dbias, t[0], t[1]


class _TeacherTrunk(torch.autograd.Function):
    """images -> (sum-pooled FE features [B,128], sum-pooled expert outputs [E,B,C][, feature maps in eval]) with the
    reference's executed gradient set as backward."""

    @staticmethod
    def forward(ctx, teacher, x, grad_on, *params):
        with _host.zero_pool(x.device):
            return _TeacherTrunk._forward(ctx, teacher, x, grad_on)

    @staticmethod
    def _forward(ctx, teacher, x, grad_on):
        B, _, H, W = x.shape
        training = teacher.training
        feats, pooled_fe = _fe_forward(teacher.feature_extractor, x, 1)
        fmaps, saved = [], []
        C = teacher.feature_dim
        pooled = _host.zeros(len(teacher.experts), B, C, device=x.device)    # per-image channel sums of every expert
        xmom = _input_moments(feats, B, H * W, feats.shape[-1]) if training else None
        for e, expert in enumerate(teacher.experts):
            h = feats
            per = []
            for b, blk in enumerate(expert):
                last = b == len(expert) - 1
                # blocks whose checkpoint segment the reference re-runs in backward update BN stats twice
                recomputed = grad_on and b > 0
                h, sv = _block_forward(blk, h, B, H, W, training, 2 if recomputed else 1, save=grad_on,
                                       pool=pooled[e] if last else None, tag=f"experts.{e}.{b}",
                                       xmom=xmom if b == 0 else None)
                per.append(sv)
            saved.append(per)
            if not training:
                fmaps.append(h)
        ctx.teacher, ctx.saved, ctx.grad_on = teacher, saved, grad_on
        ctx.dims = (B, H, W)
        ctx.feats = feats if grad_on else None
        outs = (pooled_fe, pooled, *fmaps)
        ctx.mark_non_differentiable(pooled_fe, *fmaps)
        return outs

    @staticmethod
    def backward(ctx, g_fe, g_pooled, *g_maps):
        with _host.zero_pool(ctx.feats.device):
            return _TeacherTrunk._backward(ctx, g_pooled)

    @staticmethod
    def _backward(ctx, g_pooled):
        teacher = ctx.teacher
        B, H, W = ctx.dims
        HW = H * W
        grads = []
        for e, expert in enumerate(teacher.experts):
            C = expert[0].conv1[0].out_channels
            per_block = {}
            dout = None
            gpool = g_pooled[e].contiguous().float() if g_pooled is not None \
                else torch.zeros(B, C, device=ctx.feats.device)
            for b in range(len(expert) - 1, 0, -1):
                blk, sv = expert[b], ctx.saved[e][b]
                gamma, beta = blk.conv2[2].weight.detach(), blk.conv2[2].bias.detach()
                ls = blk.layer_scale.detach().reshape(-1).contiguous()
                dpre, dz, dbias2, t1, t2 = _block_tail_backward(
                    B, HW, C, dout, gpool if dout is None else None, sv["out"], sv["a2"], sv["mean2"], sv["rstd2"],
                    gamma, ls, sv["m2"], _SLOPE, _SLOPE, want_dpre=True)
                d_ls = (gamma * t2 + beta * t1).view(1, C, 1, 1)
                d_gamma, d_beta = ls * t2, ls * t1
                dz4 = dz.view(B, H, W, C)
                dw2 = ops.conv2d_wgrad(dz4, sv["h2"].view(B, H, W, C), 3, 1, 1)
                a = sv["att"]
                # conv2's data gradient with proj_drop's backward folded into its epilogue: the proj bias gradient
                # (column sums of mask * dh2 / (1-p)) comes out of the conv, only the nq surviving rows are re-read
                dpo = torch.empty(B, a["nq_pad"], C, device=dz.device, dtype=torch.bfloat16)
                if a["nq_pad"] > a["nq"]:
                    dpo[:, a["nq"]:].zero_()        # the gather kernel writes rows < nq; padding rows must be zero
                if ops.drop_sum_ok(B, H, W, C):
                    colsum = _host.zeros(2 * C, device=dz.device)
                    dh2 = ops.conv2d_dgrad(dz4, _packed(blk.conv2[0].weight, "dgrad"), 3, 1, 1, (H, W),
                                           drop_sum=(a["seed"], a["p_proj"], colsum))
                    dbp = colsum[:C]
                else:                                # small feature maps: a separate pass sums the masked gradient
                    dh2 = ops.conv2d_dgrad(dz4, _packed(blk.conv2[0].weight, "dgrad"), 3, 1, 1, (H, W))
                    dbp = _host.zeros(C, device=dz.device)
                del dz, dz4
                check(_capi.lib().lun_proj_bwd_gather_bf16(dh2.data_ptr(), dpo.data_ptr(),
                                                           None if ops.drop_sum_ok(B, H, W, C) else dbp.data_ptr(), B,
                                                           HW, C, a["nq"], a["nq_pad"], a["seed"], float(a["p_proj"]),
                                                           _stream()), "lun_proj_bwd_gather_bf16")
                del dh2
                dwp = ops.linear_wgrad(dpo.view(B * a["nq_pad"], C), a["att_small"].view(B * a["nq_pad"], C))
                per_block[b] = [d_ls, dwp.view(C, C, 1, 1), dbp, dw2, dbias2, d_gamma, d_beta]
                dout = dpre
            blk, sv = expert[0], ctx.saved[e][0]
            if sv["has_sc"]:
                bns = blk.shortcut[1]
                _, dzs, dbs, t1, t2 = _block_tail_backward(
                    B, HW, C, dout, gpool if dout is None else None, sv["out"], sv["sc_pre"], sv["imean"],
                    sv["irstd"], bns.weight.detach(), None, None, _SLOPE, 1.0, want_dpre=True)
                cin = blk.shortcut[0].in_channels
                dws = ops.conv2d_wgrad(dzs.view(B, H, W, C), sv["x"].view(B, H, W, cin), 1, 1, 0)
                grads += [dws, dbs, t2.clone(), t1.clone()]
            for b in range(1, len(expert)):
                grads += per_block[b]
        ctx.saved = None
        ctx.feats = None
        return (None, None, None, *grads)
