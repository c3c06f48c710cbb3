"""Drop-in for the reference's `lunar_generate` module (MeryylleA/Lunaris-Orion lunar_generate.py:24-291).

Same classes / constructor signatures / sub-module tree (identical state_dict keys, parameter order, init RNG) and the
same forward contract, with the arithmetic on the B200-native kernels of liblunaris_b200.so (NHWC bf16 activations,
tcgen05 implicit-GEMM convs, fused GroupNorm+Mish). All 72 VAE tensors receive gradients, as in the reference.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _capi, _host, ops
from ._capi import check


def mish(x):
    """x * tanh(softplus(x)) (reference lunar_generate.py:24-26)."""
    return x * torch.tanh(F.softplus(x))


# ====================================================================================================== module tree
def _conv_gn_mish(cin, cout, stride=1):
    return [nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1), nn.GroupNorm(8, cout), nn.Mish()]


class ResBlock(nn.Module):
    """Parameter container mirroring lunar_generate.py:28-53."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv1 = nn.Sequential(*_conv_gn_mish(in_channels, out_channels))
        self.conv2 = nn.Sequential(*_conv_gn_mish(out_channels, out_channels))
        self.shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1) if in_channels != out_channels \
            else nn.Identity()

    def forward(self, x):
        """mish(conv2(conv1(x)) + shortcut(x)) (lunar_generate.py:47-53), incl. the 1x1 shortcut conv of the
        in != out case. Inference-only as a standalone module (the VAE's fused forward / backward is the
        differentiable path)."""
        if not x.is_cuda:
            raise _capi.LunarisB200Error("lunaris_orion_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        _host.require_no_grad("ResBlock.forward", x)
        B, cin, H, W = x.shape
        C = self.conv1[0].out_channels
        if H != W or cin % 64 or C % 64:
            raise _capi.LunarisB200Error("ResBlock kernel needs square maps and channel counts that are multiples of 64")
        a = x.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).view(B, H * W, cin)
        out, _ = _resblock_forward(self, a, B, H, C, save=False)
        return out.view(B, H, W, C).permute(0, 3, 1, 2).float()


class SelfAttention2d(nn.Module):
    """Parameter container mirroring lunar_generate.py:56-78 (the reference defines it but never instantiates it)."""

    def __init__(self, in_channels):
        super().__init__()
        self.query_conv = nn.Conv2d(in_channels, in_channels // 8, kernel_size=1)
        self.key_conv = nn.Conv2d(in_channels, in_channels // 8, kernel_size=1)
        self.value_conv = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.gamma = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        """y = gamma * softmax(q^T k) v + x (lunar_generate.py:66-78) through the flash-style tcgen05 kernels
        (csrc/flash_attn2d_sm100.cu forward + value gradient, csrc/flash_attn2d_bwd_sm100.cu query / key gradients).
        The N x N attention matrix is never formed in either direction."""
        if not x.is_cuda:
            raise _capi.LunarisB200Error("lunaris_orion_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        B, C, H, W = x.shape
        N, dq = H * W, C // 8
        if N % 128 or C % 64 or dq > 64:
            raise _capi.LunarisB200Error("SelfAttention2d kernel needs H*W % 128 == 0, C % 64 == 0, C <= 512")
        return _SelfAttention2dFn.apply(self, x, self.query_conv.weight, self.query_conv.bias, self.key_conv.weight,
                                        self.key_conv.bias, self.value_conv.weight, self.value_conv.bias, self.gamma)


class _SelfAttention2dFn(torch.autograd.Function):
    """SelfAttention2d forward / backward on the flash kernels. Gradients: x, the three 1x1 convs, gamma."""

    @staticmethod
    def forward(ctx, mod, x, wq, bq, wk, bk, wv_p, bv, gamma):
        B, C, H, W = x.shape
        N, dq = H * W, C // 8
        dev = x.device
        xf = x.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).view(B * N, C)

        def build_qk():
            w = torch.zeros(128, C, device=dev)
            w[:dq] = wq.detach().view(dq, C)
            w[64:64 + dq] = wk.detach().view(dq, C)
            return w.to(torch.bfloat16).contiguous()

        def build_qk_bias():
            b2 = torch.zeros(128, device=dev)
            b2[:dq] = bq.detach()
            b2[64:64 + dq] = bk.detach()
            return b2.contiguous()
        wqk = _cached((wq, wk), "sa_qk", build_qk)
        bqk = _cached((bq, bk), "sa_qk_bias", build_qk_bias)
        wv = _cached((wv_p,), "sa_v", lambda: wv_p.detach().view(C, C).to(torch.bfloat16).contiguous())
        qk = ops.linear_fprop(xf, wqk, bqk, out_f32=False)                       # [B*N, 128] = [q | k], zero padded
        v = ops.linear_fprop(xf, wv, _f32(bv), out_f32=False)                    # [B*N, C]
        y = torch.empty(B, N, C, device=dev, dtype=torch.bfloat16)
        need = any(ctx.needs_input_grad)
        o = torch.empty(B, N, C, device=dev, dtype=torch.bfloat16) if need else None
        lse = torch.empty(B * N, device=dev, dtype=torch.float32) if need else None
        gm = _f32(gamma)
        check(_capi.lib().lun_flash_attn2d_bf16(qk.data_ptr(), v.data_ptr(), xf.data_ptr(), y.data_ptr(), gm.data_ptr(),
                                                B, N, C, _p(o), _p(lse), _stream()), "lun_flash_attn2d_bf16")
        if need:
            ctx.save_for_backward(xf, qk, v, o, lse, wqk, wv, gm.clone())
            ctx.dims = (B, C, H, W)
        return y.view(B, H, W, C).permute(0, 3, 1, 2).float()

    @staticmethod
    def backward(ctx, dy):
        lib = _capi.lib()
        xf, qk, v, o, lse, wqk, wv, gm = ctx.saved_tensors
        B, C, H, W = ctx.dims
        N, dq = H * W, C // 8
        dev = dy.device
        dyf = dy.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).view(B * N, C)
        dsum = torch.empty(B * N, device=dev, dtype=torch.float32)
        dgamma = torch.zeros(1, device=dev, dtype=torch.float32)
        check(lib.lun_flash_attn2d_bwd_prep_bf16(dyf.data_ptr(), o.data_ptr(), dsum.data_ptr(), dgamma.data_ptr(),
                                                 B * N, C, _stream()), "lun_flash_attn2d_bwd_prep_bf16")
        dv = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
        check(lib.lun_flash_attn2d_dv_bf16(qk.data_ptr(), dyf.data_ptr(), lse.data_ptr(), gm.data_ptr(), dv.data_ptr(),
                                           B, N, C, _stream()), "lun_flash_attn2d_dv_bf16")
        dqk = torch.empty(B * N, 128, device=dev, dtype=torch.bfloat16)
        check(lib.lun_flash_attn2d_dqk_bf16(qk.data_ptr(), v.data_ptr(), dyf.data_ptr(), lse.data_ptr(),
                                            dsum.data_ptr(), gm.data_ptr(), dqk.data_ptr(), B, N, C, _stream()),
              "lun_flash_attn2d_dqk_bf16")
        # 1x1 convs: data gradients through the tcgen05 GEMM, weight gradients through the wgrad kernel
        dx = dy.detach().permute(0, 2, 3, 1).reshape(B * N, C).float()
        dx = dx + ops.linear_dgrad(dqk, wqk.t().contiguous()).float() + ops.linear_dgrad(dv, wv.t().contiguous()).float()
        dwqk = ops.linear_wgrad(dqk, xf)                                          # [128, C]
        dwv = ops.linear_wgrad(dv, xf)                                            # [C, C]
        dbqk = dqk.float().sum(0)
        grads = (None, dx.view(B, H, W, C).permute(0, 3, 1, 2),
                 dwqk[:dq].reshape(dq, C, 1, 1), dbqk[:dq], dwqk[64:64 + dq].reshape(dq, C, 1, 1), dbqk[64:64 + dq],
                 dwv.reshape(C, C, 1, 1), dv.float().sum(0), dgamma)
        return tuple(g if need else None for g, need in zip(grads, ctx.needs_input_grad))


class Encoder(nn.Module):
    """Parameter container mirroring lunar_generate.py:84-125."""

    def __init__(self, latent_dim=256):
        super().__init__()
        chans = [3, 64, 128, 256, 512]
        for i in range(4):
            setattr(self, f"down{i + 1}", nn.Sequential(*_conv_gn_mish(chans[i], chans[i + 1], stride=2),
                                                        ResBlock(chans[i + 1], chans[i + 1])))
        self.flatten = nn.Flatten()
        self.fc_mu = nn.Linear(512 * 8 * 8, latent_dim)
        self.fc_logvar = nn.Linear(512 * 8 * 8, latent_dim)

    def forward(self, x):
        _host.require_no_grad("Encoder.forward", x)
        mulv, skips, _ = _encoder_forward(self, x, save=False)
        L = self.fc_mu.out_features
        B = x.shape[0]
        sizes = ((64, 64), (128, 32), (256, 16))
        return mulv[:, :L], mulv[:, L:], [s.view(B, hw, hw, c).permute(0, 3, 1, 2).float()
                                          for s, (c, hw) in zip(skips, sizes)]


class Decoder(nn.Module):
    """Parameter container mirroring lunar_generate.py:155-192."""

    def __init__(self, latent_dim=256):
        super().__init__()
        self.fc = nn.Linear(latent_dim, 512 * 8 * 8)
        chans = [512, 256, 128, 64, 32]
        for i in range(4):
            setattr(self, f"up{i + 1}", nn.Sequential(
                nn.ConvTranspose2d(chans[i], chans[i + 1], kernel_size=4, stride=2, padding=1),
                nn.GroupNorm(8, chans[i + 1]), nn.Mish()))
        self.final_conv = nn.Conv2d(32, 3, kernel_size=3, padding=1)

    def forward(self, z, skips):
        _host.require_no_grad("Decoder.forward", z, *skips)
        B = z.shape[0]
        sk = [s.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).view(B, -1, s.shape[1]) for s in skips]
        recon, _ = _decoder_forward(self, z.detach().to(torch.bfloat16).contiguous(), sk, save=False)
        return recon


class LunarisCoreVAE(nn.Module):
    """VAE for 128x128 pixel art, drop-in for lunar_generate.py:231-291."""

    def __init__(self, latent_dim=256):
        super().__init__()
        self.latent_dim = latent_dim
        self.encoder = Encoder(latent_dim=latent_dim)
        self.decoder = Decoder(latent_dim=latent_dim)

    def reparameterize(self, mu, logvar):
        std = torch.exp(0.5 * logvar)
        eps = torch.randn_like(std)
        return mu + eps * std

    def forward(self, x):
        if not x.is_cuda:
            raise _capi.LunarisB200Error("lunaris_orion_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        B = x.shape[0]
        # first RNG draw of the step, same call and shape as the reference's randn_like (lunar_generate.py:260)
        eps = _host.draw_eps(B, self.latent_dim, x.device)
        params = list(self.parameters())
        return _VAEFn.apply(self, x.detach(), eps, *params)

    def sample(self, num_samples):
        dev = next(self.parameters()).device
        z = torch.randn(num_samples, self.latent_dim, device=dev)
        with torch.no_grad(), _host.zero_pool(dev):      # the per-image GroupNorm sums: one memset instead of five fills
            recon, _ = _decoder_forward(self.decoder, z.to(torch.bfloat16).contiguous(), [], save=False)
        return recon


# ====================================================================================================== plumbing
def _cached(key_params, kind, build):
    """Kernel-layout shadow of one or more parameters, cached on the first of them (see _host.cached)."""
    return _host.cached(key_params[0], kind, key_params, build)


def _f32(p):
    return _cached((p,), "f32", lambda: p.detach().float().contiguous())


def _wfwd(conv):
    return _cached((conv.weight,), "fwd", lambda: ops.pack_conv_weight(conv.weight))


def _wdgrad(conv):
    return _cached((conv.weight,), "dgrad", lambda: ops.pack_conv_weight_dgrad(conv.weight))


def _wT(convT):
    return _cached((convT.weight,), "Tfwd", lambda: ops.pack_convT_weight(convT.weight))


def _wTdgrad(convT):
    return _cached((convT.weight,), "Tdgrad", lambda: ops.pack_convT_weight_dgrad(convT.weight))


_stream = _capi.raw_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _nchw_cols_to_nhwc(w, C=512, S=64):
    """Linear weight over a flattened NCHW [C,8,8] feature (index c*64+s) -> same over NHWC (index s*512+c)."""
    n = w.shape[0]
    return w.view(n, C, S).permute(0, 2, 1).reshape(n, C * S)


def _conv_gn(B, hw, C, dev, conv_fn):
    """Run a conv whose output feeds a GroupNorm. conv_fn(img_stats) -> [B,hw,hw,C] bf16. Returns (t [B,HW,C], per-image
    channel sums [B,2,C]): accumulated by the conv epilogue itself when an image holds >= 128 output pixels (one
    image per tile), by a separate pass over t otherwise."""
    HW = hw * hw
    if ops.img_stats_ok(hw, hw):
        st = _host.zeros(B, 2, C, device=dev)
        return conv_fn(st).view(B, HW, C), st
    t = conv_fn(None).view(B, HW, C)
    return t, _gn_stats(t, B, HW, C)


def _gn_stats(t, B, HW, C):
    st = _host.zeros(B, 2, C, device=t.device)
    check(_capi.lib().lun_image_channel_stats_bf16(t.data_ptr(), st.data_ptr(), B, HW, C, _stream()),
          "lun_image_channel_stats_bf16")
    return st


def _gn_mish(t, st, gn, B, HW, C, res=None, add=None):
    y = torch.empty_like(t)
    check(_capi.lib().lun_gn_mish_fwd_bf16(t.data_ptr(), st.data_ptr(), _f32(gn.weight).data_ptr(),
                                           _f32(gn.bias).data_ptr(), _p(res), _p(add), y.data_ptr(), B, HW, C,
                                           gn.num_groups, gn.eps, _stream()), "lun_gn_mish_fwd_bf16")
    return y


def _gn_mish_bwd(dy, dy2, t, st, gn, B, HW, C, res=None, want_dres=False):
    """Returns (dt, dres, dgamma, dbeta)."""
    red = _host.zeros(B, 2, C, device=t.device)
    dt = torch.empty_like(t)
    dres = torch.empty_like(t) if want_dres else None
    check(_capi.lib().lun_gn_mish_bwd_bf16(dy.data_ptr(), _p(dy2), t.data_ptr(), st.data_ptr(),
                                           _f32(gn.weight).data_ptr(), _f32(gn.bias).data_ptr(), _p(res),
                                           red.data_ptr(), dt.data_ptr(), _p(dres), B, HW, C, gn.num_groups, gn.eps,
                                           _stream()), "lun_gn_mish_bwd_bf16")
    s = red.sum(0)
    return dt, dres, s[1], s[0]


def _colsum(t, P, C):
    """Per-channel sum of a bf16 [P,C] tensor (conv bias gradients)."""
    st = _host.zeros(2 * C, device=t.device)
    check(_capi.lib().lun_channel_stats_bf16(t.data_ptr(), P, C, st.data_ptr(), _stream()), "lun_channel_stats_bf16")
    return st[:C]


# ====================================================================================================== forward
def _resblock_forward(rb, a, B, hw, C, save):
    """ResBlock.forward (lunar_generate.py:47-53) on NHWC bf16 a [B,HW,Cin]; C = out channels. The 1x1 shortcut conv
    of the in != out case (never instantiated by the VAE, whose blocks have in == out) is forward-only."""
    HW = hw * hw
    c1, g1 = rb.conv1[0], rb.conv1[1]
    c2, g2 = rb.conv2[0], rb.conv2[1]
    cin = c1.in_channels
    if isinstance(rb.shortcut, nn.Identity):
        ident = a
    else:
        if save:
            raise _capi.LunarisB200Error("ResBlock with a 1x1 shortcut conv has no backward in lunaris_orion_b200")
        ident = ops.conv2d_fprop(a.view(B, hw, hw, cin), _wfwd(rb.shortcut), 1, 1, 0,
                                 bias=_f32(rb.shortcut.bias)).view(B, HW, C)
    t1, s1 = _conv_gn(B, hw, C, a.device, lambda st: ops.conv2d_fprop(a.view(B, hw, hw, cin), _wfwd(c1), 3, 1, 1,
                                                                      bias=_f32(c1.bias), img_stats=st))
    r1 = _gn_mish(t1, s1, g1, B, HW, C)
    t2, s2 = _conv_gn(B, hw, C, a.device, lambda st: ops.conv2d_fprop(r1.view(B, hw, hw, C), _wfwd(c2), 3, 1, 1,
                                                                      bias=_f32(c2.bias), img_stats=st))
    out = _gn_mish(t2, s2, g2, B, HW, C, res=ident)
    return out, (dict(a=a, t1=t1, s1=s1, r1=r1, t2=t2, s2=s2) if save else None)


def _encoder_forward(enc, x, save):
    """Encoder.forward (lunar_generate.py:127-153). x NCHW fp32. Returns (mulv [B,2L] fp32, skips, saved)."""
    lib = _capi.lib()
    B, _, H, W = x.shape
    x = x.detach().float().contiguous()
    chans = [64, 128, 256, 512]
    skips, saved = [], []
    h, hw = None, H
    for i in range(4):
        down = getattr(enc, f"down{i + 1}")
        conv, gn, rb = down[0], down[1], down[3]
        C = chans[i]
        hw //= 2
        HW = hw * hw
        if i == 0:
            t0 = torch.empty(B, HW, C, device=x.device, dtype=torch.bfloat16)
            check(lib.lun_conv3x3_c3_fwd(x.data_ptr(), _f32(conv.weight).data_ptr(), _f32(conv.bias).data_ptr(),
                                         t0.data_ptr(), B, H, W, C, 2, _stream()), "lun_conv3x3_c3_fwd")
            s0 = _gn_stats(t0, B, HW, C)
        else:
            t0, s0 = _conv_gn(B, hw, C, x.device, lambda st, h=h, i=i: ops.conv2d_fprop(
                h.view(B, hw * 2, hw * 2, chans[i - 1]), _wfwd(conv), 3, 2, 1, bias=_f32(conv.bias), img_stats=st))
        a = _gn_mish(t0, s0, gn, B, HW, C)
        out, rsv = _resblock_forward(rb, a, B, hw, C, save)
        if save:
            saved.append(dict(inp=h, t0=t0, s0=s0, rb=rsv, hw=hw, C=C))
        h = out
        if i < 3:
            skips.append(out)
    L = enc.fc_mu.out_features
    wmulv = _cached((enc.fc_mu.weight, enc.fc_logvar.weight), "mulv_fwd", lambda: _nchw_cols_to_nhwc(
        torch.cat([enc.fc_mu.weight.detach(), enc.fc_logvar.weight.detach()], 0)).to(torch.bfloat16).contiguous())
    bmulv = _cached((enc.fc_mu.bias, enc.fc_logvar.bias), "mulv_bias", lambda: torch.cat(
        [enc.fc_mu.bias.detach(), enc.fc_logvar.bias.detach()]).float().contiguous())
    flat = h.view(B, 64 * 512)
    mulv = ops.linear_fprop(flat, wmulv, bmulv, out_f32=True)
    return mulv, skips, (dict(stages=saved, flat=flat) if save else None)


def _decoder_forward(dec, z, skips, save):
    """Decoder.forward (lunar_generate.py:194-229). z bf16 [B,L]; skips: NHWC bf16 [B,HW,C] list (may be empty)."""
    lib = _capi.lib()
    B = z.shape[0]
    wfc = _cached((dec.fc.weight,), "fc_fwd", lambda: dec.fc.weight.detach().view(512, 64, -1).permute(1, 0, 2)
                  .reshape(512 * 64, -1).to(torch.bfloat16).contiguous())
    bfc = _cached((dec.fc.bias,), "fc_bias", lambda: dec.fc.bias.detach().view(512, 64).t().reshape(-1).float()
                  .contiguous())
    h = ops.linear_fprop(z, wfc, bfc, out_f32=False)           # [B, 64*512] in NHWC order
    chans = [512, 256, 128, 64, 32]
    hw = 8
    saved = []
    for i in range(4):
        up = getattr(dec, f"up{i + 1}")
        convT, gn = up[0], up[1]
        cin, C = chans[i], chans[i + 1]
        if ops.img_stats_ok(hw, hw):               # each output phase grid is hw x hw per image
            st = _host.zeros(B, 2, C, device=z.device)
            t = ops.convT4x4s2_fprop(h.view(B, hw, hw, cin), _wT(convT), bias=_f32(convT.bias), img_stats=st)
        else:
            st = None
            t = ops.convT4x4s2_fprop(h.view(B, hw, hw, cin), _wT(convT), bias=_f32(convT.bias))
        hw *= 2
        HW = hw * hw
        t = t.view(B, HW, C)
        if st is None:
            st = _gn_stats(t, B, HW, C)
        add = skips[2 - i] if (i < 3 and len(skips) >= 3 - i) else None
        if save:
            saved.append(dict(inp=h, t=t, st=st, hw=hw, C=C, cin=cin, has_skip=add is not None))
        if i == 3:
            break                                   # up4's GroupNorm + Mish is fused into the final conv below
        h = _gn_mish(t, st, gn, B, HW, C, add=add)
    fc = dec.final_conv
    recon = torch.empty(B, 3, hw, hw, device=z.device, dtype=torch.float32)
    h = torch.empty_like(t) if save else None       # the normalised 32-channel tensor exists only for the backward
    check(lib.lun_gn_mish_final_conv_tanh_fwd(t.data_ptr(), st.data_ptr(), _f32(gn.weight).data_ptr(),
                                              _f32(gn.bias).data_ptr(), _f32(fc.weight).data_ptr(),
                                              _f32(fc.bias).data_ptr(), _p(h), recon.data_ptr(), B, hw, hw,
                                              gn.num_groups, gn.eps, _stream()), "lun_gn_mish_final_conv_tanh_fwd")
    return recon, (dict(stages=saved, x_last=h, z=z) if save else None)


# ====================================================================================================== autograd
class _VAEFn(torch.autograd.Function):
    """(images, eps, *params) -> (recon, mu, logvar) with the full VAE backward on the CUDA kernels."""

    @staticmethod
    def forward(ctx, vae, x, eps, *params):
        with _host.zero_pool(x.device):
            return _VAEFn._forward(ctx, vae, x, eps)

    @staticmethod
    def _forward(ctx, vae, x, eps):
        lib = _capi.lib()
        B = x.shape[0]
        L = vae.latent_dim
        save = any(ctx.needs_input_grad)        # all False under no_grad: nothing is kept for a backward
        mulv, skips, esv = _encoder_forward(vae.encoder, x, save)
        z = torch.empty(B, L, device=x.device, dtype=torch.bfloat16)
        check(lib.lun_reparam_fwd(mulv.data_ptr(), eps.data_ptr(), z.data_ptr(), B, L, _stream()), "lun_reparam_fwd")
        recon, dsv = _decoder_forward(vae.decoder, z, skips, save)
        ctx.vae, ctx.esv, ctx.dsv = vae, esv, dsv
        if save:
            ctx.save_for_backward(x, eps, mulv, recon)     # recon is an output: saved through autograd, no cycle
        ctx.names = [n for n, _ in vae.named_parameters()]
        mu, logvar = mulv[:, :L], mulv[:, L:]
        return recon, mu, logvar

    @staticmethod
    def backward(ctx, d_recon, d_mu, d_logvar):
        with _host.zero_pool(ctx.saved_tensors[0].device):
            return _VAEFn._backward(ctx, d_recon, d_mu, d_logvar)

    @staticmethod
    def _backward(ctx, d_recon, d_mu, d_logvar):
        lib = _capi.lib()
        vae, esv, dsv = ctx.vae, ctx.esv, ctx.dsv
        enc, dec = vae.encoder, vae.decoder
        x, eps, mulv, recon = ctx.saved_tensors
        B, L = x.shape[0], vae.latent_dim
        dev = x.device
        g = {}

        # ---------------- decoder
        fc = dec.final_conv
        if d_recon is None:
            d_recon = torch.zeros_like(recon)
        d_recon = d_recon.float().contiguous()
        hw = 128
        dx = torch.empty(B, hw * hw, 32, device=dev, dtype=torch.bfloat16)
        dwf = torch.zeros_like(fc.weight, dtype=torch.float32)
        dbf = _host.zeros(3, device=dev)
        check(lib.lun_final_conv_bwd(d_recon.data_ptr(), recon.data_ptr(), dsv["x_last"].data_ptr(),
                                     _f32(fc.weight).data_ptr(), dx.data_ptr(), dwf.data_ptr(), dbf.data_ptr(), B, hw,
                                     hw, _stream()), "lun_final_conv_bwd")
        g["decoder.final_conv.weight"], g["decoder.final_conv.bias"] = dwf, dbf
        d_skips = [None, None, None]
        dh = dx
        for i in range(3, -1, -1):
            sv = dsv["stages"][i]
            up = getattr(dec, f"up{i + 1}")
            convT, gn = up[0], up[1]
            hw, C, cin = sv["hw"], sv["C"], sv["cin"]
            HW = hw * hw
            if sv["has_skip"]:
                d_skips[2 - i] = dh                      # y = gn_mish(t) + skip  ->  dskip = dy
            dt, _, dgam, dbet = _gn_mish_bwd(dh, None, sv["t"], sv["st"], gn, B, HW, C)
            g[f"decoder.up{i + 1}.1.weight"], g[f"decoder.up{i + 1}.1.bias"] = dgam, dbet
            g[f"decoder.up{i + 1}.0.bias"] = _colsum(dt, B * HW, C).clone()
            dt4 = dt.view(B, hw, hw, C)
            inp4 = sv["inp"].view(B, hw // 2, hw // 2, cin)
            g[f"decoder.up{i + 1}.0.weight"] = ops.convT4x4s2_wgrad(dt4, inp4).contiguous()
            dh = ops.convT4x4s2_dgrad(dt4, _wTdgrad(convT)).view(B, (hw // 2) ** 2, cin)
        # dh: gradient of the fc output [B, 64*512] (NHWC order)
        dh = dh.view(B, 64 * 512)
        dwfc = ops.linear_wgrad(dh, dsv["z"])                                     # [32768 (s*512+c), L]
        g["decoder.fc.weight"] = dwfc.view(64, 512, L).permute(1, 0, 2).reshape(512 * 64, L).contiguous()
        g["decoder.fc.bias"] = dh.float().sum(0).view(64, 512).t().reshape(-1).contiguous()
        wfc_t = _cached((dec.fc.weight,), "fc_dgrad", lambda: dec.fc.weight.detach().view(512, 64, -1)
                        .permute(2, 1, 0).reshape(L, 64 * 512).to(torch.bfloat16).contiguous())
        dz = ops.linear_dgrad(dh, wfc_t)                                          # [B, L] bf16

        # ---------------- reparameterisation + encoder heads
        dmulv = torch.empty(B, 2 * L, device=dev, dtype=torch.bfloat16)
        dmu = d_mu.float().contiguous() if d_mu is not None else None
        dlv = d_logvar.float().contiguous() if d_logvar is not None else None
        check(lib.lun_reparam_bwd(mulv.data_ptr(), eps.data_ptr(), dz.data_ptr(), _p(dmu), _p(dlv), dmulv.data_ptr(),
                                  B, L, _stream()), "lun_reparam_bwd")
        flat = esv["flat"]
        dw = ops.linear_wgrad(dmulv, flat)                                        # [2L, 32768 (s*512+c)]
        dw = dw.view(2 * L, 64, 512).permute(0, 2, 1).reshape(2 * L, 512 * 64)
        g["encoder.fc_mu.weight"], g["encoder.fc_logvar.weight"] = dw[:L].contiguous(), dw[L:].contiguous()
        db = dmulv.float().sum(0)
        g["encoder.fc_mu.bias"], g["encoder.fc_logvar.bias"] = db[:L].contiguous(), db[L:].contiguous()
        wmulv_t = _cached((enc.fc_mu.weight, enc.fc_logvar.weight), "mulv_dgrad", lambda: _nchw_cols_to_nhwc(
            torch.cat([enc.fc_mu.weight.detach(), enc.fc_logvar.weight.detach()], 0)).t().to(torch.bfloat16)
            .contiguous())
        d_out = ops.linear_dgrad(dmulv, wmulv_t).view(B, 64, 512)                 # gradient of down4's output

        # ---------------- encoder stages
        d_extra = None
        for i in range(3, -1, -1):
            sv = esv["stages"][i]
            down = getattr(enc, f"down{i + 1}")
            conv, gn, rb = down[0], down[1], down[3]
            hw, C = sv["hw"], sv["C"]
            HW = hw * hw
            r = sv["rb"]
            pre = f"encoder.down{i + 1}"
            # ResBlock tail: out = mish(mish(gn2(t2)) + a)
            dt2, d_a1, dgam, dbet = _gn_mish_bwd(d_out, d_extra, r["t2"], r["s2"], rb.conv2[1], B, HW, C, res=r["a"],
                                                 want_dres=True)
            g[pre + ".3.conv2.1.weight"], g[pre + ".3.conv2.1.bias"] = dgam, dbet
            g[pre + ".3.conv2.0.bias"] = _colsum(dt2, B * HW, C).clone()
            dt2_4 = dt2.view(B, hw, hw, C)
            g[pre + ".3.conv2.0.weight"] = ops.conv2d_wgrad(dt2_4, r["r1"].view(B, hw, hw, C), 3, 1, 1).contiguous()
            d_r1 = ops.conv2d_dgrad(dt2_4, _wdgrad(rb.conv2[0]), 3, 1, 1, (hw, hw)).view(B, HW, C)
            dt1, _, dgam, dbet = _gn_mish_bwd(d_r1, None, r["t1"], r["s1"], rb.conv1[1], B, HW, C)
            g[pre + ".3.conv1.1.weight"], g[pre + ".3.conv1.1.bias"] = dgam, dbet
            g[pre + ".3.conv1.0.bias"] = _colsum(dt1, B * HW, C).clone()
            dt1_4 = dt1.view(B, hw, hw, C)
            g[pre + ".3.conv1.0.weight"] = ops.conv2d_wgrad(dt1_4, r["a"].view(B, hw, hw, C), 3, 1, 1).contiguous()
            d_a2 = ops.conv2d_dgrad(dt1_4, _wdgrad(rb.conv1[0]), 3, 1, 1, (hw, hw)).view(B, HW, C)
            # stage head: a = mish(gn(t0)), t0 = conv_s2(prev)
            dt0, _, dgam, dbet = _gn_mish_bwd(d_a1, d_a2, sv["t0"], sv["s0"], gn, B, HW, C)
            g[pre + ".1.weight"], g[pre + ".1.bias"] = dgam, dbet
            if i == 0:
                dw0 = torch.zeros_like(conv.weight, dtype=torch.float32)
                db0 = _host.zeros(C, device=dev)
                check(lib.lun_conv3x3_c3_wgrad(dt0.data_ptr(), x.data_ptr(), dw0.data_ptr(), db0.data_ptr(), B,
                                               x.shape[2], x.shape[3], C, 2, _stream()), "lun_conv3x3_c3_wgrad")
                g[pre + ".0.weight"], g[pre + ".0.bias"] = dw0, db0
            else:
                cprev = sv["inp"].shape[-1]
                inp4 = sv["inp"].view(B, hw * 2, hw * 2, cprev)
                dt0_4 = dt0.view(B, hw, hw, C)
                g[pre + ".0.bias"] = _colsum(dt0, B * HW, C).clone()
                g[pre + ".0.weight"] = ops.conv2d_wgrad(dt0_4, inp4, 3, 2, 1).contiguous()
                d_out = ops.conv2d_dgrad(dt0_4, _wdgrad(conv), 3, 2, 1, (hw * 2, hw * 2)).view(B, 4 * HW, cprev)
                d_extra = d_skips[i - 1]                 # decoder's gradient into this stage's skip output
        ctx.esv = ctx.dsv = None
        grads = tuple(g[n].view_as(p) if g[n].shape != p.shape else g[n]
                      for n, p in zip(ctx.names, vae.parameters()))
        return (None, None, None) + grads


# ====================================================================================================== step losses
class _VaeLossFn(torch.autograd.Function):
    """(recon, images, mu, logvar) -> (recon_loss, kl_loss): MSE mean and -0.5*mean(1+lv-mu^2-e^lv)
    (train_hybrid.py:859-862) in one kernel, gradients in a second one."""

    @staticmethod
    def forward(ctx, recon, images, mu, logvar):
        lib = _capi.lib()
        B, L = mu.shape
        recon = recon.float().contiguous()
        images = images.float().contiguous()
        if (mu.dtype == torch.float32 and logvar.dtype == torch.float32 and mu.stride() == (2 * L, 1)
                and logvar.stride() == (2 * L, 1) and logvar.data_ptr() == mu.data_ptr() + 4 * L):
            mulv = torch.as_strided(mu, (B, 2 * L), (2 * L, 1))     # mu | logvar already packed by the fused fc
        else:
            mulv = torch.cat([mu.float(), logvar.float()], 1).contiguous()
        sums = _host.zeros(2, device=recon.device)
        check(lib.lun_vae_loss_fwd(recon.data_ptr(), images.data_ptr(), mulv.data_ptr(), sums.data_ptr(),
                                   recon.numel(), B, L, _stream()), "lun_vae_loss_fwd")
        ctx.save_for_backward(recon, images, mulv)
        return sums[0] * (1.0 / recon.numel()), sums[1] * (-0.5 / (B * L))     # python scalars: no H2D copy

    @staticmethod
    def backward(ctx, g_recon, g_kl):
        lib = _capi.lib()
        recon, images, mulv = ctx.saved_tensors
        B, L2 = mulv.shape
        L = L2 // 2
        # kl = -0.5/(BL) * sum(...): fold the -0.5 into the kernel's analytic form: d kl/d mu = mu/(BL), etc.
        g = torch.stack([g_recon if g_recon is not None else torch.zeros((), device=recon.device),
                         g_kl if g_kl is not None else torch.zeros((), device=recon.device)]).float().contiguous()
        d_recon = torch.empty_like(recon)
        d_mulv = torch.empty_like(mulv)
        check(lib.lun_vae_loss_bwd(recon.data_ptr(), images.data_ptr(), mulv.data_ptr(), g.data_ptr(),
                                   d_recon.data_ptr(), d_mulv.data_ptr(), recon.numel(), B, L, _stream()),
              "lun_vae_loss_bwd")
        return d_recon, None, d_mulv[:, :L], d_mulv[:, L:]


def vae_losses(recon, images, mu, logvar):
    """recon_loss (MSE mean) and kl_loss exactly as train_hybrid.py:859-862, through the fused CUDA kernels."""
    return _VaeLossFn.apply(recon, images, mu, logvar)


def sprites_to_tensor(u8_nhwc):
    """uint8 NHWC sprites on the GPU -> fp32 NCHW in [-1,1] (PixelArtDataset normalisation, train_hybrid.py:181-182)."""
    B, H, W, _ = u8_nhwc.shape
    out = torch.empty(B, 3, H, W, device=u8_nhwc.device, dtype=torch.float32)
    check(_capi.lib().lun_sprites_u8_to_f32(u8_nhwc.contiguous().data_ptr(), out.data_ptr(), B, H, W, _stream()),
          "lun_sprites_u8_to_f32")
    return out
