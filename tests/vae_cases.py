"""VAE parity helpers: lunaris_orion_b200.lunar_generate (CUDA) vs oracle/restatement.py (fp32 CPU; bf16-autocast GPU
run of the same op sequence as the noise calibration)."""
import contextlib

import torch
import torch.nn.functional as F

from oracle import restatement as R
from teacher_cases import images, oracle_sd, rel_err


def make_vae(dev, latent=64, seed=21):
    from lunaris_orion_b200 import lunar_generate as lg
    torch.manual_seed(seed)
    return lg.LunarisCoreVAE(latent_dim=latent).to(dev)


def _loss(recon, x, mu, lv):
    return F.mse_loss(recon.float(), x) + 0.1 * (-0.5 * torch.mean(1 + lv.float() - mu.float().pow(2) - lv.float().exp()))


def _oracle(x, sd, eps, autocast=False):
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
    with ctx:
        recon, mu, lv = R.vae_forward(x, sd, eps)
        loss = _loss(recon, x, mu, lv)
    loss.backward()
    return recon.detach().float(), mu.detach().float(), lv.detach().float()


def vae_report(dev, B=4, latent=64, calibrate=True):
    vae = make_vae(dev, latent).train()
    x = images(B, seed=31)
    torch.manual_seed(77)
    recon, mu, lv = vae(x.to(dev))
    torch.manual_seed(77)
    eps = torch.randn(B, latent, device=dev).cpu()
    _loss(recon, x.to(dev), mu, lv).backward()
    sd = oracle_sd(vae)
    r_recon, r_mu, r_lv = _oracle(x, sd, eps)
    rep = {"recon": rel_err(recon, r_recon), "mu": rel_err(mu, r_mu), "logvar": rel_err(lv, r_lv)}
    names = [n for n, _ in vae.named_parameters()]
    rep["all_grads_present"] = all(p.grad is not None for p in vae.parameters())

    def errs(get):
        return {n: (get(n) - sd[n].grad).abs().max().item() / (sd[n].grad.abs().max().item() + 1e-20) for n in names}
    mine = errs(lambda n: dict(vae.named_parameters())[n].grad.detach().cpu().float())
    rep["grad_rel_max"] = max(mine.values())
    rep["grad_worst"] = sorted(mine.items(), key=lambda kv: -kv[1])[:6]
    if calibrate:
        sdc = {k: v.detach().to(dev) for k, v in oracle_sd(vae).items()}
        for n in names:
            sdc[n].requires_grad_(True)
        c_recon, c_mu, c_lv = _oracle(x.to(dev), sdc, eps.to(dev), autocast=True)
        cal = errs(lambda n: sdc[n].grad.detach().cpu().float())
        rep["cal_recon"] = rel_err(c_recon, r_recon)
        rep["cal_mu"] = rel_err(c_mu, r_mu)
        rep["cal_grad_rel_max"] = max(cal.values())
        rep["cal_grad_worst"] = sorted(cal.items(), key=lambda kv: -kv[1])[:6]
        rep["ratio_worst"] = sorted(((n, mine[n] / (cal[n] + 1e-3)) for n in names), key=lambda kv: -kv[1])[:6]
    # decoder-only sampling path
    torch.manual_seed(5)
    s = vae.sample(3)
    torch.manual_seed(5)
    z = torch.randn(3, latent, device=dev).cpu()
    with torch.no_grad():
        rep["sample"] = rel_err(s, R.decoder_forward(z, [], sd))
    return rep
