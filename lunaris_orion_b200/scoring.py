"""Generate-and-rank path: the sample-then-score loop the reference sketches in
examples/simple_generation.py:71-134, on the B200 kernels.

The example calls `vae.decode(z)` and `teacher.assess_quality(images)`, two methods the reference classes never define
(lunar_generate.py:232-291, lunar_evaluator.py:278-462), so it cannot run as shipped. Their evident meaning is used here:
decode = `Decoder.forward(z, skips=[])` (what `LunarisCoreVAE.sample` does after drawing z, lunar_generate.py:278-291) and
assess_quality = the Teacher's eval-mode `quality_scores` averaged over its 4 quality dimensions (the quantity the trainer
uses as the quality reward, train_hybrid.py:868-871). No new arithmetic: every kernel is one of the forward kernels.
"""
import torch

from . import lunar_generate as lg


def decode(vae, z):
    """Decoder.forward(z, skips=[]) (lunar_generate.py:194-229) -> images [n,3,128,128] fp32 in (-1, 1)."""
    with torch.no_grad():
        recon, _ = lg._decoder_forward(vae.decoder, z.to(torch.bfloat16).contiguous(), [], save=False)
    return recon


def assess_quality(teacher, images):
    """One score in (0, 1) per image: mean of the eval-mode `quality_scores` [n,4] (lunar_evaluator.py:455-462)."""
    was_training = teacher.training
    teacher.eval()
    try:
        with torch.no_grad():
            return teacher(images)["quality_scores"].mean(dim=1, keepdim=True)
    finally:
        teacher.train(was_training)


def generate_and_rank(vae, teacher, num_samples=4, temperature=1.0, quality_threshold=0.7, max_attempts=3, seed=None):
    """examples/simple_generation.py:71-134 `generate` without the (unused) prompt: draw latents, decode, score, keep
    the images whose score reaches `quality_threshold`, retry up to `max_attempts * num_samples` rounds.
    Returns (images [k,3,128,128], scores [k,1]) with k <= num_samples, best first."""
    if seed is not None:
        torch.manual_seed(seed)
    dev = next(vae.parameters()).device
    kept_img, kept_sc = [], []
    have, attempts = 0, 0
    while have < num_samples and attempts < max_attempts * num_samples:
        z = torch.randn(num_samples - have, vae.latent_dim, device=dev) * temperature
        images = decode(vae, z)
        scores = assess_quality(teacher, images)
        good = scores.squeeze(1) >= quality_threshold
        take = min(int(good.sum()), num_samples - have)          # the one host sync per round (data-dependent loop)
        if take:
            kept_img.append(images[good][:take])
            kept_sc.append(scores[good][:take])
            have += take
        attempts += 1
    if not kept_img:
        return (torch.empty(0, 3, 128, 128, device=dev), torch.empty(0, 1, device=dev))
    images, scores = torch.cat(kept_img), torch.cat(kept_sc)
    order = torch.argsort(scores.squeeze(1), descending=True)
    return images[order], scores[order]
