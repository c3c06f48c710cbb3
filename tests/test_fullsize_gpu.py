"""Size-independent properties at BASELINE.json's FULL sizes (the oracle is too slow there to be the checker):

* configs[4] ("C5"): decoder-only sampling, batch 256, latent 512 - the batch of 256 equals the same latents decoded
  in four batches of 64 (GroupNorm is per image, lunar_generate.py:170-222, so the decoder has no cross-image term),
  and its first images match the fp32 CPU oracle of the decoder.
* configs[2] ("C3"): one trainer step at batch 64, latent 512 / emb 256 / feat 512 - the fused loss kernel agrees with
  torch on the reconstruction the step produced, the loss identities of train_hybrid.py:886-896 hold, the executed
  gradient set is the reference's (168 of 268 Teacher tensors stay None, all 72 VAE tensors get one), gradients are
  finite and clipped to max_grad_norm, BatchNorm counters advance by 2 (Teacher passes A and B).
"""
import math

import pytest
import torch
import torch.nn.functional as F

import teacher_cases as tc
from oracle import restatement as R


@pytest.mark.gpu
def test_c5_sampling_batch_256_is_batch_split_invariant_and_matches_the_oracle(cuda_dev):
    from lunaris_orion_b200 import lunar_generate as lg
    torch.manual_seed(3)
    vae = lg.LunarisCoreVAE(latent_dim=512).to(cuda_dev).eval()
    z = torch.randn(256, 512, generator=torch.Generator().manual_seed(42)).to(cuda_dev)
    with torch.no_grad():
        full = vae.decoder(z, [])
        parts = torch.cat([vae.decoder(z[i:i + 64], []) for i in range(0, 256, 64)])
    assert full.shape == (256, 3, 128, 128) and torch.isfinite(full).all()
    assert full.abs().max() <= 1.0                                    # tanh range
    # same kernels, same per-image arithmetic; only the order of the fp32 atomics of the GroupNorm sums may differ
    assert (full - parts).abs().max().item() <= 5e-2              # a flipped bf16 rounding upstream, at worst
    assert (full - parts).abs().mean().item() <= 1e-3
    ref = R.decoder_forward(z[:4].cpu(), [], {k: v.detach() for k, v in tc.oracle_sd(vae).items()})
    assert tc.rel_err(full[:4], ref) < 0.05                          # tolerance of test_vae_gpu.py's sampling check
    # sample() is the same path on its own latent draw (lunar_generate.py:278-291)
    torch.manual_seed(9)
    s = vae.sample(256)
    torch.manual_seed(9)
    z2 = torch.randn(256, 512, device=cuda_dev)
    with torch.no_grad():
        d = (s - vae.decoder(z2, [])).abs()
    assert s.shape == (256, 3, 128, 128) and d.max().item() <= 5e-2 and d.mean().item() <= 1e-3


@pytest.mark.gpu
def test_c3_trainer_step_at_batch_64_keeps_the_reference_invariants(cuda_dev, tmp_path):
    from lunaris_orion_b200.train_hybrid import TrainingManager, build_arg_parser
    args = build_arg_parser().parse_args([
        "--data_dir", "synthetic", "--output_dir", str(tmp_path), "--batch_size", "64",
        "--gradient_accumulation_steps", "1", "--latent_dim", "512", "--embedding_dim", "256", "--feature_dim", "512",
        "--seed", "42"])
    tm = TrainingManager(args, device=cuda_dev)
    x = tc.images(64, 77).to(cuda_dev)
    nbt0 = {k: int(v) for k, v in tm.teacher.state_dict().items() if k.endswith("num_batches_tracked")}
    w0 = tm.vae.decoder.final_conv.weight.detach().clone()
    m = tm._process_batch(x, 0)
    assert all(math.isfinite(v) for v in m.values()), m
    # fused MSE kernel vs torch on the reconstruction this step produced
    assert abs(m["recon_loss"] - F.mse_loss(tm._last_recon.float(), x).item()) <= 2e-5 * m["recon_loss"] + 1e-7
    assert abs(m["vae_loss"] - (m["recon_loss"] + 0.1 * m["kl_loss"] + m["pg_loss"])) < 1e-4
    assert abs(m["teacher_loss"] - 0.5 * m["quality_loss"]) < 1e-6
    assert abs(m["quality_loss"] + m["quality_scores"]) < 1e-6
    assert 0.0 < m["quality_scores"] < 1.0 and m["kl_loss"] >= 0.0
    assert abs(m["advantage"]) < 1e-6 and abs(m["pg_loss"]) < 1e-6      # first step: baseline == reward
    named = list(tm.teacher.named_parameters())
    assert len(named) == 268 and sum(p.grad is None for _, p in named) == 168
    vg = [p.grad for p in tm.vae.parameters()]
    assert len(vg) == 72 and all(g is not None and torch.isfinite(g).all() for g in vg)
    assert all(torch.isfinite(p.grad).all() for _, p in named if p.grad is not None)
    # gradients are left clipped in place (train_hybrid.py:913-915)
    for model in (tm.vae, tm.teacher):
        norm = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters() if p.grad is not None)).item()
        assert norm <= args.max_grad_norm * (1 + 1e-4), norm
    nbt1 = {k: int(v) for k, v in tm.teacher.state_dict().items() if k.endswith("num_batches_tracked")}
    steps = {nbt1[k] - nbt0[k] for k in nbt0}
    assert steps <= {2, 3} and 2 in steps, steps                       # +2 / +2 / +3 per step
    assert not torch.equal(w0, tm.vae.decoder.final_conv.weight.detach())
