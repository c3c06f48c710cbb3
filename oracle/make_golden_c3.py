"""Generate tests/golden/golden_c3.pt: the UNMODIFIED reference trainer at the BENCHMARKED architecture (BASELINE.json
configs[2], "C3": latent 512 / emb 256 / feat 512, head_dim 64) on a bounded batch.

    train_hybrid.py --force_cpu --batch_size 4 --latent_dim 512 --embedding_dim 256 --feature_dim 512
                    --vae_lr 3e-4 --teacher_lr 2e-4

Run in the build container only (needs /root/reference; about ten minutes on 8 host threads):
    python oracle/make_golden_c3.py
Two real `TrainingManager._process_batch` calls (SURVEY.md App. C.1 harness) on 4 seeded sprites, dropout
probabilities set to 0 at run time. Besides what golden_c1.pt keeps (12 metrics per step, learning rates, grad-None
set, BatchNorm counters) the fixture stores
  * 64-sample fingerprints of every gradient of step 0 and of every updated parameter (value / sign agreement is
    checked on the samples, tests/test_c3_gpu.py),
  * the pass-B Teacher outputs per sample (quality_scores, semantic_score, expert_weights) of both steps, so the
    ill-conditioned sigmoid heads are compared as logits,
  * fingerprints of the reconstruction.
Test infrastructure only.
"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import images  # noqa: E402
from oracle.ref_harness import drive_reference_trainer, write_sprites  # noqa: E402

CFG = dict(feat=512, emb=256, latent=512, B=4, seed=42, img_seed=17, eps_seed=231, vae_lr=3e-4, teacher_lr=2e-4,
           steps=2, samples=64)


def fingerprint(t, k=64):
    """sum, |sum| and k evenly spaced samples (tests/teacher_cases.py::fingerprint with a wider sample)."""
    t = t.detach().double().flatten()
    idx = torch.linspace(0, t.numel() - 1, min(k, t.numel())).long()
    return {"sum": t.sum().item(), "abs": t.abs().sum().item(), "n": t.numel(), "samples": t[idx].float().clone()}


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    out = {"cfg": CFG, "torch": torch.__version__, "steps": []}
    x = images(CFG["B"], CFG["img_seed"])
    k = CFG["samples"]
    with tempfile.TemporaryDirectory() as d:
        data = os.path.join(d, "data")
        write_sprites(data, 10)
        tm = drive_reference_trainer(CFG, data, os.path.join(d, "out"))
        for m in tm.teacher.modules():
            if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)):
                m.p = 0.0
        calls = []
        hook = tm.teacher.register_forward_hook(lambda mod, inp, o: calls.append(
            {kk: o[kk].detach().clone() for kk in ("quality_scores", "semantic_score", "expert_weights")}))
        recons = []
        hook2 = tm.vae.register_forward_hook(lambda mod, inp, o: recons.append(o[0].detach().clone()))
        torch.manual_seed(CFG["eps_seed"])
        for s in range(CFG["steps"]):
            calls.clear()
            recons.clear()
            t0 = time.time()
            metrics = tm._process_batch(x.clone(), s)
            rec = {"metrics": metrics, "seconds": time.time() - t0,
                   "vae_lr": tm.vae_optimizer.param_groups[0]["lr"],
                   "teacher_lr": tm.teacher_optimizer.param_groups[0]["lr"],
                   "pass_b": calls[1], "pass_a": calls[0], "recon_fp": fingerprint(recons[0], k),
                   "teacher_nbt": {kk: int(v) for kk, v in tm.teacher.state_dict().items()
                                   if kk.endswith("num_batches_tracked")}}
            if s == 0:
                rec.update({
                    "teacher_none": sorted(n for n, p in tm.teacher.named_parameters() if p.grad is None),
                    "vae_grads": {n: fingerprint(p.grad, k) for n, p in tm.vae.named_parameters()},
                    "teacher_grads": {n: fingerprint(p.grad, k) for n, p in tm.teacher.named_parameters()
                                      if p.grad is not None},
                    "vae_params_after": {n: fingerprint(p, k) for n, p in tm.vae.named_parameters()},
                    "teacher_params_after": {n: fingerprint(p, k) for n, p in tm.teacher.named_parameters()
                                             if p.grad is not None},
                })
            out["steps"].append(rec)
            print(s, "%.0f s" % rec["seconds"], metrics, flush=True)
        hook.remove()
        hook2.remove()
    path = os.path.join(ROOT, "tests", "golden", "golden_c3.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
