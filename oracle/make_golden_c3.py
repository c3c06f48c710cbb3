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
from oracle import reference_loader  # noqa: E402
from oracle.make_golden import images  # noqa: E402

CFG = dict(feat=512, emb=256, latent=512, B=4, seed=42, img_seed=17, eps_seed=231, vae_lr=3e-4, teacher_lr=2e-4,
           steps=2, samples=64)


def fingerprint(t, k=64):
    """sum, |sum| and k evenly spaced samples (tests/teacher_cases.py::fingerprint with a wider sample)."""
    t = t.detach().double().flatten()
    idx = torch.linspace(0, t.numel() - 1, min(k, t.numel())).long()
    return {"sum": t.sum().item(), "abs": t.abs().sum().item(), "n": t.numel(), "samples": t[idx].float().clone()}


def drive_reference_trainer(cfg, data_dir, out_dir, extra=()):
    """build the unmodified reference TrainingManager through its own main()."""
    reference_loader.load()
    if reference_loader.REF not in sys.path:
        sys.path.insert(0, reference_loader.REF)
    import train_hybrid as th
    if not getattr(th, "_lun_dl_shim", False):
        _DL = th.DataLoader
        th.DataLoader = lambda ds, **kw: _DL(ds, **{**kw, "timeout": 0 if kw.get("num_workers", 0) == 0
                                                      else kw.get("timeout", 0)})
        th._lun_dl_shim = True
    cap = {}
    th.TrainingManager.train = lambda self: cap.__setitem__("tm", self)
    argv = sys.argv
    sys.argv = ["train_hybrid.py", "--data_dir", data_dir, "--output_dir", out_dir, "--force_cpu",
                "--batch_size", str(cfg["B"]), "--gradient_accumulation_steps", str(cfg.get("accum", 1)),
                "--num_workers", "0", "--latent_dim", str(cfg["latent"]), "--embedding_dim", str(cfg["emb"]),
                "--feature_dim", str(cfg["feat"]), "--seed", str(cfg["seed"]),
                "--vae_lr", str(cfg.get("vae_lr", 1e-4)), "--teacher_lr", str(cfg.get("teacher_lr", 1e-4))] + list(extra)
    try:
        th.main()
    finally:
        sys.argv = argv
    return cap["tm"]


def write_sprites(data, n):
    os.makedirs(data, exist_ok=True)
    np.save(os.path.join(data, "sprites_000.npy"),
            np.random.default_rng(1234).integers(0, 256, (n, 128, 128, 3), dtype=np.uint8))
    with open(os.path.join(data, "labels_000.csv"), "w") as f:
        f.write("filename,category,prompt,seed,pixel_size,guidance_scale,pag_scale,num_steps\n")
        for i in range(n):
            f.write(f"s{i}.png,cat,prompt,{i},8,7.5,3.0,20\n")


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
