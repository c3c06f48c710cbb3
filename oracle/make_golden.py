"""Generate tests/golden/golden_small.pt by EXECUTING the unmodified reference (/root/reference) on CPU.

Run in the build container only:  python oracle/make_golden.py
The fixture pins (a) the init of the drop-in modules (same seed -> same state_dict), (b) the oracle restatement and
(c) the CUDA path on the GPU box, where /root/reference does not exist.

Contents: config + seeds, state_dict fingerprints, reference Teacher outputs (eval / train) with gradient
fingerprints and BN counters, reference VAE outputs / gradient fingerprints, and the 12 metrics + LR of one real
`TrainingManager._process_batch` (SURVEY.md App. C.1 harness; dropout probabilities set to 0 at run time so the
step is RNG-free apart from the VAE epsilon).
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_loader  # noqa: E402

CFG = dict(feat=64, emb=32, latent=64, B=2, seed=1234, img_seed=5)


def images(B, seed):
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (B, 3, 128, 128), generator=g, dtype=torch.uint8)
    return u8.float() / 127.5 - 1.0


def fingerprint(t):
    t = t.detach().double().flatten()
    idx = torch.linspace(0, t.numel() - 1, 8).long()
    return {"sum": t.sum().item(), "abs": t.abs().sum().item(), "n": t.numel(), "samples": t[idx].float().clone()}


def sd_fingerprint(m):
    return {k: fingerprint(v) for k, v in m.state_dict().items() if v.is_floating_point()}


def main():
    lg, le = reference_loader.load()
    out = {"cfg": CFG, "torch": torch.__version__}
    x = images(CFG["B"], CFG["img_seed"])

    # ---- modules
    torch.manual_seed(CFG["seed"])
    vae = lg.LunarisCoreVAE(latent_dim=CFG["latent"])
    teacher = le.LunarMoETeacher(feature_dim=CFG["feat"], embedding_dim=CFG["emb"], dropout_rate=0.0)
    out["vae_init"], out["teacher_init"] = sd_fingerprint(vae), sd_fingerprint(teacher)
    out["teacher_keys"], out["vae_keys"] = list(teacher.state_dict().keys()), list(vae.state_dict().keys())

    teacher.eval()
    with torch.no_grad():
        ev = teacher(x)
    out["teacher_eval"] = {k: ev[k].clone() for k in ("quality_scores", "expert_weights", "style_embedding",
                                                        "prompt_embedding", "semantic_score")}
    out["teacher_eval"]["feature_maps_fp"] = [fingerprint(f) for f in ev["feature_maps"]]
    teacher.train()
    tr = teacher(x)
    (-tr["quality_scores"].mean() * 0.5).backward()
    out["teacher_train"] = {k: tr[k].detach().clone() for k in ("quality_scores", "expert_weights", "semantic_score")}
    out["teacher_train_grads"] = {n: (None if p.grad is None else fingerprint(p.grad))
                                  for n, p in teacher.named_parameters()}
    out["teacher_train_nbt"] = {k: int(v) for k, v in teacher.state_dict().items() if k.endswith("num_batches_tracked")}
    out["teacher_keys_after_forward"] = list(teacher.state_dict().keys())

    vae.train()
    torch.manual_seed(77)
    recon, mu, lv = vae(x)
    loss = torch.nn.functional.mse_loss(recon, x) + 0.1 * (-0.5 * torch.mean(1 + lv - mu.pow(2) - lv.exp()))
    loss.backward()
    out["vae_train"] = {"recon_fp": fingerprint(recon), "mu": mu.detach().clone(), "logvar": lv.detach().clone(),
                        "loss": loss.item(), "eps_seed": 77}
    out["vae_train_grads"] = {n: fingerprint(p.grad) for n, p in vae.named_parameters()}

    # ---- one real TrainingManager._process_batch (App. C.1)
    sys.path.insert(0, reference_loader.REF)
    import train_hybrid as th
    _DL = th.DataLoader
    th.DataLoader = lambda ds, **kw: _DL(ds, **{**kw, "timeout": 0 if kw.get("num_workers", 0) == 0
                                                  else kw.get("timeout", 0)})
    cap = {}
    th.TrainingManager.train = lambda self: cap.__setitem__("tm", self)
    with tempfile.TemporaryDirectory() as d:
        data = os.path.join(d, "data")
        os.makedirs(data)
        np.save(os.path.join(data, "sprites_000.npy"),
                np.random.default_rng(1234).integers(0, 256, (10, 128, 128, 3), dtype=np.uint8))
        with open(os.path.join(data, "labels_000.csv"), "w") as f:
            f.write("filename,category,prompt,seed,pixel_size,guidance_scale,pag_scale,num_steps\n")
            for i in range(10):
                f.write(f"s{i}.png,cat,prompt,{i},8,7.5,3.0,20\n")
        argv = sys.argv
        sys.argv = ["train_hybrid.py", "--data_dir", data, "--output_dir", os.path.join(d, "out"), "--force_cpu",
                    "--batch_size", str(CFG["B"]), "--gradient_accumulation_steps", "1", "--num_workers", "0",
                    "--latent_dim", str(CFG["latent"]), "--embedding_dim", str(CFG["emb"]),
                    "--feature_dim", str(CFG["feat"]), "--seed", "42"]
        th.main()
        sys.argv = argv
        tm = cap["tm"]
        for m in tm.teacher.modules():
            if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)):
                m.p = 0.0
        out["trainer_init"] = {"vae": sd_fingerprint(tm.vae), "teacher": sd_fingerprint(tm.teacher)}
        torch.manual_seed(123)
        metrics = tm._process_batch(x.clone(), 0)
        out["trainer_step"] = {"metrics": metrics, "eps_seed": 123,
                               "vae_lr": tm.vae_optimizer.param_groups[0]["lr"],
                               "teacher_none": sorted(n for n, p in tm.teacher.named_parameters() if p.grad is None),
                               "teacher_nbt": {k: int(v) for k, v in tm.teacher.state_dict().items()
                                               if k.endswith("num_batches_tracked")}}
        m2 = tm._process_batch(x.clone(), 1)
        out["trainer_step2"] = {"metrics": m2, "vae_lr": tm.vae_optimizer.param_groups[0]["lr"]}
        tm._save_checkpoint()
        ck = torch.load(os.path.join(tm.checkpoints_dir, "latest.pt"), weights_only=True)
        out["checkpoint_keys"] = sorted(ck.keys())
        out["checkpoint_args_keys"] = sorted(ck["args"].keys())
    path = os.path.join(ROOT, "tests", "golden", "golden_small.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")
    print(out["trainer_step"]["metrics"])


if __name__ == "__main__":
    main()
