"""Generate tests/golden/golden_c1.pt: the UNMODIFIED reference trainer at BASELINE.json configs[0] ("C1").

    train_hybrid.py --force_cpu --batch_size 8 --latent_dim 256 --embedding_dim 128 --feature_dim 256

Run in the build container only (needs /root/reference; about five minutes on 8 host threads):
    python oracle/make_golden_c1.py
Three real `TrainingManager._process_batch` calls (SURVEY.md App. C.1 harness) on 8 seeded sprites, dropout
probabilities set to 0 at run time (torch's Philox stream cannot be matched by a fused kernel; everything else -
epsilon draw, reward baseline, clip, AdamW, cosine step - is the reference's). The fixture keeps the 12 metrics of all
three steps, the learning rates, the grad-None set, BatchNorm counters and a fingerprint (sum, |sum|, 8 samples) of every
gradient of step 0 and of every parameter after the optimizer step. Test infrastructure only.
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
from oracle.make_golden import fingerprint, images  # noqa: E402

CFG = dict(feat=256, emb=128, latent=256, B=8, seed=42, img_seed=11, eps_seed=123)


def main():
    reference_loader.load()
    sys.path.insert(0, reference_loader.REF)
    import train_hybrid as th
    _DL = th.DataLoader
    th.DataLoader = lambda ds, **kw: _DL(ds, **{**kw, "timeout": 0 if kw.get("num_workers", 0) == 0
                                                  else kw.get("timeout", 0)})
    cap = {}
    th.TrainingManager.train = lambda self: cap.__setitem__("tm", self)
    torch.set_num_threads(os.cpu_count() or 1)
    out = {"cfg": CFG, "torch": torch.__version__}
    x = images(CFG["B"], CFG["img_seed"])
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
                    "--feature_dim", str(CFG["feat"]), "--seed", str(CFG["seed"])]
        th.main()
        sys.argv = argv
        tm = cap["tm"]
        for m in tm.teacher.modules():
            if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)):
                m.p = 0.0
        torch.manual_seed(CFG["eps_seed"])
        t0 = time.time()
        metrics = tm._process_batch(x.clone(), 0)
        out["seconds_step0"] = time.time() - t0
        out["step0"] = {
            "metrics": metrics,
            "vae_lr": tm.vae_optimizer.param_groups[0]["lr"], "teacher_lr": tm.teacher_optimizer.param_groups[0]["lr"],
            "teacher_none": sorted(n for n, p in tm.teacher.named_parameters() if p.grad is None),
            "teacher_nbt": {k: int(v) for k, v in tm.teacher.state_dict().items() if k.endswith("num_batches_tracked")},
            # gradients as left on the parameters after the step (clipped in place by clip_grad_norm_)
            "vae_grads": {n: fingerprint(p.grad) for n, p in tm.vae.named_parameters()},
            "teacher_grads": {n: fingerprint(p.grad) for n, p in tm.teacher.named_parameters() if p.grad is not None},
            "vae_params_after": {n: fingerprint(p) for n, p in tm.vae.named_parameters()},
            "teacher_params_after": {n: fingerprint(p) for n, p in tm.teacher.named_parameters() if p.grad is not None},
        }
        m2 = tm._process_batch(x.clone(), 1)
        out["step1"] = {"metrics": m2, "vae_lr": tm.vae_optimizer.param_groups[0]["lr"],
                        "teacher_lr": tm.teacher_optimizer.param_groups[0]["lr"]}
        m3 = tm._process_batch(x.clone(), 2)
        out["step2"] = {"metrics": m3, "vae_lr": tm.vae_optimizer.param_groups[0]["lr"],
                        "teacher_lr": tm.teacher_optimizer.param_groups[0]["lr"],
                        "teacher_nbt": {k: int(v) for k, v in tm.teacher.state_dict().items()
                                        if k.endswith("num_batches_tracked")}}
    path = os.path.join(ROOT, "tests", "golden", "golden_c1.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; step 0 took %.1f s" % out["seconds_step0"])
    print(out["step0"]["metrics"])
    print(out["step1"]["metrics"])
    print(out["step2"]["metrics"])


if __name__ == "__main__":
    main()
