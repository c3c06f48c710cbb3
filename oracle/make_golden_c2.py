"""Generate tests/golden/golden_c2.pt: the UNMODIFIED reference trainer at BASELINE.json configs[1] ("C2").

    train_hybrid.py --batch_size 16 --gradient_accumulation_steps 4 --latent_dim 256 --embedding_dim 128
                    --feature_dim 256            (here with --force_cpu: fp32 on the host cores)

Run in the build container only (needs /root/reference; about a quarter of an hour on 8 host threads):
    python oracle/make_golden_c2.py
Five real `TrainingManager._process_batch` calls (SURVEY.md App. C.1 harness): micro-batches 0-3 are one accumulation
window (four DIFFERENT batches of 16 seeded sprites, optimizer step on the fourth, `train_hybrid.py:906-926`), the fifth
call runs on the updated weights. Dropout probabilities are set to 0 at run time (torch's Philox stream cannot be matched
by a fused kernel); epsilon draw, reward baseline, 1/accum scaling, per-micro-batch zero_grad, clip,
AdamW and the cosine step are the reference's. The fixture keeps the 12 metrics of every call, learning rates, the
BatchNorm counters, and fingerprints (sum, |sum|, 8 samples) of a parameter set before / after the window, of every
gradient left on the parameters at the boundary and of every parameter after the optimizer step.
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
from oracle.make_golden import fingerprint, images  # noqa: E402

CFG = dict(feat=256, emb=128, latent=256, B=16, accum=4, seed=42, img_seed=31, eps_seed=321, calls=5)


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
    out = {"cfg": CFG, "torch": torch.__version__, "calls": []}
    with tempfile.TemporaryDirectory() as d:
        data = os.path.join(d, "data")
        os.makedirs(data)
        n = 20                                     # >= ceil(B / 0.9) sprites so the 90/10 split holds a full batch
        np.save(os.path.join(data, "sprites_000.npy"),
                np.random.default_rng(1234).integers(0, 256, (n, 128, 128, 3), dtype=np.uint8))
        with open(os.path.join(data, "labels_000.csv"), "w") as f:
            f.write("filename,category,prompt,seed,pixel_size,guidance_scale,pag_scale,num_steps\n")
            for i in range(n):
                f.write(f"s{i}.png,cat,prompt,{i},8,7.5,3.0,20\n")
        argv = sys.argv
        sys.argv = ["train_hybrid.py", "--data_dir", data, "--output_dir", os.path.join(d, "out"), "--force_cpu",
                    "--batch_size", str(CFG["B"]), "--gradient_accumulation_steps", str(CFG["accum"]),
                    "--num_workers", "0", "--latent_dim", str(CFG["latent"]), "--embedding_dim", str(CFG["emb"]),
                    "--feature_dim", str(CFG["feat"]), "--seed", str(CFG["seed"])]
        th.main()
        sys.argv = argv
        tm = cap["tm"]
        for m in tm.teacher.modules():
            if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)):
                m.p = 0.0
        torch.manual_seed(CFG["eps_seed"])
        t0 = time.time()
        for i in range(CFG["calls"]):
            x = images(CFG["B"], CFG["img_seed"] + i)
            metrics = tm._process_batch(x, i)
            rec = {"metrics": metrics, "vae_lr": tm.vae_optimizer.param_groups[0]["lr"],
                   "teacher_lr": tm.teacher_optimizer.param_groups[0]["lr"], "global_step": int(tm.global_step),
                   "probe_weight": fingerprint(tm.vae.decoder.final_conv.weight)}
            if i == CFG["accum"] - 1:              # the accumulation boundary: clip + AdamW + scheduler ran
                rec.update({
                    "teacher_none": sorted(k for k, p in tm.teacher.named_parameters() if p.grad is None),
                    "teacher_nbt": {k: int(v) for k, v in tm.teacher.state_dict().items()
                                    if k.endswith("num_batches_tracked")},
                    "vae_grads": {k: fingerprint(p.grad) for k, p in tm.vae.named_parameters()},
                    "teacher_grads": {k: fingerprint(p.grad) for k, p in tm.teacher.named_parameters()
                                      if p.grad is not None},
                    "vae_params_after": {k: fingerprint(p) for k, p in tm.vae.named_parameters()},
                    "teacher_params_after": {k: fingerprint(p) for k, p in tm.teacher.named_parameters()
                                             if p.grad is not None},
                })
            out["calls"].append(rec)
            print(i, "%.0f s" % (time.time() - t0), metrics, flush=True)
        out["seconds"] = time.time() - t0
    path = os.path.join(ROOT, "tests", "golden", "golden_c2.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
