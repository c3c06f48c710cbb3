"""Drive the UNMODIFIED reference trainer (train_hybrid.TrainingManager) without its CLI loop - SURVEY.md App. C.1.

The bare CLI crashes on torch >= 2.11 with num_workers 0 (DataLoader timeout assertion; on CUDA also
prefetch_factor / persistent_workers without workers), so the harness shims those DataLoader kwargs, replaces `TrainingManager.train` by a capture, lets the reference's own `main()` parse
flags / seed / construct everything, and hands back the TrainingManager whose `_process_batch` the caller then drives.
Works from sources (/root/reference) and from oracle/_ref bytecode. TEST / BENCH INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np

from . import reference_loader

LABEL_HEADER = "filename,category,prompt,seed,pixel_size,guidance_scale,pag_scale,num_steps"


def write_sprites(data_dir, n, seed=1234):
    """SURVEY.md 8(d) synthetic data in the reference's on-disk format: sprites_000.npy + labels_000.csv."""
    os.makedirs(data_dir, exist_ok=True)
    np.save(os.path.join(data_dir, "sprites_000.npy"),
            np.random.default_rng(seed).integers(0, 256, (n, 128, 128, 3), dtype=np.uint8))
    with open(os.path.join(data_dir, "labels_000.csv"), "w") as f:
        f.write(LABEL_HEADER + "\n")
        for i in range(n):
            f.write(f"s{i}.png,cat,prompt,{i},8,7.5,3.0,20\n")


def import_reference_trainer(dropin=False):
    """`import train_hybrid` from the reference. dropin=True puts lunaris_orion_b200/dropin AHEAD of the reference on
    sys.path first, so the reference's `from lunar_generate import LunarisCoreVAE` /
    `from lunar_evaluator import LunarMoETeacher` (train_hybrid.py:45-46) bind the B200-native modules. Only valid in
    a process that has not imported those two module names yet."""
    if not reference_loader.available():
        raise RuntimeError("reference not available: run oracle/make_ref.py in the build container")
    if dropin:
        for name in ("lunar_generate", "lunar_evaluator", "train_hybrid"):
            if name in sys.modules:
                raise RuntimeError(f"{name} is already imported: use a fresh process for the drop-in seam")
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, os.path.join(root, "lunaris_orion_b200", "dropin"))
    # the reference directory is consulted last (sources and oracle/_ref bytecode alike), after everything on sys.path
    reference_loader.install_finder()
    import train_hybrid as th
    if not getattr(th, "_lun_dl_shim", False):
        _DL = th.DataLoader

        def loader(ds, **kw):
            # num_workers 0: torch >= 2.11 asserts timeout == 0 and rejects the prefetch_factor /
            # persistent_workers / multiprocessing_context the reference passes on CUDA (train_hybrid.py:561-570)
            if kw.get("num_workers", 0) == 0:
                kw = {**kw, "timeout": 0, "prefetch_factor": None, "persistent_workers": False,
                      "multiprocessing_context": None}
            return _DL(ds, **kw)
        th.DataLoader = loader
        th._lun_dl_shim = True
    return th


def drive_reference_trainer(cfg, data_dir, out_dir, device="cpu", dropin=False, extra=()):
    """cfg: dict(B, latent, emb, feat, seed[, accum, vae_lr, teacher_lr]). Returns the reference TrainingManager."""
    th = import_reference_trainer(dropin)
    cap = {}
    orig_train = th.TrainingManager.train
    th.TrainingManager.train = lambda self: cap.__setitem__("tm", self)
    argv = sys.argv
    sys.argv = ["train_hybrid.py", "--data_dir", data_dir, "--output_dir", out_dir,
                "--batch_size", str(cfg["B"]), "--gradient_accumulation_steps", str(cfg.get("accum", 1)),
                "--num_workers", "0", "--latent_dim", str(cfg["latent"]), "--embedding_dim", str(cfg["emb"]),
                "--feature_dim", str(cfg["feat"]), "--seed", str(cfg.get("seed", 42)),
                "--vae_lr", str(cfg.get("vae_lr", 1e-4)), "--teacher_lr", str(cfg.get("teacher_lr", 1e-4))]
    if device == "cpu":
        sys.argv.append("--force_cpu")
    sys.argv += list(extra)
    try:
        th.main()
    finally:
        sys.argv = argv
        th.TrainingManager.train = orig_train
    return cap["tm"]
