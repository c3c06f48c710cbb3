"""Checkpoints are interchangeable with the UNMODIFIED reference trainer in both directions (SURVEY.md 8b "State /
checkpoint layout", 4 "checkpoint round-trip through the reference's _load_checkpoint"). Runs on CPU and only where the
reference checkout exists (the build container): the reference `TrainingManager` takes one real step, saves
`latest.pt`; our `TrainingManager` (constructed on the CPU - no kernel runs) loads it; then the other way round through
the reference's own `_load_checkpoint` (train_hybrid.py:791-836, `weights_only=True`)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import reference_loader

pytestmark = pytest.mark.skipif(not reference_loader.available(), reason="reference checkout not present")

DIMS = dict(latent=64, emb=32, feat=64, B=2)


@pytest.fixture(scope="module")
def ref_trainer(tmp_path_factory):
    """The reference TrainingManager after one real _process_batch (SURVEY App. C.1 harness, CPU)."""
    d = tmp_path_factory.mktemp("refrun")
    reference_loader.load()
    sys.path.insert(0, reference_loader.REF)
    import train_hybrid as th
    _DL = th.DataLoader
    th.DataLoader = lambda ds, **kw: _DL(ds, **{**kw, "timeout": 0 if kw.get("num_workers", 0) == 0
                                                  else kw.get("timeout", 0)})
    cap = {}
    orig_train = th.TrainingManager.train
    th.TrainingManager.train = lambda self: cap.__setitem__("tm", self)
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
                "--batch_size", str(DIMS["B"]), "--gradient_accumulation_steps", "1", "--num_workers", "0",
                "--latent_dim", str(DIMS["latent"]), "--embedding_dim", str(DIMS["emb"]),
                "--feature_dim", str(DIMS["feat"]), "--seed", "42"]
    try:
        th.main()
    finally:
        sys.argv = argv
        th.TrainingManager.train = orig_train
        th.DataLoader = _DL
        sys.path.remove(reference_loader.REF)
    tm = cap["tm"]
    x = torch.rand(DIMS["B"], 3, 128, 128, generator=torch.Generator().manual_seed(3)) * 2 - 1
    tm._process_batch(x, 0)
    tm._save_checkpoint()
    return tm, str(tm.checkpoints_dir / "latest.pt"), str(d)


def _ours(tmp, extra=()):
    from lunaris_orion_b200.train_hybrid import TrainingManager, build_arg_parser
    args = build_arg_parser().parse_args([
        "--data_dir", "synthetic", "--output_dir", tmp, "--batch_size", str(DIMS["B"]),
        "--gradient_accumulation_steps", "1", "--latent_dim", str(DIMS["latent"]),
        "--embedding_dim", str(DIMS["emb"]), "--feature_dim", str(DIMS["feat"]), "--seed", "7", *extra])
    return TrainingManager(args, device=torch.device("cpu"))


def _same_state(loaded, source):
    """`loaded`: state_dict of a fresh module after load_state_dict(strict=False); `source`: the module that was saved.
    A module that has run a forward carries 12 lazily created `rel_pos_cache` buffers (lunar_evaluator.py:143,184) which
    a fresh module - the reference's own or ours - cannot receive (registered as None); everything else must match."""
    extra = [k for k in source if k not in loaded]
    assert all(k.endswith("rel_pos_cache") for k in extra), extra
    assert [k for k in source if k in loaded] == list(loaded.keys())
    for k in loaded:
        assert loaded[k].shape == source[k].shape, k
        if k.endswith("last_spatial_shapes"):              # zeros(2) float buffer that the forward replaces by int64 [H, W]
            assert torch.equal(loaded[k].double(), source[k].double()), k
        else:
            assert loaded[k].dtype == source[k].dtype and torch.equal(loaded[k], source[k]), k


def test_our_trainer_loads_a_reference_checkpoint(ref_trainer, tmp_path):
    ref_tm, path, _ = ref_trainer
    ours = _ours(str(tmp_path))                       # different seed: everything must come from the file
    ours._load_checkpoint(path)
    assert ours.global_step == ref_tm.global_step == 1
    assert ours.best_loss == ref_tm.best_loss
    _same_state(ours.vae.state_dict(), ref_tm.vae.state_dict())
    _same_state(ours.teacher.state_dict(), ref_tm.teacher.state_dict())       # 391 keys incl. rel_pos_cache buffers
    for mine, theirs in ((ours.vae_optimizer, ref_tm.vae_optimizer), (ours.teacher_optimizer, ref_tm.teacher_optimizer)):
        sa, sb = mine.state_dict(), theirs.state_dict()
        assert sa["state"].keys() == sb["state"].keys()           # same parameter indices (incl. the None-grad holes)
        for k in sa["state"]:
            for f in ("step", "exp_avg", "exp_avg_sq"):
                assert torch.equal(sa["state"][k][f], sb["state"][k][f]), (k, f)
        assert sa["param_groups"][0]["lr"] == sb["param_groups"][0]["lr"]
        for q, st in mine.state.items():                          # bias-correction step counters restored per tensor
            assert mine._host_step(q, st) == 1
    assert ours.vae_scheduler.state_dict()["last_epoch"] == ref_tm.vae_scheduler.state_dict()["last_epoch"]


def test_reference_trainer_loads_our_checkpoint(ref_trainer, tmp_path):
    ref_tm, path, _ = ref_trainer
    ours = _ours(str(tmp_path))
    ours._load_checkpoint(path)
    with torch.no_grad():                               # make our file distinguishable from the one it came from
        for p in ours.vae.parameters():
            p.add_(0.125)
    ours.global_step = 5
    mine = ours._save_checkpoint()
    ck = torch.load(mine, weights_only=True)            # the reference loads with weights_only=True
    assert sorted(ck.keys()) == sorted(torch.load(path, weights_only=True).keys())
    assert ref_tm._load_checkpoint(mine) is True        # train_hybrid.py:791-836 returns False on any exception
    assert ref_tm.global_step == 5
    _same_state(ours.vae.state_dict(), ref_tm.vae.state_dict())
    _same_state(ours.teacher.state_dict(), ref_tm.teacher.state_dict())


def test_unmodified_reference_trainer_builds_our_modules_through_the_import_seam(tmp_path):
    """the reference trainer binds exactly two names (train_hybrid.py:45-46). With those two names
    pointing at the drop-in classes, the UNMODIFIED TrainingManager constructs our modules (same constructor calls
    :394-404), builds its optimizers / schedulers over their parameters, runs enable_checkpointing's attribute probes
    (:407-424, which must find nothing), and saves a checkpoint that a reference trainer with the reference modules
    loads back completely (no missing / unexpected keys besides the lazily created buffers). CPU only: no forward."""
    reference_loader.load()
    sys.path.insert(0, reference_loader.REF)
    import train_hybrid as th
    from lunaris_orion_b200 import lunar_evaluator as ours_le, lunar_generate as ours_lg
    _DL = th.DataLoader
    th.DataLoader = lambda ds, **kw: _DL(ds, **{**kw, "timeout": 0 if kw.get("num_workers", 0) == 0
                                                  else kw.get("timeout", 0)})
    orig = (th.TrainingManager.train, th.LunarisCoreVAE, th.LunarMoETeacher)
    cap = {}
    th.TrainingManager.train = lambda self: cap.setdefault("tms", []).append(self)
    data = os.path.join(tmp_path, "data")
    os.makedirs(data)
    np.save(os.path.join(data, "sprites_000.npy"),
            np.random.default_rng(1234).integers(0, 256, (10, 128, 128, 3), dtype=np.uint8))
    with open(os.path.join(data, "labels_000.csv"), "w") as f:
        f.write("filename,category,prompt,seed,pixel_size,guidance_scale,pag_scale,num_steps\n")
        for i in range(10):
            f.write(f"s{i}.png,cat,prompt,{i},8,7.5,3.0,20\n")

    def run(outdir):
        argv = sys.argv
        sys.argv = ["train_hybrid.py", "--data_dir", data, "--output_dir", os.path.join(tmp_path, outdir),
                    "--force_cpu", "--batch_size", "2", "--gradient_accumulation_steps", "1", "--num_workers", "0",
                    "--latent_dim", str(DIMS["latent"]), "--embedding_dim", str(DIMS["emb"]),
                    "--feature_dim", str(DIMS["feat"]), "--seed", "42"]
        try:
            th.main()
        finally:
            sys.argv = argv
        return cap["tms"][-1]
    try:
        th.LunarisCoreVAE, th.LunarMoETeacher = ours_lg.LunarisCoreVAE, ours_le.LunarMoETeacher     # the seam
        tm_ours = run("ours")
        th.LunarisCoreVAE, th.LunarMoETeacher = orig[1], orig[2]
        tm_ref = run("ref")
    finally:
        th.TrainingManager.train, th.LunarisCoreVAE, th.LunarMoETeacher = orig
        th.DataLoader = _DL
        sys.path.remove(reference_loader.REF)
    assert type(tm_ours.vae) is ours_lg.LunarisCoreVAE and type(tm_ours.teacher) is ours_le.LunarMoETeacher
    # same seed, same construction order -> the drop-in modules start from the reference's exact weights
    _same_state(tm_ours.vae.state_dict(), tm_ref.vae.state_dict())
    _same_state(tm_ours.teacher.state_dict(), tm_ref.teacher.state_dict())
    assert len(tm_ours.vae_optimizer.param_groups[0]["params"]) == len(tm_ref.vae_optimizer.param_groups[0]["params"])
    tm_ours.global_step = 9
    tm_ours._save_checkpoint()
    assert tm_ref._load_checkpoint(str(tm_ours.checkpoints_dir / "latest.pt")) is True
    assert tm_ref.global_step == 9
