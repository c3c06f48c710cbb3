"""The UNMODIFIED reference trainer runs real steps on the drop-in modules on the B200: a fresh
process puts lunaris_orion_b200/dropin ahead of the reference (oracle/_ref bytecode on the GPU box) on sys.path, the
reference's own main() builds TrainingManager, and its `_process_batch` - torch AdamW, clip_grad_norm_, python-float
baseline, F.mse_loss, `.item()` metrics, `images.requires_grad_()` and all (train_hybrid.py:838-954) - drives our
kernels. With dropout off its metrics must agree with lunaris_orion_b200's own trainer on the same seeds, and with the
golden metrics the reference trainer produced with the reference modules."""
import json
import os
import subprocess
import sys

import pytest
import torch

import teacher_cases as tc
from oracle import reference_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gold = torch.load(os.path.join(ROOT, "tests", "golden", "golden_small.pt"), weights_only=False)
CFG = gold["cfg"]


@pytest.mark.gpu
@pytest.mark.skipif(not reference_loader.available(), reason="oracle/_ref not built (run oracle/make_ref.py)")
def test_unmodified_reference_trainer_steps_on_the_dropin_modules(cuda_dev, tmp_path):
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_reference_on_dropin.py"), "--steps", "2", "--dropout", "0",
           "--batch", str(CFG["B"]), "--latent", str(CFG["latent"]), "--emb", str(CFG["emb"]), "--feat", str(CFG["feat"]),
           "--img-seed", str(CFG["img_seed"]), "--eps-seed", str(gold["trainer_step"]["eps_seed"]), "--cpu-eps"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    rep = json.loads(out.stdout.strip().splitlines()[-1])
    # the reference's trainer, our modules, through the two-name import seam
    assert rep["vae_class"] == "lunaris_orion_b200.lunar_generate.LunarisCoreVAE"
    assert rep["teacher_class"] == "lunaris_orion_b200.lunar_evaluator.LunarMoETeacher"
    assert all(os.path.join("lunaris_orion_b200", "dropin") in f for f in rep["shim_modules"])
    assert "lunaris_orion_b200" not in rep["trainer_module"] and rep["optimizer_class"] == "AdamW"
    assert rep["device"].startswith("cuda") and rep["launches"] > 200 and rep["global_step"] == 2
    assert rep["teacher_none"] == 168 and rep["teacher_state_keys"] == 391
    assert rep["checkpoint_keys"] == gold["checkpoint_keys"]
    assert abs(rep["vae_lr"] - gold["trainer_step2"]["vae_lr"]) < 1e-12
    # (a) vs the golden: reference trainer + reference modules on CPU fp32 (same seeds, same sprites)
    m0, ref0 = rep["steps"][0], gold["trainer_step"]["metrics"]
    for k in ("recon_loss", "kl_loss", "vae_loss"):
        assert abs(m0[k] - ref0[k]) <= 0.03 * abs(ref0[k]) + 1e-4, (k, m0[k], ref0[k])
    for k in ("quality_scores", "quality_reward", "quality_loss"):
        assert abs(m0[k] - ref0[k]) <= 0.05, (k, m0[k], ref0[k])
    # (b) vs our own trainer on the same modules / seeds: same kernels, so only the host arithmetic differs
    from lunaris_orion_b200.train_hybrid import TrainingManager, build_arg_parser
    args = build_arg_parser().parse_args([
        "--data_dir", "synthetic", "--output_dir", str(tmp_path), "--batch_size", str(CFG["B"]),
        "--gradient_accumulation_steps", "1", "--latent_dim", str(CFG["latent"]), "--embedding_dim", str(CFG["emb"]),
        "--feature_dim", str(CFG["feat"]), "--seed", "42"])
    tm = TrainingManager(args, device=cuda_dev)
    for m in tm.teacher.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)):
            m.p = 0.0
    x = tc.images(CFG["B"], CFG["img_seed"]).to(cuda_dev)
    with tc.reference_eps(gold["trainer_step"]["eps_seed"]):
        mine = [tm._process_batch(x, i) for i in range(2)]
    for s in range(2):
        for k in ("recon_loss", "kl_loss", "vae_loss", "quality_scores", "teacher_loss", "baseline"):
            a, b = rep["steps"][s][k], mine[s][k]
            assert abs(a - b) <= 0.02 * abs(b) + 2e-3, (s, k, a, b)
