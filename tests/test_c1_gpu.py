"""BASELINE.json configs[0] ("C1": batch 8, latent 256 / emb 128 / feat 256) at FULL size: three trainer steps of
lunaris_orion_b200.train_hybrid.TrainingManager on the B200 vs tests/golden/golden_c1.pt, which holds what the
UNMODIFIED reference trainer produced for the same seeds and sprites on CPU fp32 (oracle/make_golden_c1.py).

Tolerances are for a bf16-operand / fp32-accumulate path against fp32: losses 3 % relative; head outputs compared as
probabilities with an absolute bound (ill-conditioned sigmoid heads, SURVEY.md 7 hard part 6); gradients compared
through their fingerprints (signed sum, absolute sum, 8 samples).
"""
import json
import os

import pytest
import torch

import teacher_cases as tc

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_c1.pt")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="golden_c1.pt not generated")


_fp = tc.fingerprint


@pytest.mark.gpu
def test_two_trainer_steps_at_c1_match_the_reference_trainer(cuda_dev, tmp_path):
    from lunaris_orion_b200.train_hybrid import TrainingManager, build_arg_parser
    gold = torch.load(PATH, weights_only=False)
    cfg = gold["cfg"]
    args = build_arg_parser().parse_args([
        "--data_dir", "synthetic", "--output_dir", str(tmp_path), "--batch_size", str(cfg["B"]),
        "--gradient_accumulation_steps", "1", "--latent_dim", str(cfg["latent"]), "--embedding_dim", str(cfg["emb"]),
        "--feature_dim", str(cfg["feat"]), "--seed", str(cfg["seed"])])
    tm = TrainingManager(args, device=cuda_dev)
    for m in tm.teacher.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)):
            m.p = 0.0
    x = tc.images(cfg["B"], cfg["img_seed"]).to(cuda_dev)
    torch.manual_seed(cfg["eps_seed"])
    m0 = tm._process_batch(x, 0)
    ref0 = gold["step0"]

    report = {"step0": {k: (m0[k], ref0["metrics"][k]) for k in ref0["metrics"]}}
    for k in ("recon_loss", "kl_loss", "vae_loss"):
        assert abs(m0[k] - ref0["metrics"][k]) <= 0.03 * abs(ref0["metrics"][k]) + 1e-4, (k, m0[k], ref0["metrics"][k])
    for k in ("quality_scores", "quality_reward", "quality_loss"):
        assert abs(m0[k] - ref0["metrics"][k]) <= 0.05, (k, m0[k], ref0["metrics"][k])
    assert abs(m0["advantage"]) < 1e-6 and abs(m0["pg_loss"]) < 1e-6
    # the executed gradient set, BatchNorm counters and the schedule are exact
    none = sorted(n for n, p in tm.teacher.named_parameters() if p.grad is None)
    assert none == ref0["teacher_none"]
    sd = tm.teacher.state_dict()
    for k, v in ref0["teacher_nbt"].items():
        assert int(sd[k]) == v, k
    assert abs(tm.vae_optimizer.param_groups[0]["lr"] - ref0["vae_lr"]) < 1e-12
    assert abs(tm.teacher_optimizer.param_groups[0]["lr"] - ref0["teacher_lr"]) < 1e-12

    # gradients (as clip_grad_norm_ left them): aggregate L1 agreement and per-tensor magnitude agreement
    worst = {}
    for name, model, key in (("vae", tm.vae, "vae_grads"), ("teacher", tm.teacher, "teacher_grads")):
        num = den = 0.0
        per = {}
        for n, p in model.named_parameters():
            if p.grad is None:
                continue
            r, o = ref0[key][n], _fp(p.grad)
            num += abs(o["abs"] - r["abs"])
            den += r["abs"]
            per[n] = abs(o["abs"] - r["abs"]) / (r["abs"] + 1e-30)
        worst[name] = {"aggregate_l1": num / den, "worst_tensor": max(per.items(), key=lambda kv: kv[1])}
    report["grads"] = worst
    # parameters after clip + AdamW: Adam's first step moves every element by ~lr regardless of gradient scale, so
    # the updated weights pin the SIGN pattern of the gradients; compare the fingerprints' samples
    moved = {}
    for name, model, key in (("vae", tm.vae, "vae_params_after"), ("teacher", tm.teacher, "teacher_params_after")):
        bad = tot = 0
        for n, p in model.named_parameters():
            if n not in ref0[key]:
                continue
            d = (_fp(p)["samples"] - ref0[key][n]["samples"]).abs()
            lr = ref0["vae_lr" if name == "vae" else "teacher_lr"]
            bad += int((d > 2.5 * lr + 1e-7).sum())          # > one full opposite-sign Adam step (2 lr) + slack
            tot += d.numel()
        moved[name] = bad / tot
    report["param_samples_off_by_more_than_one_adam_step"] = moved

    m1 = tm._process_batch(x, 1)
    ref1 = gold["step1"]
    report["step1"] = {k: (m1[k], ref1["metrics"][k]) for k in ref1["metrics"]}
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump(report, open(os.path.join(out_dir, "c1_parity_report.json"), "w"), indent=1, default=str)

    assert worst["vae"]["aggregate_l1"] < 0.05, worst
    assert worst["teacher"]["aggregate_l1"] < 0.10, worst
    assert moved["vae"] == 0.0 and moved["teacher"] == 0.0, moved
    # second step runs on the UPDATED weights (optimizer boundary + schedule in the loop)
    for k in ("recon_loss", "kl_loss", "vae_loss"):
        assert abs(m1[k] - ref1["metrics"][k]) <= 0.03 * abs(ref1["metrics"][k]) + 1e-4, (k, m1[k], ref1["metrics"][k])
    for k in ("quality_scores", "quality_reward"):
        assert abs(m1[k] - ref1["metrics"][k]) <= 0.08, (k, m1[k], ref1["metrics"][k])
    assert abs(tm.vae_optimizer.param_groups[0]["lr"] - ref1["vae_lr"]) < 1e-12
    # third consecutive step (">= 3 consecutive steps"): two optimizer updates deep, the bf16 path and the
    # fp32 reference have drifted apart a little more; BatchNorm counters and the schedule stay exact
    if "step2" in gold:
        m2 = tm._process_batch(x, 2)
        ref2 = gold["step2"]
        report["step2"] = {k: (m2[k], ref2["metrics"][k]) for k in ref2["metrics"]}
        if os.path.isdir(out_dir):
            json.dump(report, open(os.path.join(out_dir, "c1_parity_report.json"), "w"), indent=1, default=str)
        for k in ("recon_loss", "vae_loss"):
            assert abs(m2[k] - ref2["metrics"][k]) <= 0.05 * abs(ref2["metrics"][k]) + 1e-4, (k, m2[k], ref2["metrics"][k])
        assert abs(m2["kl_loss"] - ref2["metrics"]["kl_loss"]) <= 0.10 * abs(ref2["metrics"]["kl_loss"]) + 1e-3
        assert abs(m2["quality_scores"] - ref2["metrics"]["quality_scores"]) <= 0.10
        assert abs(tm.vae_optimizer.param_groups[0]["lr"] - ref2["vae_lr"]) < 1e-12
        sd = tm.teacher.state_dict()
        for k, v in ref2["teacher_nbt"].items():
            assert int(sd[k]) == v, k
