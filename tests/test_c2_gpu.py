"""BASELINE.json configs[1] ("C2": batch 16, latent 256 / emb 128 / feat 256, gradient accumulation 4) at FULL size:
five `_process_batch` calls of lunaris_orion_b200.train_hybrid.TrainingManager on the B200 vs tests/golden/golden_c2.pt,
which holds what the UNMODIFIED reference trainer produced for the same seeds and sprites on CPU fp32
(oracle/make_golden_c2.py). Calls 0-3 are one accumulation window on four different batches (weights must not move
until the fourth; the gradient that reaches the optimizer is the fourth micro-batch's, scaled by 1/4 - the reference
zeroes gradients at the top of every micro-batch); call 4 runs on the updated weights.

Tolerances are those of tests/test_c1_gpu.py (bf16 operands / fp32 accumulation against fp32): losses 3 % relative,
sigmoid head means 0.10 absolute (measured over the five calls: 0.002-0.047; the heads sit behind a LayerNorm of pooled
features whose spread across channels is tiny on noise sprites, SURVEY.md 7 hard part 6), gradient fingerprints 5 % (VAE) / 10 % (Teacher) aggregate L1, updated weights
within one Adam step of the reference's. The advantage is (reward - EMA baseline) * 0.1 with reward = quality mean +
0.5 * semantic score, i.e. a difference of ill-conditioned sigmoid-head means (SURVEY.md 7 hard part 6; the C1 report
shows the semantic head alone moving by 0.12 between fp32 and bf16): absolute bound 0.02.
"""
import json
import os

import pytest
import torch

import teacher_cases as tc

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_c2.pt")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="golden_c2.pt not generated")


_fp = tc.fingerprint


@pytest.mark.gpu
def test_accumulation_window_at_c2_matches_the_reference_trainer(cuda_dev, tmp_path):
    from lunaris_orion_b200.train_hybrid import TrainingManager, build_arg_parser
    gold = torch.load(PATH, weights_only=False)
    cfg = gold["cfg"]
    accum = cfg["accum"]
    args = build_arg_parser().parse_args([
        "--data_dir", "synthetic", "--output_dir", str(tmp_path), "--batch_size", str(cfg["B"]),
        "--gradient_accumulation_steps", str(accum), "--latent_dim", str(cfg["latent"]),
        "--embedding_dim", str(cfg["emb"]), "--feature_dim", str(cfg["feat"]), "--seed", str(cfg["seed"])])
    tm = TrainingManager(args, device=cuda_dev)
    for m in tm.teacher.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)):
            m.p = 0.0
    probe = tm.vae.decoder.final_conv.weight
    w0 = probe.detach().clone()
    lr0 = tm.vae_optimizer.param_groups[0]["lr"]
    eps = tc.reference_eps(cfg["eps_seed"])        # the reference's CPU-generator noise, call after call
    report = {"calls": []}
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")

    def dump():
        if os.path.isdir(out_dir):
            json.dump(report, open(os.path.join(out_dir, "c2_parity_report.json"), "w"), indent=1, default=str)

    for i, ref in enumerate(gold["calls"]):
        x = tc.images(cfg["B"], cfg["img_seed"] + i).to(cuda_dev)
        with eps:
            m = tm._process_batch(x, i)
        rm = ref["metrics"]
        report["calls"].append({k: (m[k], rm[k]) for k in rm})
        dump()
        drift = 1.0 if i < accum else 2.0          # call 4 runs on weights one optimizer step deep
        for k in ("recon_loss", "kl_loss"):
            assert abs(m[k] - rm[k]) <= 0.03 * drift * abs(rm[k]) + 1e-4, (i, k, m[k], rm[k])
        for k in ("quality_scores", "quality_reward"):
            assert abs(m[k] - rm[k]) <= 0.10, (i, k, m[k], rm[k])
        assert abs(m["advantage"] - rm["advantage"]) <= 0.02 * drift, (i, m["advantage"], rm["advantage"])
        assert abs(m["pg_loss"] - rm["pg_loss"]) <= 0.02 * drift * abs(rm["recon_loss"]) + 1e-5, (i, m["pg_loss"])
        # reported losses carry the 1/accum factor (train_hybrid.py:886-896)
        assert abs(m["vae_loss"] * accum - (m["recon_loss"] + 0.1 * m["kl_loss"] + m["pg_loss"])) < 1e-3, (i, m)
        assert abs(m["vae_loss"] - rm["vae_loss"]) <= 0.03 * drift * abs(rm["vae_loss"]) + 0.01 / accum, (i, m["vae_loss"])
        assert abs(tm.vae_optimizer.param_groups[0]["lr"] - ref["vae_lr"]) < 1e-12, i
        assert abs(tm.teacher_optimizer.param_groups[0]["lr"] - ref["teacher_lr"]) < 1e-12, i
        if i < accum - 1:
            # inside the window: no optimizer step, no scheduler step
            assert torch.equal(w0, probe.detach()), i
            assert tm.vae_optimizer.param_groups[0]["lr"] == lr0
            assert abs(_fp(probe)["abs"] - ref["probe_weight"]["abs"]) <= 1e-6 * ref["probe_weight"]["abs"]
        if i == accum - 1:
            assert not torch.equal(w0, probe.detach())
            none = sorted(n for n, p in tm.teacher.named_parameters() if p.grad is None)
            assert none == ref["teacher_none"]
            sd = tm.teacher.state_dict()
            for k, v in ref["teacher_nbt"].items():
                assert int(sd[k]) == v, k
            worst = {}
            for name, model, key in (("vae", tm.vae, "vae_grads"), ("teacher", tm.teacher, "teacher_grads")):
                num = den = 0.0
                per = {}
                for n, p in model.named_parameters():
                    if p.grad is None:
                        continue
                    r, o = ref[key][n], _fp(p.grad)
                    num += abs(o["abs"] - r["abs"])
                    den += r["abs"]
                    per[n] = abs(o["abs"] - r["abs"]) / (r["abs"] + 1e-30)
                worst[name] = {"aggregate_l1": num / den, "worst_tensor": max(per.items(), key=lambda kv: kv[1])}
            report["grads"] = worst
            # value / sign agreement of the fingerprints' gradient samples, and updated parameters: Adam's first step
            # moves every element by lr * sign(g), so a sample further than ONE lr from the reference's updated value
            # stepped in the opposite direction (a fully sign-flipped gradient would put every sample 2 lr away)
            agree, flips = {}, {}
            for name, model, key, pkey in (("vae", tm.vae, "vae_grads", "vae_params_after"),
                                           ("teacher", tm.teacher, "teacher_grads", "teacher_params_after")):
                agree[name] = tc.sample_agreement(((n, p.grad) for n, p in model.named_parameters()
                                                   if p.grad is not None), ref[key])
                bad = tot = 0
                lr = lr0 if name == "vae" else args.teacher_lr
                for n, p in model.named_parameters():
                    if n not in ref[pkey] or n.endswith("shortcut.0.bias"):
                        continue
                    d = (_fp(p)["samples"] - ref[pkey][n]["samples"]).abs()
                    bad += int((d > 1.0 * lr).sum())
                    tot += d.numel()
                flips[name] = bad / tot
            report["grad_sample_agreement"] = agree
            report["param_samples_stepped_in_opposite_direction"] = flips
            dump()
            assert worst["vae"]["aggregate_l1"] < 0.05, worst
            assert worst["teacher"]["aggregate_l1"] < 0.10, worst
            # VAE gradients: value, sign and direction on the samples. Teacher gradients all pass through d sigmoid(quality
            # logits) of the ill-conditioned heads (kaiming-fan_out MLPs on LayerNormed features, logits up to +-17: SURVEY.md 7
            # hard part 6), which rescales every sample's contribution; they are held to direction / sign here and elementwise
            # behind the heads (tests/test_c3_gpu.py trunk test, tests/test_dropout_parity_gpu.py)
            assert agree["vae"]["cosine"] > 0.999 and agree["vae"]["sign_agree"] > 0.99, agree
            assert agree["vae"]["value_within_10pct"] > 0.95 and flips["vae"] < 0.02, (agree, flips)
            assert agree["teacher"]["cosine"] > 0.97 and agree["teacher"]["sign_agree"] > 0.95, agree
            assert flips["teacher"] < 0.08, flips
