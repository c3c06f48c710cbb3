"""One hybrid step through lunaris_orion_b200.train_hybrid.TrainingManager on the GPU vs (a) the 12 metrics of the
real reference trainer stored in tests/golden, (b) the CPU oracle, plus checkpoint layout / resume and LR schedule."""
import os

import pytest
import torch

import teacher_cases as tc

gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_small.pt"),
                  weights_only=False)
CFG = gold["cfg"]


def _manager(dev, tmp_path, extra=()):
    from lunaris_orion_b200.train_hybrid import TrainingManager, build_arg_parser
    args = build_arg_parser().parse_args([
        "--data_dir", "synthetic", "--output_dir", str(tmp_path), "--batch_size", str(CFG["B"]),
        "--gradient_accumulation_steps", "1", "--latent_dim", str(CFG["latent"]), "--embedding_dim", str(CFG["emb"]),
        "--feature_dim", str(CFG["feat"]), "--seed", "42", *extra])
    tm = TrainingManager(args, device=dev)
    for m in tm.teacher.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)):
            m.p = 0.0
    return tm


@pytest.mark.gpu
def test_step_metrics_match_reference_trainer_golden(cuda_dev, tmp_path):
    tm = _manager(cuda_dev, tmp_path)
    x = tc.images(CFG["B"], CFG["img_seed"]).to(cuda_dev)
    eps = tc.reference_eps(gold["trainer_step"]["eps_seed"])     # the golden's CPU-generator noise
    with eps:
        m = tm._process_batch(x, 0)
    ref = gold["trainer_step"]["metrics"]
    # bf16 tolerance: losses 3 % relative; quality / semantic outputs are sigmoid values of ill-conditioned heads
    for k in ("recon_loss", "kl_loss", "vae_loss"):
        assert abs(m[k] - ref[k]) <= 0.03 * abs(ref[k]) + 1e-4, (k, m[k], ref[k])
    for k in ("quality_scores", "quality_reward", "quality_loss"):
        assert abs(m[k] - ref[k]) <= 0.05, (k, m[k], ref[k])
    assert abs(m["advantage"]) < 1e-6 and abs(m["pg_loss"]) < 1e-6           # first step: baseline == reward mean
    none = sorted(n for n, p in tm.teacher.named_parameters() if p.grad is None)
    assert none == gold["trainer_step"]["teacher_none"]
    assert all(p.grad is not None for p in tm.vae.parameters())
    sd = tm.teacher.state_dict()
    for k, v in gold["trainer_step"]["teacher_nbt"].items():
        assert int(sd[k]) == v, k
    assert abs(tm.vae_optimizer.param_groups[0]["lr"] - gold["trainer_step"]["vae_lr"]) < 1e-12
    with eps:
        m2 = tm._process_batch(x, 1)
    assert abs(tm.vae_optimizer.param_groups[0]["lr"] - gold["trainer_step2"]["vae_lr"]) < 1e-12
    ref2 = gold["trainer_step2"]["metrics"]
    for k in ("recon_loss", "vae_loss"):                      # second step: on the updated weights
        assert abs(m2[k] - ref2[k]) <= 0.05 * abs(ref2[k]) + 1e-4, (k, m2[k], ref2[k])


@pytest.mark.gpu
def test_checkpoint_layout_and_resume(cuda_dev, tmp_path):
    tm = _manager(cuda_dev, tmp_path)
    x = tc.images(CFG["B"], 11).to(cuda_dev)
    tm._process_batch(x, 0)
    path = tm._save_checkpoint()
    ck = torch.load(path, weights_only=True)
    assert sorted(ck.keys()) == gold["checkpoint_keys"]
    assert sorted(ck["args"].keys()) == gold["checkpoint_args_keys"]
    assert list(ck["teacher_state_dict"].keys()) == gold["teacher_keys_after_forward"]
    assert list(ck["vae_state_dict"].keys()) == gold["vae_keys"]
    tm2 = _manager(cuda_dev, tmp_path, extra=("--resume_from", path))
    assert tm2.global_step == 1
    for (k, a), (_, b) in zip(tm.vae.state_dict().items(), tm2.vae.state_dict().items()):
        assert torch.equal(a, b), k
    assert abs(tm2.vae_optimizer.param_groups[0]["lr"] - tm.vae_optimizer.param_groups[0]["lr"]) < 1e-12


@pytest.mark.gpu
def test_grad_accumulation_quirk(cuda_dev, tmp_path):
    """--gradient_accumulation_steps 2: no optimizer step on the first micro-batch, grads zeroed each micro-batch and
    scaled by 1/2."""
    tm = _manager(cuda_dev, tmp_path)
    tm.args.gradient_accumulation_steps = 2
    x = tc.images(CFG["B"], 12).to(cuda_dev)
    w0 = tm.vae.decoder.final_conv.weight.detach().clone()
    m = tm._process_batch(x, 0)
    assert torch.equal(w0, tm.vae.decoder.final_conv.weight.detach())
    assert abs(m["vae_loss"] * 2 - (m["recon_loss"] + 0.1 * m["kl_loss"] + m["pg_loss"])) < 1e-4
    tm._process_batch(x, 1)
    assert not torch.equal(w0, tm.vae.decoder.final_conv.weight.detach())


@pytest.mark.gpu
def test_training_makes_progress_and_stays_finite(cuda_dev, tmp_path):
    """Ten optimizer steps on a fixed batch with the reference defaults (dropout on): reconstruction loss goes down,
    every metric stays finite, the Teacher's frozen tensors do not move and its live ones do."""
    from lunaris_orion_b200.train_hybrid import TrainingManager, build_arg_parser
    args = build_arg_parser().parse_args([
        "--data_dir", "synthetic", "--output_dir", str(tmp_path), "--batch_size", "4",
        "--gradient_accumulation_steps", "1", "--latent_dim", "64", "--embedding_dim", "32", "--feature_dim", "64",
        "--seed", "7"])                              # reference default learning rates (1e-4)
    tm = TrainingManager(args, device=cuda_dev)
    x = tc.images(4, 21).to(cuda_dev)
    frozen0 = tm.teacher.experts[0][1].conv1[0].weight.detach().clone()
    live0 = tm.teacher.experts[0][1].conv2[0].weight.detach().clone()
    hist = [tm._process_batch(x, i) for i in range(10)]
    for m in hist:
        assert all(v == v and abs(v) < 1e6 for v in m.values()), m
    assert hist[-1]["recon_loss"] < 0.9 * hist[0]["recon_loss"], (hist[0]["recon_loss"], hist[-1]["recon_loss"])
    assert torch.equal(frozen0, tm.teacher.experts[0][1].conv1[0].weight.detach())
    assert not torch.equal(live0, tm.teacher.experts[0][1].conv2[0].weight.detach())
    assert tm.global_step == 10


@pytest.mark.gpu
def test_clip_adamw_matches_torch_clip_plus_adamw(cuda_dev):
    """lunaris_orion_b200.optim.ClipAdamW == clip_grad_norm_(max_norm) + torch.optim.AdamW over several steps, with a
    parameter that never receives a gradient (reference None-set) and an identical state_dict layout."""
    from lunaris_orion_b200.optim import ClipAdamW
    g = torch.Generator().manual_seed(0)
    shapes = [(64, 32, 3, 3), (512,), (1, 48, 1, 1), (300, 7)]
    pa = [torch.randn(s, generator=g).to(cuda_dev).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    dead_a = torch.zeros(5, device=cuda_dev, requires_grad=True)
    dead_b = torch.zeros(5, device=cuda_dev, requires_grad=True)
    oa = ClipAdamW(pa + [dead_a], lr=3e-3, weight_decay=0.01, max_grad_norm=1.0)
    ob = torch.optim.AdamW(pb + [dead_b], lr=3e-3, weight_decay=0.01, betas=(0.9, 0.999))
    for step in range(4):
        for p, q in zip(pa, pb):
            gr = torch.randn(p.shape, generator=g).to(cuda_dev) * (10.0 if step % 2 == 0 else 0.01)
            p.grad, q.grad = gr.clone(), gr.clone()
        torch.nn.utils.clip_grad_norm_(pb + [dead_b], 1.0)
        ob.step()
        oa.step()
        for p, q in zip(pa, pb):
            assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), step
            assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-8), step
            assert p._version > step, "the raw-pointer update must be visible to autograd's version counter"
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa["state"].keys() == sb["state"].keys()
    for k in sa["state"]:
        assert sa["state"][k].keys() == sb["state"][k].keys()
        assert float(sa["state"][k]["step"]) == float(sb["state"][k]["step"]) == 4.0
        assert torch.allclose(sa["state"][k]["exp_avg_sq"], sb["state"][k]["exp_avg_sq"], rtol=1e-5, atol=1e-10)
    assert set(sb["param_groups"][0].keys()) <= set(sa["param_groups"][0].keys())
    assert dead_a.grad is None and len(oa.state[dead_a]) == 0


@pytest.mark.gpu
def test_clip_adamw_param_groups_late_gradients_and_grad_scale(cuda_dev):
    """Per-group hyper-parameters, a parameter whose first gradient arrives two steps late (its own bias-correction
    step count, like torch's per-tensor state['step']), and grad_scale = 1/world on summed gradients == torch
    clip + AdamW on the averaged gradients."""
    from lunaris_orion_b200 import _capi
    from lunaris_orion_b200.optim import ClipAdamW
    g = torch.Generator().manual_seed(3)
    shapes = [(40, 9), (33,), (5, 4, 3, 3)]
    pa = [torch.randn(s, generator=g).to(cuda_dev).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]

    def groups(ps):
        return [dict(params=ps[:1], lr=1e-2, weight_decay=0.1, betas=(0.8, 0.99)), dict(params=ps[1:], lr=1e-3)]
    oa = ClipAdamW(groups(pa), weight_decay=0.01, max_grad_norm=0.5)
    oa.grad_scale = 0.25
    ob = torch.optim.AdamW(groups(pb), weight_decay=0.01)
    for step in range(5):
        for i, (p, q) in enumerate(zip(pa, pb)):
            if i == 2 and step < 2:
                p.grad = q.grad = None                   # late first gradient
                continue
            gr = torch.randn(p.shape, generator=g).to(cuda_dev) * 3.0
            p.grad, q.grad = gr.clone() * 4.0, gr.clone()             # ours holds the 4-rank SUM, torch the average
        torch.nn.utils.clip_grad_norm_(pb, 0.5)
        ob.step()
        oa.step()
        for p, q in zip(pa, pb):
            assert torch.allclose(p, q, rtol=3e-5, atol=3e-6), step
    assert float(oa.state[pa[2]]["step"]) == float(ob.state[pb[2]]["step"]) == 3.0
    assert float(oa.state[pa[0]]["step"]) == 5.0
    # state_dict round trip keeps the per-tensor counters (host mirror re-read from the restored state)
    oc = ClipAdamW(groups([p.detach().clone().requires_grad_(True) for p in pa]), weight_decay=0.01, max_grad_norm=0.5)
    oc.load_state_dict(oa.state_dict())
    assert [oc._host_step(p, oc.state[p]) for gr_ in oc.param_groups for p in gr_["params"]] == [5, 5, 3]
    with pytest.raises(_capi.LunarisB200Error):
        bad = ClipAdamW([torch.zeros(3, device=cuda_dev, requires_grad=True)])
        bad.param_groups[0]["amsgrad"] = True
        bad.param_groups[0]["params"][0].grad = torch.ones(3, device=cuda_dev)
        bad.step()


@pytest.mark.gpu
def test_forward_after_an_optimizer_step_runs_on_the_updated_weights(cuda_dev, tmp_path):
    """The kernels consume packed bf16 shadows of the fp32 parameters, cached per parameter version. After a fused
    clip + AdamW step (which writes the parameters through raw pointers) the next forward must see the NEW weights:
    its outputs equal those of freshly built modules loaded from the updated state_dict (to the run-to-run noise of
    the atomically accumulated GroupNorm / BatchNorm sums), and differ clearly from the pre-step outputs."""
    from lunaris_orion_b200 import lunar_evaluator as le, lunar_generate as lg
    tm = _manager(cuda_dev, tmp_path)
    x = tc.images(CFG["B"], CFG["img_seed"]).to(cuda_dev)

    def outputs(vae, teacher):
        vae.eval(), teacher.eval()
        with torch.no_grad():
            _, mu, _ = vae(x)
            return mu.float().clone(), tc.logit(teacher(x)["quality_scores"])
    mu_before, q_before = outputs(tm.vae, tm.teacher)
    tm.vae.train(), tm.teacher.train()
    torch.manual_seed(1)
    tm._process_batch(x, 0)                       # forward (fills the caches) + backward + optimizer step
    mu_live, q_live = outputs(tm.vae, tm.teacher)
    vae2 = lg.LunarisCoreVAE(CFG["latent"]).to(cuda_dev)
    t2 = le.LunarMoETeacher(feature_dim=CFG["feat"], embedding_dim=CFG["emb"], dropout_rate=0.0).to(cuda_dev)
    vae2.load_state_dict(tm.vae.state_dict())
    t2.load_state_dict(tm.teacher.state_dict(), strict=False)
    mu_new, q_new = outputs(vae2, t2)
    scale = mu_new.abs().max().item()
    moved = (mu_new - mu_before).abs().max().item()
    assert (mu_live - mu_new).abs().max().item() < 0.02 * scale, "live modules ran on stale packed weights"
    assert moved > 0.2 * scale, "one AdamW step must move mu visibly (32768-wide fc), else this test proves nothing"
    assert (q_live - q_new).abs().max().item() < 0.05 * (q_new.abs().max().item() + 1.0)


@pytest.mark.gpu
def test_cli_trains_an_epoch_from_sprite_files_and_resumes(cuda_dev, tmp_path):
    """The reference's command line end to end (train_hybrid.py:1076-1135 flags, generate.py:858-904 on-disk format):
    sprites_000.npy + labels_000.csv -> one epoch through main() -> latest.pt / best.pt -> a second run with
    --resume_from continues from the saved global step."""
    import numpy as np
    from lunaris_orion_b200 import train_hybrid as th
    data = tmp_path / "data"
    data.mkdir()
    np.save(data / "sprites_000.npy", np.random.default_rng(1234).integers(0, 256, (12, 128, 128, 3), dtype=np.uint8))
    with open(data / "labels_000.csv", "w") as f:
        f.write("filename,category,prompt,seed,pixel_size,guidance_scale,pag_scale,num_steps\n")
        for i in range(12):
            f.write(f"s{i}.png,cat,prompt,{i},8,7.5,3.0,20\n")
    out = tmp_path / "out"
    argv = ["--data_dir", str(data), "--output_dir", str(out), "--batch_size", "4", "--num_epochs", "1",
            "--gradient_accumulation_steps", "1", "--latent_dim", "64", "--embedding_dim", "32", "--feature_dim", "64"]
    th.main(argv)
    ck = torch.load(out / "checkpoints" / "latest.pt", weights_only=True)
    # 12 sprites -> 90/10 split like the reference (train_hybrid.py:551-555): 10 training sprites, 2 full batches of 4
    assert ck["global_step"] == 2 and (out / "checkpoints" / "best.pt").exists()
    assert all(torch.isfinite(v).all() for v in ck["vae_state_dict"].values() if v.is_floating_point())
    args = th.build_arg_parser().parse_args(argv + ["--resume_from", str(out / "checkpoints" / "latest.pt")])
    tm = th.TrainingManager(args, device=cuda_dev)
    assert tm.global_step == 2
    assert abs(tm.vae_optimizer.param_groups[0]["lr"] - ck["vae_optimizer"]["param_groups"][0]["lr"]) < 1e-12
    m = tm._process_batch(tc.images(4, 2).to(cuda_dev), 0)
    assert all(v == v for v in m.values()) and tm.global_step == 3
