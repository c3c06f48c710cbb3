"""Dropout ON (the configuration bench.py times): one `_process_batch` with the reference's default dropout_rate 0.1
vs the CPU oracle fed the SAME masks ("mask injection", SURVEY.md 4 / 7 hard part 3).

The CUDA path draws its elementwise dropouts (feature-extractor Dropout, attn_drop, proj_drop; reference
lunar_evaluator.py:97,212,225) from a stateless counter RNG keyed by per-launch seeds, its Dropout2d / head dropouts
(:245,252,359,371) from torch bernoulli draws. The step's dropout trace records the seeds and keep-masks; the masks
are rebuilt on the host by tests/dropout_rng.py (a numpy restatement of csrc/elem_common.cuh) and injected into
oracle.restatement.train_step(masks_a=..., masks_b=...). Compared: the 12 metrics, the reconstruction and ALL 100 + 72
gradient tensors elementwise. Tolerance: 3x the drift of the same oracle executed under bf16 autocast on the GPU
(what a bf16 execution of the reference op sequence itself loses against fp32), plus a small floor.
"""
import numpy as np
import pytest
import torch

import dropout_rng as dr
import teacher_cases as tc
from oracle import restatement as R

P = 0.1
CFG = dict(B=2, latent=64, emb=32, feat=64)


@pytest.mark.gpu
def test_counter_rng_restatement_is_bit_exact(cuda_dev):
    """The kernels' keep decisions (csrc/elem_common.cuh drop_keep8) == tests/dropout_rng.keep, element for element:
    an all-ones tensor through the dropout of lun_affine_fwd_bf16 and through lun_proj_expand_bf16."""
    from lunaris_orion_b200 import _capi
    from lunaris_orion_b200.lunar_evaluator import _affine
    B, HW, C = 2, 4096, 192
    x = torch.ones(B, HW, C, device=cuda_dev, dtype=torch.bfloat16)
    for seed in (1, 0x1234567890ABCDEF, 2 ** 62 - 1):
        y = _affine(x, B, HW, C, None, None, seed=seed, drop_p=P)
        kept = (y.float().cpu().numpy() != 0).reshape(-1)
        want = dr.keep_range(seed, B * HW * C, P)
        assert np.array_equal(kept, want), seed
        assert abs(kept.mean() - (1 - P)) < 5e-3
        vals = torch.unique(y.float())
        assert vals.numel() == 2 and vals[0] == 0 and abs(float(vals[1]) - 1 / (1 - P)) < 0.01
    # proj_expand: rows >= nq are dropout(bias)
    Cp, nq, nq_pad = 64, 4096 // 32 + 31, 160
    small = torch.ones(B, nq_pad, Cp, device=cuda_dev, dtype=torch.bfloat16)
    bias = torch.ones(Cp, device=cuda_dev)
    h2 = torch.empty(B, HW, Cp, device=cuda_dev, dtype=torch.bfloat16)
    seed = 987654321012345
    _capi.check(_capi.lib().lun_proj_expand_bf16(small.data_ptr(), bias.data_ptr(), h2.data_ptr(), B, HW, Cp, nq,
                                                 nq_pad, seed, P, _capi.raw_stream()), "lun_proj_expand_bf16")
    kept = (h2.float().cpu().numpy() != 0).reshape(-1)
    assert np.array_equal(kept, dr.keep_range(seed, B * HW * Cp, P))


def _split_passes(trace):
    """The step runs the Teacher twice with the same structure: pass A events first, then pass B."""
    assert len(trace) % 2 == 0
    half = len(trace) // 2
    a, b = trace[:half], trace[half:]
    assert [(k, t) for k, t, _ in a] == [(k, t) for k, t, _ in b]
    return a, b


def _grad_errs(get, ref_grads):
    out = {}
    for n, rg in ref_grads.items():
        g = get(n)
        if n.endswith("shortcut.0.bias"):          # exact gradient is zero (bias in front of a train-mode BatchNorm)
            scale = ref_grads[n.replace("bias", "weight")].abs().max().item()
        else:
            scale = rg.abs().max().item()
        out[n] = (g - rg).abs().max().item() / (scale + 1e-20)
    return out


@pytest.mark.gpu
def test_step_with_dropout_on_matches_oracle_with_injected_masks(cuda_dev, tmp_path):
    from lunaris_orion_b200 import _host
    from lunaris_orion_b200.train_hybrid import TrainingManager, build_arg_parser
    B, feat = CFG["B"], CFG["feat"]
    # accumulation 2 and batch_idx 0: both backward passes run, no optimizer step -> raw (1/2-scaled) gradients stay
    args = build_arg_parser().parse_args([
        "--data_dir", "synthetic", "--output_dir", str(tmp_path), "--batch_size", str(B),
        "--gradient_accumulation_steps", "2", "--latent_dim", str(CFG["latent"]), "--embedding_dim", str(CFG["emb"]),
        "--feature_dim", str(feat), "--seed", "42"])
    tm = TrainingManager(args, device=cuda_dev)
    assert tm.teacher.dropout_rate == P
    # The default init makes the heads ill-conditioned (kaiming-fan_out MLPs on LayerNormed features: logits up to +-17,
    # saturated sigmoids - SURVEY.md 7 hard part 6): a bf16-sized logit difference then rescales each sample's gradient
    # by tens of percent. The arithmetic is compared at a well-conditioned point instead: the heads' Linear weights are
    # shrunk by 10x on both sides (logits of order 1); the ill-conditioned default is what the golden tests run.
    with torch.no_grad():
        for n, p_ in tm.teacher.named_parameters():
            if n.startswith(("gate", "quality_heads", "semantic_head")) and n.endswith("weight") and p_.dim() == 2:
                p_.mul_(0.1)
    x = tc.images(B, seed=41)
    vsd, tsd = tc.oracle_sd(tm.vae), tc.oracle_sd(tm.teacher)          # pre-step weights and BatchNorm buffers
    vsd_cal, tsd_cal = tc.oracle_sd(tm.vae), tc.oracle_sd(tm.teacher)  # second copy for the bf16 calibration run

    trace = []
    _host.set_dropout_trace(trace)
    try:
        torch.manual_seed(321)
        m = tm._process_batch(x.to(cuda_dev), 0)
    finally:
        _host.set_dropout_trace(None)
    torch.manual_seed(321)
    eps = torch.randn(B, CFG["latent"], device=cuda_dev).cpu()
    ev_a, ev_b = _split_passes(trace)
    tags = {t for _, t, _ in ev_a}
    # every dropout site of the reference is covered: FE, 12 x (Dropout2d, attn_drop, proj_drop, Dropout2d), 8 heads
    assert len(ev_a) == 1 + 12 * 4 + 8, sorted(tags)
    masks_a = dr.oracle_masks(ev_a, B, 128, 128, feat, P)
    masks_b = dr.oracle_masks(ev_b, B, 128, 128, feat, P)

    ref, ref_recon, _ = R.train_step(x, vsd, tsd, eps, masks_a=masks_a, masks_b=masks_b, accum=2)
    ref_grads = {"vae." + n: vsd[n].grad for n, _ in tm.vae.named_parameters()}
    ref_grads.update({"teacher." + n: tsd[n].grad for n, _ in tm.teacher.named_parameters() if tsd[n].grad is not None})
    mine = {"vae." + n: p.grad.detach().cpu().float() for n, p in tm.vae.named_parameters()}
    mine.update({"teacher." + n: p.grad.detach().cpu().float() for n, p in tm.teacher.named_parameters()
                 if p.grad is not None})
    assert set(mine) == set(ref_grads)
    assert len(mine) == 72 + 100

    # calibration: the same oracle, same masks, executed under bf16 autocast on the GPU
    dev = cuda_dev
    vc = {k: v.detach().to(dev) for k, v in vsd_cal.items()}
    tcal = {k: v.detach().to(dev) for k, v in tsd_cal.items()}
    for n, _ in tm.vae.named_parameters():
        vc[n].requires_grad_(True)
    for n, _ in tm.teacher.named_parameters():
        tcal[n].requires_grad_(True)
    to_dev = lambda d: {k: v.to(dev) for k, v in d.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        cal, cal_recon, _ = R.train_step(x.to(dev), vc, tcal, eps.to(dev), masks_a=to_dev(masks_a),
                                         masks_b=to_dev(masks_b), accum=2)
    cal_grads = {"vae." + n: vc[n].grad.detach().cpu().float() for n, _ in tm.vae.named_parameters()}
    cal_grads.update({"teacher." + n: tcal[n].grad.detach().cpu().float() for n, _ in tm.teacher.named_parameters()
                      if tcal[n].grad is not None})

    e_mine = _grad_errs(lambda n: mine[n], ref_grads)
    e_cal = _grad_errs(lambda n: cal_grads[n], ref_grads)
    worst_mine, worst_cal = max(e_mine.values()), max(e_cal.values())
    report = {"metrics": {k: (m[k], ref[k], cal[k]) for k in ref},
              "recon": (tc.rel_err(tm._last_recon, ref_recon), tc.rel_err(cal_recon, ref_recon)),
              "grad_worst_mine": sorted(e_mine.items(), key=lambda kv: -kv[1])[:6],
              "grad_worst_cal": sorted(e_cal.items(), key=lambda kv: -kv[1])[:6]}
    import json
    import os
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump(report, open(os.path.join(out_dir, "dropout_parity_report.json"), "w"), indent=1, default=str)

    for k in ("recon_loss", "kl_loss", "vae_loss"):
        assert abs(m[k] - ref[k]) <= 0.03 * abs(ref[k]) + 1e-4, (k, m[k], ref[k])
    for k in ("quality_scores", "quality_reward", "quality_loss", "semantic_reward"):
        tol = 3 * abs(cal[k] - ref[k]) + 0.03
        assert abs(m[k] - ref[k]) <= tol, (k, m[k], ref[k], cal[k])
    assert abs(m["advantage"]) < 1e-6 and abs(m["pg_loss"]) < 1e-6
    assert report["recon"][0] <= 3 * report["recon"][1] + 0.01, report["recon"]
    # ALL 72 + 100 gradient tensors elementwise against the oracle. VAE tensors: each within 3x its own bf16 calibration
    # (+2 % of its scale). Teacher tensors: BatchNorm centres every feature over the batch, so the per-image pooled
    # features that feed the heads sum to ~zero over the batch and every weight gradient downstream of them is a small
    # difference of large per-sample terms - the bf16-autocast oracle itself is 40-90 % off on several of them. A
    # per-tensor 3x bound on a ratio of two such noise responses is a coin flip; the bound is put on the distribution
    # instead: ratio r = err / (calibration err + 2 %), median <= 2, 90th percentile <= 5, worst <= 10 (a wrong mask,
    # seed or scale puts dozens of tensors at r >> 10; the single worst tensor - a gate weight whose calibration error
    # happens to be small - moved between 5.3 and 6.1 from box to box because the pooling atomics reorder the sums),
    # and every tensor's direction (cosine) stays above 0.8.
    bad = {n: (e_mine[n], e_cal[n]) for n in ref_grads if n.startswith("vae.") and e_mine[n] > 3 * e_cal[n] + 0.02}
    assert not bad, bad
    ratios = sorted(((e_mine[n] / (e_cal[n] + 0.02), n) for n in ref_grads if n.startswith("teacher.")), reverse=True)
    report["teacher_ratio_worst"] = ratios[:6]
    report["teacher_ratio_median"] = ratios[len(ratios) // 2][0]
    report["teacher_ratio_p90"] = ratios[len(ratios) // 10][0]
    cos = {}
    for n, rg in ref_grads.items():
        if n.startswith("teacher.") and not n.endswith("shortcut.0.bias"):
            g, r = mine[n].flatten().double(), rg.flatten().double()
            cos[n] = float((g @ r) / (g.norm() * r.norm() + 1e-30))
    report["teacher_cosine_worst"] = sorted(cos.items(), key=lambda kv: kv[1])[:6]
    if os.path.isdir(out_dir):
        json.dump(report, open(os.path.join(out_dir, "dropout_parity_report.json"), "w"), indent=1, default=str)
    assert report["teacher_ratio_median"] <= 2.0, report["teacher_ratio_median"]
    assert report["teacher_ratio_p90"] <= 5.0, ratios[:12]         # seven runs on four boxes: 1.8 - 3.1
    assert ratios[0][0] <= 10.0, ratios[:6]
    assert min(cos.values()) > 0.8, report["teacher_cosine_worst"]
    assert worst_mine <= 3 * worst_cal + 0.05, (report["grad_worst_mine"], report["grad_worst_cal"])


@pytest.mark.gpu
@pytest.mark.parametrize("feat", [64, 256])
def test_trunk_with_dropout_on_matches_oracle_elementwise(cuda_dev, feat):
    """Every dropout of the Teacher trunk ON (p = 0.1: feature-extractor Dropout, both Dropout2d, attn_drop inside
    attn_fold, proj_drop and its replay in the conv2 data-gradient epilogue), masks rebuilt from the step's dropout
    trace and injected into the oracle: pooled expert outputs and ALL live trunk gradient tensors elementwise, within 3x
    the bf16-autocast calibration (same masks)."""
    rep = tc.trunk_report(cuda_dev, B=2, feat=feat, calibrate=True, dropout=P)
    import json
    import os
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump(rep, open(os.path.join(out_dir, "dropout_trunk_report_feat%d.json" % feat), "w"), indent=1, default=str)
    assert rep["grad_keys_equal"]
    assert rep["pool"] <= 3 * rep["cal_pool"] + 2e-3, rep
    assert rep["grad_rel_max"] <= 3 * rep["cal_grad_rel_max"] + 0.02, (rep["grad_worst"], rep["cal_grad_worst"])
