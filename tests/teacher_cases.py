"""Teacher parity helpers: lunaris_orion_b200.lunar_evaluator (CUDA) vs oracle/restatement.py (CPU fp32)."""
import torch

from oracle import restatement as R


def make_teacher(dev, feat=64, emb=32, dropout=0.0, seed=11):
    from lunaris_orion_b200 import lunar_evaluator as le
    torch.manual_seed(seed)
    return le.LunarMoETeacher(feature_dim=feat, embedding_dim=emb, dropout_rate=dropout).to(dev)


def oracle_sd(module):
    sd = {k: v.detach().cpu().float().clone() if v.is_floating_point() else v.detach().cpu().clone()
          for k, v in module.state_dict().items()}
    for n, _ in module.named_parameters():
        sd[n].requires_grad_(True)
    return sd


def images(B, seed=5):
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (B, 3, 128, 128), generator=g, dtype=torch.uint8)
    return u8.float() / 127.5 - 1.0


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


def fingerprint(t):
    """sum, |sum| and 8 evenly spaced samples of a tensor - the form the golden fixtures store gradients / weights in
    (oracle/make_golden.py::fingerprint)."""
    t = t.detach().double().flatten().cpu()
    idx = torch.linspace(0, t.numel() - 1, 8).long()
    return {"sum": t.sum().item(), "abs": t.abs().sum().item(), "n": t.numel(), "samples": t[idx].float()}


def logit(p):
    p = p.detach().double().cpu().clamp(1e-12, 1 - 1e-12)
    return torch.log(p / (1 - p))


def logit_c(p=None, logits=None, lim=12.0):
    """Logit clamped to +-lim: an fp32 probability cannot resolve logits beyond ~+-15 (1 - p underflows the spacing
    of floats near 1), so saturated heads are compared on the clamped scale. Pass probabilities or raw logits."""
    z = logit(p) if logits is None else logits.detach().double().cpu()
    return z.clamp(-lim, lim)


def eval_report(dev, B=2):
    t = make_teacher(dev).eval()
    sd = oracle_sd(t)
    x = images(B)
    with torch.no_grad():
        out = t(x.to(dev))
        ref = R.teacher_forward(x, sd, training=False)
    rep = {}
    for k in ("expert_weights", "style_embedding", "prompt_embedding"):
        rep[k] = rel_err(out[k], ref[k])
    rep["quality_logit_abs"] = (logit(out["quality_scores"]) - ref["quality_logits"].double()).abs().max().item()
    rep["quality_logit_scale"] = ref["quality_logits"].abs().max().item()
    rep["semantic_logit_abs"] = (logit(out["semantic_score"]) - ref["semantic_logit"].double()).abs().max().item()
    rep["semantic_logit_scale"] = ref["semantic_logit"].abs().max().item()
    rep["feature_maps"] = max(rel_err(a, b) for a, b in zip(out["feature_maps"], ref["feature_maps"]))
    return rep


def train_report(dev, B=2):
    t = make_teacher(dev).train()
    sd = oracle_sd(t)
    x = images(B, seed=6)
    out = t(x.to(dev))
    (-out["quality_scores"].mean() * 0.5).backward()
    ref = R.teacher_forward(x, sd, training=True)
    (-ref["quality_scores"].mean() * 0.5).backward()
    rep = {"quality_logit_abs": (logit(out["quality_scores"]) - ref["quality_logits"].double()).abs().max().item(),
           "quality_logit_scale": ref["quality_logits"].abs().max().item()}
    none_mine = {n for n, p in t.named_parameters() if p.grad is None}
    none_ref = {n for n, _ in t.named_parameters() if sd[n].grad is None}
    rep["none_set_equal"] = none_mine == none_ref
    rep["n_none"] = len(none_mine)
    grads = {}
    for n, p in t.named_parameters():
        if p.grad is None or sd[n].grad is None:
            continue
        ref_g = sd[n].grad
        if n.endswith("shortcut.0.bias"):
            scale = sd[n.replace("bias", "weight")].grad.abs().max().item()   # exact gradient is zero
        else:
            scale = ref_g.abs().max().item()
        grads[n] = (p.grad.detach().cpu().float() - ref_g).abs().max().item() / (scale + 1e-20)
    rep["grad_rel_max"] = max(grads.values())
    rep["grad_worst"] = sorted(grads.items(), key=lambda kv: -kv[1])[:6]
    bn = {}
    mine = t.state_dict()
    for k, v in mine.items():
        if k.endswith("num_batches_tracked"):
            bn[k] = float(int(v) != int(sd[k]))
        elif "running_" in k:
            bn[k] = rel_err(v, sd[k])
    rep["bn_buffers_max"] = max(bn.values())
    rep["bn_worst"] = [(k, v, mine[k].flatten()[:3].tolist(), sd[k].flatten()[:3].tolist())
                       for k, v in sorted(bn.items(), key=lambda kv: -kv[1])[:8]]
    return rep


def _trunk_oracle(x, sd, G, autocast_dev=None, masks=None):
    """Oracle trunk loss sum_e <pooled_e, G_e> / HW and its parameter gradients (fp32 CPU, or bf16 autocast on GPU
    to calibrate how far a bf16 execution of the reference op sequence drifts from fp32)."""
    import contextlib
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast_dev is not None else contextlib.nullcontext()
    with ctx:
        fe, outs = R.teacher_trunk(x, sd, training=True, masks=masks)
        pooled = [o.float().sum((2, 3)) for o in outs]
        loss = sum((p * g).sum() for p, g in zip(pooled, G)) / (x.shape[2] * x.shape[3])
    loss.backward()
    return fe.detach().float().sum((2, 3)), [p.detach() for p in pooled]


def trunk_report(dev, B=2, feat=64, calibrate=True, dropout=0.0):
    """Trunk forward + backward (pooled expert outputs contracted with fixed random cotangents G) vs the oracle.
    dropout > 0: every dropout of the trunk is ON; the masks the CUDA path drew are rebuilt from its dropout trace
    (tests/dropout_rng.py) and injected into the oracle, so the comparison stays elementwise."""
    from lunaris_orion_b200 import _host, lunar_evaluator as le
    t = make_teacher(dev, feat=feat, dropout=dropout).train()
    x = images(B, seed=8)
    g = torch.Generator().manual_seed(9)
    G = [torch.randn(B, feat, generator=g) for _ in range(4)]
    HW = 128 * 128
    # mine
    params = le._trunk_grad_params(t)
    trace = []
    _host.set_dropout_trace(trace if dropout > 0 else None)
    try:
        res = le._TeacherTrunk.apply(t, x.to(dev), True, *params)
    finally:
        _host.set_dropout_trace(None)
    masks = None
    if dropout > 0:
        import dropout_rng
        assert len(trace) == 1 + 12 * 4, [e[1] for e in trace]
        masks = dropout_rng.oracle_masks(trace, B, 128, 128, feat, dropout)
    loss = sum((res[1][e] * gg.to(dev)).sum() for e, gg in enumerate(G)) / HW
    loss.backward()
    mine_pool = [res[1][e].detach().cpu() for e in range(4)]
    mine_fe = res[0].detach().cpu()
    mine_grads = {n: p.grad.detach().cpu().float() for n, p in t.named_parameters() if p.grad is not None}
    # oracle fp32 (CPU)
    sd = oracle_sd(t)
    for k in sd:
        if "running_" in k or k.endswith("num_batches_tracked"):
            pass
    t2 = make_teacher(dev, feat=feat, dropout=dropout)   # fresh buffers for the oracle (same seed => same params)
    sd = oracle_sd(t2)
    ref_fe, ref_pool = _trunk_oracle(x, sd, G, masks=masks)
    ref_grads = {n: sd[n].grad for n, _ in t2.named_parameters() if sd[n].grad is not None}
    rep = {"fe_pool": rel_err(mine_fe, ref_fe),
           "pool": max(rel_err(a, b) for a, b in zip(mine_pool, ref_pool)),
           "grad_keys_equal": set(mine_grads) == set(ref_grads)}

    def grad_errs(gr):
        out = {}
        for n, rg in ref_grads.items():
            if n not in gr:
                continue
            scale = ref_grads[n.replace("bias", "weight")].abs().max().item() if n.endswith("shortcut.0.bias") \
                else rg.abs().max().item()
            out[n] = (gr[n] - rg).abs().max().item() / (scale + 1e-20)
        return out
    ge = grad_errs(mine_grads)
    rep["grad_rel_max"] = max(ge.values())
    rep["grad_worst"] = sorted(ge.items(), key=lambda kv: -kv[1])[:5]
    if calibrate:
        sdc = {k: v.detach().to(dev) for k, v in oracle_sd(make_teacher(dev, feat=feat, dropout=dropout)).items()}
        for n, _ in t2.named_parameters():
            sdc[n].requires_grad_(True)
        cal_fe, cal_pool = _trunk_oracle(x.to(dev), sdc, [gg.to(dev) for gg in G], autocast_dev=dev,
                                         masks=None if masks is None else {k: v.to(dev) for k, v in masks.items()})
        cal_grads = {n: sdc[n].grad.detach().cpu().float() for n in ref_grads if sdc[n].grad is not None}
        ce = grad_errs(cal_grads)
        rep["cal_pool"] = max(rel_err(a, b) for a, b in zip(cal_pool, ref_pool))
        rep["cal_grad_rel_max"] = max(ce.values())
        rep["cal_grad_worst"] = sorted(ce.items(), key=lambda kv: -kv[1])[:5]
    return rep


def fingerprint_k(t, k):
    """fingerprint() with k samples (golden_c3.pt keeps 64 per tensor, oracle/make_golden_c3.py)."""
    t = t.detach().double().flatten().cpu()
    idx = torch.linspace(0, t.numel() - 1, k).long()       # k comes from the fixture (duplicates for tiny tensors)
    return {"sum": t.sum().item(), "abs": t.abs().sum().item(), "n": t.numel(), "samples": t[idx].float()}


def sample_agreement(named_tensors, ref_fps, skip=("shortcut.0.bias",)):
    """Value / sign agreement between tensors of the CUDA path and the golden fingerprints' evenly spaced samples.
    Returns the cosine between the two pooled sample vectors (each tensor normalised by the reference's mean |g| so
    that small tensors count), the fraction of significant samples (|ref| above 10 % of the tensor's mean |ref|) whose
    sign agrees, and the fraction of samples within 10 % of |ref| + 10 % of the tensor's mean |ref|.
    Tensors whose exact value is zero (`skip`: conv biases feeding a train-mode BatchNorm) hold rounding noise only."""
    mine, ref, sign_ok, sign_n, val_ok, val_n = [], [], 0, 0, 0, 0
    per = {}
    for name, t in named_tensors:
        if name not in ref_fps or any(name.endswith(s) for s in skip):
            continue
        r = ref_fps[name]
        k = r["samples"].numel()
        o = fingerprint_k(t, k)["samples"].double()
        rs = r["samples"].double()
        scale = r["abs"] / r["n"] + 1e-30
        mine.append(o / scale)
        ref.append(rs / scale)
        sig = rs.abs() > 0.1 * scale
        ok = int((torch.sign(o[sig]) == torch.sign(rs[sig])).sum())
        sign_ok += ok
        sign_n += int(sig.sum())
        v = (o - rs).abs() <= 0.1 * rs.abs() + 0.1 * scale
        val_ok += int(v.sum())
        val_n += k
        per[name] = float(((o - rs).abs().max() / (rs.abs().max() + 1e-30)))
    a, b = torch.cat(mine), torch.cat(ref)
    return {"cosine": float((a @ b) / (a.norm() * b.norm() + 1e-30)), "sign_agree": sign_ok / max(sign_n, 1),
            "n_significant": sign_n, "value_within_10pct": val_ok / max(val_n, 1), "n_samples": val_n,
            "worst_tensor": max(per.items(), key=lambda kv: kv[1])}



class reference_eps:
    """Context manager: LunarisCoreVAE.forward takes its reparameterisation noise from the CPU generator stream the
    reference trainer used for the golden fixtures (`torch.manual_seed(seed)` once, then one `randn_like(std)` per
    step on CPU; with dropout probabilities 0 nothing else draws from it) instead of the CUDA generator."""

    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)

    def __enter__(self):
        from lunaris_orion_b200 import _host
        _host.set_eps_source(lambda b, l, dev: torch.randn(b, l, generator=self.gen))
        return self

    def __exit__(self, *exc):
        from lunaris_orion_b200 import _host
        _host.set_eps_source(None)
        return False
