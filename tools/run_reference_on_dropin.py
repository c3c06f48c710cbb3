"""INTEGRATION.md 1 end to end: the UNMODIFIED reference trainer (train_hybrid.py from /root/reference or the oracle/_ref
bytecode) with lunaris_orion_b200/dropin ahead of it on sys.path, so its two imports (train_hybrid.py:45-46) bind the
B200-native modules. Runs `--steps` real `_process_batch` calls on cuda:0 and prints one JSON line: module types,
metrics per step, our kernel-launch count, the checkpoint keys it saved.

    python tools/run_reference_on_dropin.py --steps 2 [--dropout 0.0] [--batch 2 --latent 64 --emb 32 --feat 64]
"""
import argparse
import contextlib
import io
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--steps", type=int, default=2)
    p.add_argument("--batch", type=int, default=2)
    p.add_argument("--latent", type=int, default=64)
    p.add_argument("--emb", type=int, default=32)
    p.add_argument("--feat", type=int, default=64)
    p.add_argument("--dropout", type=float, default=None, help="override every dropout probability (0 = deterministic)")
    p.add_argument("--img-seed", type=int, default=5)
    p.add_argument("--eps-seed", type=int, default=123)
    p.add_argument("--cpu-eps", action="store_true",
                   help="take the VAE noise from the CPU generator stream (what the golden fixtures were made with)")
    a = p.parse_args()
    import torch
    from oracle import ref_harness
    d = tempfile.mkdtemp(prefix="lunaris_dropin_")
    ref_harness.write_sprites(os.path.join(d, "data"), 10)
    cfg = dict(B=a.batch, latent=a.latent, emb=a.emb, feat=a.feat, seed=42)
    log = io.StringIO()
    with contextlib.redirect_stdout(log):
        tm = ref_harness.drive_reference_trainer(cfg, os.path.join(d, "data"), os.path.join(d, "out"), device="cuda",
                                                 dropin=True)
    import logging
    logging.disable(logging.CRITICAL)
    import lunar_evaluator
    import lunar_generate
    import train_hybrid
    from lunaris_orion_b200 import _capi
    if a.dropout is not None:
        for m in tm.teacher.modules():
            if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)):
                m.p = a.dropout
    g = torch.Generator().manual_seed(a.img_seed)
    x = (torch.randint(0, 256, (a.batch, 3, 128, 128), generator=g, dtype=torch.uint8).float() / 127.5 - 1.0).to(tm.device)
    l0 = _capi.lib().lun_launch_count()
    torch.manual_seed(a.eps_seed)
    if a.cpu_eps:
        from lunaris_orion_b200 import _host
        gen = torch.Generator().manual_seed(a.eps_seed)
        _host.set_eps_source(lambda b, l, dev: torch.randn(b, l, generator=gen))
    steps = [tm._process_batch(x, i) for i in range(a.steps)]
    launches = _capi.lib().lun_launch_count() - l0
    tm._save_checkpoint()
    ck = torch.load(str(tm.checkpoints_dir / "latest.pt"), weights_only=True)
    print(json.dumps({
        "trainer_module": getattr(train_hybrid, "__file__", "?"),
        "vae_class": f"{type(tm.vae).__module__}.{type(tm.vae).__name__}",
        "teacher_class": f"{type(tm.teacher).__module__}.{type(tm.teacher).__name__}",
        "shim_modules": [lunar_generate.__file__, lunar_evaluator.__file__],
        "optimizer_class": type(tm.vae_optimizer).__name__, "device": str(tm.device),
        "steps": steps, "launches": int(launches), "global_step": int(tm.global_step),
        "vae_lr": tm.vae_optimizer.param_groups[0]["lr"],
        "teacher_none": sum(1 for p in tm.teacher.parameters() if p.grad is None),
        "checkpoint_keys": sorted(ck.keys()), "teacher_state_keys": len(ck["teacher_state_dict"])}))


if __name__ == "__main__":
    main()
