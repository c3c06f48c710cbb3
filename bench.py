#!/usr/bin/env python
"""Headline benchmark: hybrid VAE + Teacher training step, 128x128x3 synthetic sprites, bf16, images/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2] / configs[3], "C3"): batch 64 per GPU, latent 512 / emb 256 / feat 512, one
`_process_batch` (VAE fwd, Teacher pass A + pass B, two backwards, clip, AdamW, cosine step) per step.
  value : images/s with the step's images already resident in HBM (CUDA-event timing, max over ranks)
  e2e   : same metric through TrainingManager with HOST uint8 sprites: pinned host->device copy + normalise inside
          the timed region and the packed 12 metrics read back every step
  roofline : the dominant kernel (tcgen05 implicit-GEMM 3x3 conv, 512->512 at 128x128) timed alone with CUDA events;
             algorithmic FLOPs per launch = 2 * (B*16384) * 512 * 4608
  cpu_baseline / --impl reference : the oracle port of the reference's CPU path (oracle/restatement.py, torch fp32 on
             all host threads) on a bounded sample of the same architecture
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(batch=64, latent=512, emb=256, feat=512)
GF_PER_IMG = 5443.3   # algorithmic GFLOP per image per step, as-executed, recompute excluded
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel (3x3 512->512, B=64, production
# epilogue) from `ncu --set full` (profiles/r01b_ncu_conv_fprop_3x3_512_B64.txt); algorithmic bytes: 2.152e9
NCU_TRAFFIC = {(512, 64): 1.458739e9 + 1.048327e9}
NCU_TENSOR_PCT = {(512, 64): 95.1}     # sm__pipe_tensor_cycles_active % of elapsed, same capture (99.5 % of active)


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def _args_ns(batch, latent, emb, feat):
    from lunaris_orion_b200.train_hybrid import build_arg_parser
    return build_arg_parser().parse_args([
        "--data_dir", "synthetic", "--batch_size", str(batch), "--gradient_accumulation_steps", "1",
        "--latent_dim", str(latent), "--embedding_dim", str(emb), "--feature_dim", str(feat),
        "--vae_lr", "3e-4", "--teacher_lr", "2e-4"])


# ====================================================================================================== our arm
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from lunaris_orion_b200 import _capi, ops
    from lunaris_orion_b200.train_hybrid import TrainingManager
    from lunaris_orion_b200.lunar_generate import sprites_to_tensor

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != a.gpus:
        if a.gpus != 1:
            raise SystemExit(f"--gpus {a.gpus} needs torchrun with {a.gpus} ranks (WORLD_SIZE={world})")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    B = a.batch
    tm = TrainingManager(_args_ns(B, a.latent, a.emb, a.feat), device=dev)
    lib = _capi.lib()

    rng = np.random.default_rng(1234 + rank)
    host_u8 = torch.from_numpy(rng.integers(0, 256, (4, B, 128, 128, 3), dtype=np.uint8)).pin_memory()
    dev_imgs = [(host_u8[i].to(dev).permute(0, 3, 1, 2).float() / 127.5 - 1.0).contiguous() for i in range(4)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step_fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident(i):
        tm._process_batch(dev_imgs[i % 4], i, return_tensor=True)

    host_metrics = torch.empty(12, dtype=torch.float32).pin_memory()

    def step_e2e(i):
        x = sprites_to_tensor(host_u8[i % 4].to(dev, non_blocking=True))
        m = tm._process_batch(x, i, return_tensor=True)
        host_metrics.copy_(m, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for i in range(a.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.lun_launch_count()
    ms = timed(step_resident, a.steps)
    launches = lib.lun_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    step_e2e(0)
    ms_e2e = timed(step_e2e, a.steps)
    metrics = dict(zip(("recon_loss", "kl_loss", "quality_loss"), host_metrics[:3].tolist()))

    # ---- roofline of the dominant kernel, timed alone on this stream
    roof = None
    if rank == 0:
        C = a.feat
        x = torch.randn(B, 128, 128, C, device=dev).to(torch.bfloat16)
        wp = ops.pack_conv_weight(torch.randn(C, C, 3, 3, device=dev) * 0.02)
        bias = torch.zeros(C, device=dev)
        st = torch.zeros(2 * C, device=dev)
        y = torch.empty(B, 128, 128, C, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.conv2d_fprop(x, wp, 3, 1, 1, bias=bias, act_leaky=True, stats=st, out=y)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            ops.conv2d_fprop(x, wp, 3, 1, 1, bias=bias, act_leaky=True, stats=st, out=y)
        e1.record()
        torch.cuda.synchronize()
        kms = e0.elapsed_time(e1) / 10
        fl = 2.0 * B * 16384 * C * C * 9
        pk = _peaks()
        peak = pk["bf16_tflops"] if pk else 1590.0
        roof = {"bound": "tensor", "kernel": "conv_fprop_kernel 3x3 %d->%d @128x128 B=%d (bias+LeakyReLU+BN stats)" % (C, C, B),
                "achieved": round(fl / kms / 1e9, 1), "peak": peak, "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, burst)" if pk else "fallback",
                "unit": "TFLOP/s", "frac": round(fl / kms / 1e9 / peak, 4), "traffic": NCU_TRAFFIC.get((C, B)),
                "tensor_pipe_active_pct_ncu": NCU_TENSOR_PCT.get((C, B)), "ms_per_launch": round(kms, 4),
                "step_model_flop_frac_of_sustained": None}
        del x, y

    if rank != 0:
        return
    n_img = B * world * a.steps
    value = n_img / (ms / 1e3)
    pk = _peaks()
    if roof is not None and pk:
        roof["step_model_flop_frac_of_sustained"] = round(value / world * GF_PER_IMG / 1e3 / pk["bf16_tflops_sustained"], 4) \
            if (a.feat, a.latent) == (512, 512) else None
    line = {
        "metric": "train images/sec @128x128 bf16 (hybrid VAE+Teacher step)", "value": round(value, 2),
        "unit": "images/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": round(ms / a.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "%s: batch %d/GPU, latent %d, emb %d, feat %d, one _process_batch per step; "
                               "inputs (4 rotating batches) + activations >> L2"
                               % ("C3 high-end" if (B, a.latent, a.emb, a.feat) == (64, 512, 256, 512) else "custom shapes",
                                  B, a.latent, a.emb, a.feat),
                   "global_batch": B * world, "parallelism": "dp%d" % world, "l2": "working set >> 126 MB L2"},
        "e2e": {"value": round(n_img / (ms_e2e / 1e3), 2), "unit": "images/s",
                "h2d_bytes_per_step": B * 128 * 128 * 3, "d2h_bytes_per_step": 48},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "last_metrics": metrics,
    }
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(a, steps=1, warmup=0)
    print(json.dumps(line), flush=True)


# ====================================================================================================== reference arm
def cpu_reference(a, steps, warmup, budget_s=150.0):
    """Oracle port of the reference's CPU path (torch fp32, all host threads): full steps on batch `--cpu-batch`
    (default 4) of the same architecture, at most `budget_s` seconds of them. Returns the cpu_baseline object."""
    import torch
    from oracle import restatement as R
    from lunaris_orion_b200.lunar_evaluator import LunarMoETeacher
    from lunaris_orion_b200.lunar_generate import LunarisCoreVAE
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    vae = LunarisCoreVAE(a.latent)
    teacher = LunarMoETeacher(feature_dim=a.feat, embedding_dim=a.emb, dropout_rate=0.0)

    def leaf_sd(m):
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        for n, _ in m.named_parameters():
            sd[n].requires_grad_(True)
        return sd
    vsd, tsd = leaf_sd(vae), leaf_sd(teacher)
    Bs = a.cpu_batch
    x = torch.rand(Bs, 3, 128, 128) * 2 - 1

    def one():
        for sd in (vsd, tsd):
            for v in sd.values():
                v.grad = None
        R.train_step(x, vsd, tsd, torch.randn(Bs, a.latent))
    t_first = None
    for _ in range(warmup):
        t0 = time.time()
        one()
        t_first = time.time() - t0
    done, t_total = 0, 0.0
    for _ in range(steps):
        t0 = time.time()
        one()
        t_total += time.time() - t0
        done += 1
        if t_total + (t_total / done) > budget_s:
            break
    val = Bs * done / t_total
    return {"value": round(val, 4), "unit": "images/s", "cores": cores, "kind": "port", "timed_steps": done,
            "sample": "oracle/restatement.py train_step (VAE fwd/bwd, Teacher pass A + pass B + bwd) on batch %d of the "
                      "C3 architecture, fp32, %d threads, %.1f s/step" % (Bs, torch.get_num_threads(), t_total / done)}


def eager_gpu_reference(a):
    """SURVEY.md 8(d) "existing kernel" bar: the reference's op sequence (oracle port, loop-free, so faster than the
    reference's own 16k-iteration Python loop) as stock PyTorch eager kernels under bf16 autocast on this B200:
    cuDNN / cuBLAS / ATen, no kernel of this repo. Reported beside, never instead of, the CPU reference arm."""
    import torch
    from oracle import restatement as R
    from lunaris_orion_b200.lunar_evaluator import LunarMoETeacher
    from lunaris_orion_b200.lunar_generate import LunarisCoreVAE
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.manual_seed(42)
    vae = LunarisCoreVAE(a.latent)
    teacher = LunarMoETeacher(feature_dim=a.feat, embedding_dim=a.emb, dropout_rate=0.0)

    def leaf_sd(m):
        sd = {k: v.detach().clone().to(dev) for k, v in m.state_dict().items()}
        for n, _ in m.named_parameters():
            sd[n].requires_grad_(True)
        return sd
    vsd, tsd = leaf_sd(vae), leaf_sd(teacher)
    Bs = a.eager_batch
    x = torch.rand(Bs, 3, 128, 128, device=dev) * 2 - 1
    torch.backends.cudnn.benchmark = True

    def one():
        for sd in (vsd, tsd):
            for v in sd.values():
                v.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            R.train_step(x, vsd, tsd, torch.randn(Bs, a.latent, device=dev))
    for _ in range(max(a.warmup, 2)):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({
        "impl": "reference", "device": "cuda-eager", "metric": "train images/sec @128x128 bf16 (hybrid VAE+Teacher step)",
        "value": round(Bs / ms * 1e3, 2), "unit": "images/s", "n_gpus": 1, "steps": a.steps, "warmup": max(a.warmup, 2),
        "ms_per_step": round(ms, 2), "higher_is_better": True, "dtype": "bf16 autocast", "data": "synthetic",
        "config": {"workload": "C3 architecture (latent %d, emb %d, feat %d), batch %d; oracle port of the reference op "
                               "sequence (no optimizer step, no dropout RNG) on stock PyTorch eager kernels"
                               % (a.latent, a.emb, a.feat, Bs)},
        "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}), flush=True)


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    if a.ref_device == "cuda":
        return eager_gpu_reference(a)
    cb = cpu_reference(a, steps=a.steps, warmup=min(a.warmup, 1))
    line = {
        "impl": "reference", "metric": "train images/sec @128x128 bf16 (hybrid VAE+Teacher step)",
        "value": cb["value"], "unit": "images/s", "n_gpus": a.gpus, "steps": cb["timed_steps"], "warmup": min(a.warmup, 1),
        "ms_per_step": round(1e3 * a.cpu_batch / cb["value"], 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C3 high-end architecture (latent %d, emb %d, feat %d), reference CPU path on host cores; "
                               "each step = one full _process_batch on a bounded sample of batch %d"
                               % (a.latent, a.emb, a.feat, a.cpu_batch)},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--batch", type=int, default=CFG["batch"])
    p.add_argument("--latent", type=int, default=CFG["latent"])
    p.add_argument("--emb", type=int, default=CFG["emb"])
    p.add_argument("--feat", type=int, default=CFG["feat"])
    p.add_argument("--cpu-batch", dest="cpu_batch", type=int, default=4)
    p.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    p.add_argument("--ref-device", dest="ref_device", default="cpu", choices=["cpu", "cuda"],
                   help="--impl reference only: 'cuda' times the same op sequence as stock PyTorch eager kernels on the GPU")
    p.add_argument("--eager-batch", dest="eager_batch", type=int, default=16)
    a = p.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
