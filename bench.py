#!/usr/bin/env python
"""Headline benchmark: hybrid VAE + Teacher training step, 128x128x3 synthetic sprites, bf16, images/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2] / configs[3], "C3"): batch 64 per GPU, latent 512 / emb 256 / feat 512, dropout on
(reference defaults), one `_process_batch` (VAE fwd, Teacher pass A + pass B, two backwards, clip, AdamW, cosine step)
per step.
  value : images/s with the step's images already resident in HBM (CUDA-event timing, max over ranks)
  e2e   : same metric through the repo's public path: `sprites_*.npy` + `labels_*.csv` on disk -> SpriteLoader
          (memory-mapped gather into pinned buffers, host->device copy, GPU normalise) -> TrainingManager._process_batch
          -> the packed 12 metrics copied back and read on the host every step
  roofline : the dominant kernel (tcgen05 implicit-GEMM 3x3 conv, 512->512 at 128x128) timed alone with CUDA events;
             algorithmic FLOPs per launch = 2 * (B*16384) * 512 * 4608; the same shape through cuDNN beside it
  c2 / c5  : BASELINE.json configs[1] (batch 16, accumulation 4, feat 256) and configs[4] (decoder-only sampling,
             batch 256) measured in the same run (N=1 only)
  dp_check : N>1 only - after the timed region one extra step verifies that the reduced gradient equals the mean of the
             all-gathered per-rank gradients and that every rank holds bit-identical parameters
  cpu_baseline / --impl reference : the UNMODIFIED reference trainer (oracle/_ref, byte-compiled from /root/reference by
             oracle/make_ref.py) on the host cores: real `TrainingManager._process_batch` calls, --force_cpu semantics,
             on a bounded batch of the same architecture
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(batch=64, latent=512, emb=256, feat=512)
GF_PER_IMG = 5443.3   # algorithmic GFLOP per image per step, as-executed, recompute excluded
GF_PER_IMG_C2 = 1423.0
PRE_ROLL_S = 10.0     # untimed steps on top of --warmup: keep stepping until this many seconds have passed
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel (3x3 512->512, B=64, production
# epilogue) from the round's `ncu --set full` capture of tools/prof_conv.py (algorithmic bytes: 2.152e9); a constant
# from that capture, not re-measured by this run
NCU_CAPTURE = {(512, 64): {"traffic": 1.114875e9 + 1.045206e9, "tensor_pipe_active_pct": 97.2,
                           "file": "profiles/r02c_ncu_conv_fprop_and_wgrad_3x3_512_B64.txt"}}


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def executed_gflop_per_image(feat, latent, heads=8):
    """GEMM work the step actually issues per image (dead work pruned): 3x3 / 1x1 convs of the trunk in
    both Teacher passes, the K/V-free attention GEMMs on the surviving rows, the executed backward (conv2 dgrad + wgrad
    and proj wgrad of blocks 1,2, shortcut wgrad), feature extractor, and 3x the VAE forward."""
    C, P = feat, 16384
    nq = P // 32 + 31
    nq_pad = (nq + 7) // 8 * 8
    conv = lambda cin, cout, k: 2.0 * P * cin * cout * k * k
    fold = 2.0 * nq_pad * C * (heads * C) + 2.0 * nq_pad * C * C + 2.0 * nq_pad * C * C + nq * 2 * 2.0 * 32 * heads * C
    fwd_expert = conv(128, C, 3) + conv(128, C, 1) + 5 * conv(C, C, 3) + 3 * fold
    fe = 1.08e9
    fwd = 4 * fwd_expert + fe
    bwd = 4 * (2 * (2 * conv(C, C, 3) + 2.0 * nq_pad * C * C) + conv(128, C, 1))
    vae = 3 * (4.09e9 if latent == 512 else 4.04e9)
    return (2 * fwd + bwd + vae) / 1e9


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during a timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

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
        return self

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
        pw = []
        for r in self.rows:
            try:
                pw.append(float(r[6]))
            except Exception:
                pass
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_median": sorted(pw)[len(pw) // 2] if pw else None}


def _args_ns(batch, latent, emb, feat, data_dir="synthetic", accum=1):
    from lunaris_orion_b200.train_hybrid import build_arg_parser
    return build_arg_parser().parse_args([
        "--data_dir", data_dir, "--batch_size", str(batch), "--gradient_accumulation_steps", str(accum),
        "--latent_dim", str(latent), "--embedding_dim", str(emb), "--feature_dim", str(feat),
        "--vae_lr", "3e-4", "--teacher_lr", "2e-4"])


def _write_sprite_files(d, n, files=2):
    """The reference's on-disk format (generate.py:858-904): sprites_*.npy uint8 NHWC + labels_*.csv, SURVEY 8(d) data."""
    import numpy as np
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(1234)
    per = (n + files - 1) // files
    for k in range(files):
        np.save(os.path.join(d, f"sprites_{k:03d}.npy"), rng.integers(0, 256, (per, 128, 128, 3), dtype=np.uint8))
        with open(os.path.join(d, f"labels_{k:03d}.csv"), "w") as f:
            f.write("filename,category,prompt,seed,pixel_size,guidance_scale,pag_scale,num_steps\n")
            for i in range(per):
                f.write(f"s{k}_{i}.png,cat,prompt,{i},8,7.5,3.0,20\n")


# ====================================================================================================== our arm
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from lunaris_orion_b200 import _capi, ops
    from lunaris_orion_b200.train_hybrid import TrainingManager

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != a.gpus and a.gpus != 1:
        raise SystemExit(f"--gpus {a.gpus} needs torchrun with {a.gpus} ranks (WORLD_SIZE={world})")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = a.batch
    data_dir = os.path.join(tempfile.gettempdir(), "lunaris_bench_sprites_%s" % os.environ.get("MASTER_PORT", "solo"))
    if rank == 0:
        _write_sprite_files(data_dir, int(math.ceil(world * B * 4 / 0.9)) + 2)
    if world > 1:
        dist.barrier()
    tm = TrainingManager(_args_ns(B, a.latent, a.emb, a.feat, data_dir=data_dir), device=dev)
    lib = _capi.lib()

    rng = np.random.default_rng(1234 + rank)
    host_u8 = torch.from_numpy(rng.integers(0, 256, (4, B, 128, 128, 3), dtype=np.uint8))
    dev_imgs = [(host_u8[i].to(dev).permute(0, 3, 1, 2).float() / 127.5 - 1.0).contiguous() for i in range(4)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local).start() if rank == 0 else None
        e0.record()
        for i in range(steps):
            step_fn(i)
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), clocks

    def step_resident(i):
        tm._process_batch(dev_imgs[i % 4], i, return_tensor=True)

    host_metrics = torch.empty(12, dtype=torch.float32).pin_memory()
    batches = tm.train_loader.forever()

    def step_e2e(i):
        x = next(batches)                                  # disk (page cache) -> pinned -> H2D -> normalise kernel
        m = tm._process_batch(x, i, return_tensor=True)
        host_metrics.copy_(m, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(host_metrics[0])                      # the step's result is read on the host

    # W warm-up steps as asked, plus a pre-roll: under the 1000 W cap the SM clock keeps climbing for the first seconds
    # of a run (medians 1462 -> 1597 -> 1620 MHz over three consecutive 5-step loops of one process), which made the
    # first timed loop 2-5 % slower than every later one - round 1's "e2e > value". After ~10 s the three loops agree to
    # 1 % (246.6 / 245.2 / 247.5 images/s); the A-B-A loop below bounds what is left.
    for i in range(a.warmup):
        step_resident(i)
    torch.cuda.synchronize()
    t_w = time.time()
    for i in range(2):                  # two settled steps give the step time (the W warm-up steps include one-off costs)
        step_resident(i)
    torch.cuda.synchronize()
    # number of pre-roll steps from rank 0's timing, the same on every rank (each step holds collectives)
    n_pre = torch.tensor([min(60, max(3, int(PRE_ROLL_S / max((time.time() - t_w) / 2, 1e-3))))], device=dev)
    if world > 1:
        dist.broadcast(n_pre, 0)
    pre_roll = 2 + int(n_pre.item())
    for i in range(pre_roll - 2):
        step_resident(i)
    l0 = lib.lun_launch_count()
    ms, clocks = timed(step_resident, a.steps)
    launches = lib.lun_launch_count() - l0
    step_e2e(0)
    ms_e2e, clocks_e2e = timed(step_e2e, a.steps)
    metrics = dict(zip(("recon_loss", "kl_loss", "quality_loss"), host_metrics[:3].tolist()))
    ms_again, clocks_again = timed(step_resident, a.steps)         # A-B-A: drift / ordering between the two loops

    # ---- data-parallel check (one extra, untimed step)
    dp_check = None
    if world > 1:
        dp_check = _dp_check(tm, dev_imgs[0], world, dev)

    # ---- roofline of the dominant kernel, timed alone on this stream, with cuDNN on the same shape beside it
    roof = None
    if rank == 0:
        C = a.feat
        x = torch.randn(B, 128, 128, C, device=dev).to(torch.bfloat16)
        w = torch.randn(C, C, 3, 3, device=dev) * 0.02
        wp = ops.pack_conv_weight(w)
        bias = torch.zeros(C, device=dev)
        st = torch.zeros(2 * C, device=dev)
        y = torch.empty(B, 128, 128, C, device=dev, dtype=torch.bfloat16)

        def time_it(fn, n=10):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        kms = time_it(lambda: ops.conv2d_fprop(x, wp, 3, 1, 1, bias=bias, act_leaky=True, stats=st, out=y))
        fl = 2.0 * B * 16384 * C * C * 9
        pk = _peaks()
        peak = pk["bf16_tflops"] if pk else 1590.0
        cap = NCU_CAPTURE.get((C, B), {})
        roof = {"bound": "tensor",
                "kernel": "conv_fprop_kernel 3x3 %d->%d @128x128 B=%d (bias+LeakyReLU+BN stats)" % (C, C, B),
                "achieved": round(fl / kms / 1e9, 1), "peak": peak,
                "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, burst)" if pk else "fallback",
                "unit": "TFLOP/s", "frac": round(fl / kms / 1e9 / peak, 4), "ms_per_launch": round(kms, 4),
                "traffic": cap.get("traffic"), "traffic_source": cap.get("file"),
                "tensor_pipe_active_pct_ncu": cap.get("tensor_pipe_active_pct")}
        if world == 1 and not a.no_extras:
            try:                                           # cuDNN: the reference's kernel for this call site
                import torch.nn.functional as F
                torch.backends.cudnn.benchmark = True
                xc = x.permute(0, 3, 1, 2)                 # NHWC storage = channels_last
                wc = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
                bc = bias.to(torch.bfloat16)
                cms = time_it(lambda: F.conv2d(xc, wc, bc, padding=1))
                roof["vs_cudnn"] = {"cudnn_tflops": round(fl / cms / 1e9, 1), "cudnn_ms": round(cms, 4),
                                    "ratio": round(cms / kms, 3),
                                    "note": "torch F.conv2d bf16 channels_last, cudnn.benchmark on, bias only (no "
                                            "LeakyReLU / BN statistics, which ours computes in the same launch)"}
            except Exception as e:                         # pragma: no cover
                roof["vs_cudnn"] = {"error": str(e)[:200]}
        del x, y

    extras = {}
    if world == 1 and not a.no_extras:
        extras["c5"] = _bench_c5(tm, dev)
        extras["c2"] = _bench_c2(dev)
        if (a.feat, B) == (512, 64):
            try:
                extras["hbm_passes"] = _bench_hbm_passes(dev, B, a.feat)
            except Exception as e:                         # pragma: no cover
                extras["hbm_passes"] = {"error": str(e)[:200]}

    if rank != 0:
        return
    n_img = B * world * a.steps
    value = n_img / (ms / 1e3)
    pk = _peaks()
    if roof is not None and pk and (a.feat, a.latent) == (512, 512):
        sus = pk["bf16_tflops_sustained"]
        per_gpu = value / world
        ex = executed_gflop_per_image(a.feat, a.latent)
        roof["step_algorithmic_gflop_per_img"] = GF_PER_IMG
        roof["step_executed_gflop_per_img"] = round(ex, 1)
        # algorithmic: the reference's op sequence incl. the dead work this path prunes (NOT a utilisation);
        # executed: GEMM FLOPs actually issued / sustained cuBLAS rate
        roof["step_algorithmic_flop_frac_of_sustained"] = round(per_gpu * GF_PER_IMG / 1e3 / sus, 4)
        roof["step_executed_flop_frac_of_sustained"] = round(per_gpu * ex / 1e3 / sus, 4)
    line = {
        "metric": "train images/sec @128x128 bf16 (hybrid VAE+Teacher step)", "value": round(value, 2),
        "unit": "images/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": round(ms / a.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "%s: batch %d/GPU, latent %d, emb %d, feat %d, one _process_batch per step; "
                               "inputs (4 rotating batches) + activations >> L2"
                               % ("C3 high-end" if (B, a.latent, a.emb, a.feat) == (64, 512, 256, 512) else "custom shapes",
                                  B, a.latent, a.emb, a.feat),
                   "global_batch": B * world, "parallelism": "dp%d" % world, "l2": "working set >> 126 MB L2",
                   "untimed_steps_before_timing": a.warmup + pre_roll},
        "e2e": {"value": round(n_img / (ms_e2e / 1e3), 2), "unit": "images/s",
                "h2d_bytes_per_step": tm.train_loader.h2d_bytes_per_batch, "d2h_bytes_per_step": 48,
                "ms_per_step": round(ms_e2e / a.steps, 3), "clocks": clocks_e2e,
                "path": "sprites_*.npy on disk -> SpriteLoader (mmap gather, pinned ring, copy stream) -> "
                        "lun_sprites_u8_to_f32 -> TrainingManager._process_batch -> 12 metrics read on the host"},
        "value_repeat_after_e2e": {"value": round(n_img / (ms_again / 1e3), 2), "ms_per_step": round(ms_again / a.steps, 3),
                                   "clocks": clocks_again,
                                   "note": "same resident loop run again after the e2e loop (A-B-A): the spread "
                                           "between the two resident runs bounds what ordering / clock drift explains"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "last_metrics": metrics,
    }
    if dp_check is not None:
        line["dp_check"] = dp_check
    line.update(extras)
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(a, steps=1, warmup=1, batch=a.cpu_batch_sample, budget_s=60.0)
    print(json.dumps(line), flush=True)


def _dp_check(tm, images, world, dev):
    """SURVEY.md 8(e) parity: DP gradients == mean over ranks of the per-rank gradients; identical parameters."""
    import torch
    import torch.distributed as dist
    res = {}

    def after_reduce(t):
        worst, n = 0.0, 0
        for i, flat in enumerate(t.reducer.flat):
            local = t.reducer.local[i]
            gathered = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(gathered, local)
            mean = torch.stack(gathered).double().mean(0)
            got = flat.double() / world                    # buckets hold SUMS; the optimizer folds 1/world in
            scale = mean.abs().max().item() + 1e-30
            worst = max(worst, (got - mean).abs().max().item() / scale)
            n += len(t.reducer.buckets[i])
            differs = (gathered[0] - gathered[-1]).abs().max().item() > 0
            res["ranks_hold_different_local_grads"] = res.get("ranks_hold_different_local_grads", False) or differs
        res["grad_vs_mean_of_rank_grads_max_rel"] = worst
        res["tensors_checked"] = n
    tm.reducer.keep_local = True
    tm._after_reduce = after_reduce
    tm._process_batch(images, 0, return_tensor=True)
    tm.reducer.keep_local = False
    tm._after_reduce = None
    tm.reducer.local.clear()
    worst = torch.zeros(1, device=dev)
    n = 0
    for m in (tm.vae, tm.teacher):
        for p in m.parameters():
            ref = p.detach().clone()
            dist.broadcast(ref, 0)
            worst = torch.maximum(worst, (p.detach() - ref).abs().max().reshape(1))
            n += 1
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    res["max_abs_param_diff_vs_rank0"] = float(worst.item())
    res["params_bit_identical_across_ranks"] = float(worst.item()) == 0.0
    res["param_tensors_compared"] = n
    res["ok"] = bool(res["params_bit_identical_across_ranks"] and res["grad_vs_mean_of_rank_grads_max_rel"] < 1e-5
                     and res["ranks_hold_different_local_grads"])
    return res


def _bench_c5(tm, dev):
    """BASELINE.json configs[4]: decoder-only sampling (lunar_generate.py:278-291), batch 256, latent 512."""
    import torch
    n, iters = 256, 200                      # 0.15 s of sampling: 20 iterations (15 ms) swung by +-8 % with the clock state
    with torch.no_grad():
        for _ in range(20):
            tm.vae.sample(n)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            tm.vae.sample(n)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ips = n / ms * 1e3
    pk = _peaks() or {"bf16_tflops_sustained": 1331.6, "hbm_gbs": 6556.2}
    return {"metric": "decoder-only sampling images/sec (batch 256, latent %d)" % tm.vae.latent_dim,
            "value": round(ips, 1), "unit": "images/s", "ms_per_batch": round(ms, 4),
            "tensor_frac_of_sustained": round(ips * 1.136e9 / 1e12 / pk["bf16_tflops_sustained"], 4),
            "hbm_frac_of_peak": round(ips * 4.3e6 / 1e9 / pk["hbm_gbs"], 4),
            "model": "1.136 GFLOP and ~4.3 MB of compulsory traffic per image (SURVEY.md 8d); includes torch.randn of z"}


def _bench_hbm_passes(dev, B, C):
    """The bandwidth-bound Teacher passes at the step's shapes ([B,16384,C] bf16), each timed ALONE with CUDA events on
    the launching stream against the measured copy peak: compulsory bytes (tensors read + written once) / time."""
    import ctypes
    import torch
    from lunaris_orion_b200 import _capi, lunar_evaluator as le
    lib = _capi.lib()
    HW = 16384
    pk = _peaks() or {"hbm_gbs": 6556.2}
    x, idn, x2 = (torch.randn(B, HW, C, device=dev).to(torch.bfloat16) for _ in range(3))
    sc = torch.rand(C, device=dev) + 0.5
    sh = torch.randn(C, device=dev)
    ls = torch.full((C,), 0.1, device=dev)
    m2 = (torch.rand(B, C, device=dev) > 0.1).float() * 1.109375
    pool = torch.zeros(B, C, device=dev)
    mean, rstd, gamma = torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.ones(C, device=dev)
    gb = x.numel() * 2 / 1e9
    s0 = torch.cuda.current_stream().cuda_stream

    def timed(fn, iters=8):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    out = {}

    def put(name, ms, nbytes_gb, what):
        out[name] = {"ms": round(ms, 4), "gbytes": round(nbytes_gb, 3), "achieved_gbs": round(nbytes_gb / ms * 1e3, 1),
                     "frac_of_hbm_peak": round(nbytes_gb / ms * 1e3 / pk["hbm_gbs"], 4), "traffic": what}
    ms = timed(lambda: le._affine(x, B, HW, C, sc, sh, mask2d=m2, ls=ls, identity=idn, pool=pool))
    put("block_tail_fwd", ms, 3 * gb, "BN apply + Dropout2d + layer scale + residual + leaky_relu + pooling: 2 read, 1 written")
    t = torch.zeros(2, C, device=dev)
    dpre, dz, dbias = torch.empty_like(x), torch.empty_like(x), torch.zeros(C, device=dev)
    ms = timed(lambda: _capi.check(lib.lun_block_bwd_reduce_bf16(
        x.data_ptr(), None, idn.data_ptr(), x2.data_ptr(), mean.data_ptr(), rstd.data_ptr(), m2.data_ptr(),
        dpre.data_ptr(), t[0].data_ptr(), t[1].data_ptr(), B, HW, C, ctypes.c_float(0.2), s0), "reduce"))
    put("block_tail_bwd_reduce", ms, 4 * gb, "3 read, 1 written")
    ms = timed(lambda: _capi.check(lib.lun_block_bwd_apply_bf16(
        dpre.data_ptr(), None, None, x2.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), ls.data_ptr(),
        m2.data_ptr(), t[0].data_ptr(), t[1].data_ptr(), dz.data_ptr(), dbias.data_ptr(), B, HW, C,
        ctypes.c_float(0.2), ctypes.c_float(0.2), s0), "apply"))
    put("block_tail_bwd_apply", ms, 3 * gb, "2 read, 1 written")
    nq = HW // 32 + 31
    nq_pad = (nq + 7) // 8 * 8
    qt = torch.randn(B, nq_pad, 8 * C, device=dev).to(torch.bfloat16) * 0.05
    xbar = torch.empty_like(qt)
    ms = timed(lambda: _capi.check(lib.lun_attn_fold_rows_bf16(
        x.data_ptr(), sc.data_ptr(), sh.data_ptr(), m2.data_ptr(), qt.data_ptr(), xbar.data_ptr(), B, HW, C, 8, nq_pad,
        12345, ctypes.c_float(0.1), s0), "fold"))
    put("attn_fold", ms, gb + 2 * qt.numel() * 2 / 1e9, "attention input read once + folded queries / outputs")
    small = torch.zeros(B, nq_pad, C, device=dev, dtype=torch.bfloat16)
    bias = torch.zeros(C, device=dev)
    ms = timed(lambda: _capi.check(lib.lun_proj_expand_bf16(
        small.data_ptr(), bias.data_ptr(), dz.data_ptr(), B, HW, C, nq, nq_pad, 1, ctypes.c_float(0.1), s0), "expand"))
    put("proj_expand", ms, gb, "write-only stream with one dropout hash per two elements")
    st = torch.zeros(2 * C, device=dev)
    ms = timed(lambda: _capi.check(lib.lun_channel_stats_bf16(x.data_ptr(), B * HW, C, st.data_ptr(), s0), "stats"))
    put("channel_stats", ms, gb, "1 read")
    out["peak_gbs"] = pk["hbm_gbs"]
    out["peak_source"] = "MEASURED_PEAKS.json hbm_gbs (torch copy_, read + write bytes)"
    return out


def _bench_c2(dev):
    """BASELINE.json configs[1]: batch 16, latent 256 / emb 128 / feat 256, --gradient_accumulation_steps 4."""
    import torch
    from lunaris_orion_b200.train_hybrid import TrainingManager
    B, accum = 16, 4
    tm = TrainingManager(_args_ns(B, 256, 128, 256, accum=accum), device=dev)
    xs = [torch.rand(B, 3, 128, 128, device=dev) * 2 - 1 for _ in range(4)]
    for i in range(2 * accum):
        tm._process_batch(xs[i % 4], i, return_tensor=True)
    calls = 8 * accum
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(calls):
        tm._process_batch(xs[i % 4], i, return_tensor=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / calls
    ips = B / ms * 1e3
    pk = _peaks() or {"bf16_tflops_sustained": 1331.6}
    return {"metric": "train images/sec, batch 16, accumulation 4, feat 256 (every micro-batch counted)",
            "value": round(ips, 1), "unit": "images/s", "ms_per_micro_batch": round(ms, 3),
            "algorithmic_flop_frac_of_sustained": round(ips * GF_PER_IMG_C2 / 1e3 / pk["bf16_tflops_sustained"], 4)}


# ====================================================================================================== reference arm
def cpu_reference(a, steps, warmup, batch, budget_s=150.0):
    """The reference's own CPU path on the host cores. With oracle/_ref (or /root/reference) present: the UNMODIFIED
    reference TrainingManager built by its own main() (--force_cpu, dropout on, AdamW, schedulers; SURVEY.md App. C.1
    harness) - every step is one real `_process_batch` on a bounded batch of the benchmarked architecture.
    Otherwise (kind "port") the oracle restatement. Returns the cpu_baseline object."""
    import torch
    from oracle import reference_loader
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if reference_loader.available():
        from oracle import ref_harness
        d = tempfile.mkdtemp(prefix="lunaris_ref_")
        ref_harness.write_sprites(os.path.join(d, "data"), max(10, int(math.ceil(batch / 0.9)) + 1))
        cfg = dict(B=batch, latent=a.latent, emb=a.emb, feat=a.feat, seed=42, vae_lr=3e-4, teacher_lr=2e-4)
        import contextlib
        import io
        import logging
        with contextlib.redirect_stdout(io.StringIO()):
            tm = ref_harness.drive_reference_trainer(cfg, os.path.join(d, "data"), os.path.join(d, "out"))
        logging.disable(logging.CRITICAL)
        g = torch.Generator().manual_seed(7)
        x = torch.randint(0, 256, (batch, 3, 128, 128), generator=g, dtype=torch.uint8).float() / 127.5 - 1.0
        one = lambda i: tm._process_batch(x.clone(), i)
        kind = "reference"
        what = ("UNMODIFIED reference train_hybrid.TrainingManager._process_batch (--force_cpu, fp32, dropout on, clip "
                "+ AdamW + scheduler) from %s" % ("oracle/_ref bytecode" if reference_loader.kind() == "bytecode"
                                                   else reference_loader.REF))
    else:
        from oracle import restatement as R
        from lunaris_orion_b200.lunar_evaluator import LunarMoETeacher
        from lunaris_orion_b200.lunar_generate import LunarisCoreVAE
        torch.manual_seed(42)
        vae = LunarisCoreVAE(a.latent)
        teacher = LunarMoETeacher(feature_dim=a.feat, embedding_dim=a.emb, dropout_rate=0.0)

        def leaf_sd(m):
            sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
            for n, _ in m.named_parameters():
                sd[n].requires_grad_(True)
            return sd
        vsd, tsd = leaf_sd(vae), leaf_sd(teacher)
        x = torch.rand(batch, 3, 128, 128) * 2 - 1

        def one(i):
            for sd in (vsd, tsd):
                for v in sd.values():
                    v.grad = None
            R.train_step(x, vsd, tsd, torch.randn(batch, a.latent))
        kind = "port"
        what = "oracle/restatement.py train_step (reference unavailable on this box: oracle/_ref missing)"
    for i in range(warmup):
        one(i)
    done, t_total = 0, 0.0
    for i in range(steps):
        t0 = time.time()
        one(warmup + i)
        t_total += time.time() - t0
        done += 1
        if t_total + (t_total / done) > budget_s:
            break
    val = batch * done / t_total
    return {"value": round(val, 4), "unit": "images/s", "cores": cores, "kind": kind, "timed_steps": done,
            "warmup_steps": warmup, "batch": batch,
            "sample": "%s; %d timed step(s) after %d warm-up on batch %d of the architecture latent %d / emb %d / feat "
                      "%d, %d threads, %.1f s/step" % (what, done, warmup, batch, a.latent, a.emb, a.feat,
                                                        torch.get_num_threads(), t_total / done)}


def eager_gpu_reference(a):
    """SURVEY.md 8(d) "existing kernel" bar: the UNMODIFIED reference trainer on this B200 as stock PyTorch eager
    kernels (cuDNN / cuBLAS / ATen, its own 16k-iteration attention loop), bf16 autocast recipe of SURVEY 8(c):
    `torch.set_autocast_dtype('cuda', bf16)` before the reference modules are imported (their decorators freeze the
    dtype at import) and an outer bf16 autocast around `_process_batch`. Never the driver's arm."""
    import contextlib
    import io
    import logging
    import torch
    from oracle import ref_harness, reference_loader
    if not reference_loader.available():
        print(json.dumps({"impl": "reference", "device": "cuda-eager", "unavailable": "oracle/_ref missing"}))
        return
    torch.set_autocast_dtype("cuda", torch.bfloat16)
    Bs = a.eager_batch
    d = tempfile.mkdtemp(prefix="lunaris_ref_")
    ref_harness.write_sprites(os.path.join(d, "data"), int(math.ceil(Bs / 0.9)) + 2)
    cfg = dict(B=Bs, latent=a.latent, emb=a.emb, feat=a.feat, seed=42, vae_lr=3e-4, teacher_lr=2e-4)
    with contextlib.redirect_stdout(io.StringIO()):
        tm = ref_harness.drive_reference_trainer(cfg, os.path.join(d, "data"), os.path.join(d, "out"), device="cuda")
    logging.disable(logging.CRITICAL)
    dev = tm.device
    x = torch.rand(Bs, 3, 128, 128, device=dev) * 2 - 1

    def one(i):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            tm._process_batch(x, i)
    for i in range(max(min(a.warmup, 2), 1)):
        one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        one(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({
        "impl": "reference", "device": "cuda-eager", "kind": "reference",
        "metric": "train images/sec @128x128 bf16 (hybrid VAE+Teacher step)",
        "value": round(Bs / ms * 1e3, 2), "unit": "images/s", "n_gpus": 1, "steps": a.steps,
        "ms_per_step": round(ms, 2), "higher_is_better": True, "dtype": "bf16 autocast", "data": "synthetic",
        "config": {"workload": "C3 architecture (latent %d, emb %d, feat %d), batch %d; UNMODIFIED reference "
                               "TrainingManager._process_batch on stock PyTorch eager kernels" % (a.latent, a.emb, a.feat, Bs)},
        "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}), flush=True)


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    if a.ref_config == "c1":        # BASELINE configs[0], the reference's own CPU-runnable case
        a.latent, a.emb, a.feat, a.cpu_batch, a.eager_batch = 256, 128, 256, 8, 8
    if a.ref_device == "cuda":
        return eager_gpu_reference(a)
    cb = cpu_reference(a, steps=a.steps, warmup=min(a.warmup, 1), batch=a.cpu_batch, budget_s=a.ref_budget)
    line = {
        "impl": "reference", "metric": "train images/sec @128x128 bf16 (hybrid VAE+Teacher step)",
        "value": cb["value"], "unit": "images/s", "n_gpus": a.gpus, "steps": cb["timed_steps"],
        "warmup": cb["warmup_steps"], "ms_per_step": round(1e3 * a.cpu_batch / cb["value"], 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": ("C1 (BASELINE configs[0]: batch %d, latent %d, emb %d, feat %d), reference CPU path on host "
                                "cores, full batch; steps stop at a %d s budget"
                                % (a.cpu_batch, a.latent, a.emb, a.feat, int(a.ref_budget))) if a.ref_config == "c1" else
                   ("C3 high-end architecture (latent %d, emb %d, feat %d), reference CPU path on host cores; "
                    "each step = one full _process_batch on a bounded sample of batch %d (a batch-64 step "
                    "of the reference takes about nine minutes on 16 cores; steps stop at a %d s budget)"
                    % (a.latent, a.emb, a.feat, a.cpu_batch, int(a.ref_budget)))},
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
    p.add_argument("--cpu-batch", dest="cpu_batch", type=int, default=4,
                   help="--impl reference: batch of each reference step (bounded sample of the 64-image workload)")
    p.add_argument("--cpu-batch-sample", dest="cpu_batch_sample", type=int, default=1,
                   help="batch of the cpu_baseline object inside our own line (about 10-30 s of CPU work)")
    p.add_argument("--ref-budget", dest="ref_budget", type=float, default=170.0,
                   help="--impl reference: stop timing new steps once this many seconds are spent")
    p.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    p.add_argument("--no-extras", dest="no_extras", action="store_true", help="skip the c2 / c5 / vs_cudnn sub-runs")
    p.add_argument("--ref-device", dest="ref_device", default="cpu", choices=["cpu", "cuda"],
                   help="--impl reference only: 'cuda' times the unmodified reference as stock PyTorch eager kernels on the GPU")
    p.add_argument("--eager-batch", dest="eager_batch", type=int, default=16)
    p.add_argument("--ref-config", dest="ref_config", default="c3", choices=["c3", "c1"],
                   help="--impl reference only: c3 = our arm's architecture on a bounded batch (default, what the driver "
                        "compares); c1 = BASELINE configs[0] at its full batch of 8 (SURVEY.md 8d)")
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
