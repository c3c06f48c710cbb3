"""B200-native hybrid trainer with the reference's CLI, step semantics and checkpoint layout
(MeryylleA/Lunaris-Orion train_hybrid.py:230-1147), plus batch-sharded data parallelism (one process per GPU,
bucketed NCCL gradient all-reduce overlapped with the second backward pass) that the reference lacks.

`TrainingManager._process_batch` reproduces train_hybrid.py:838-954 as executed: grads are zeroed on every
micro-batch (so --gradient_accumulation_steps k only scales the k-th micro-batch's gradient by 1/k),
the Teacher runs twice (no-grad pass A on the images, pass B on recon.detach()), two backward passes, clip + AdamW +
per-step cosine warm restarts on accumulation boundaries. The reward baseline lives on the device (same arithmetic as
the reference's python float), so a step has a single device->host copy: the packed 12 metrics.
"""
import argparse
import os
import time
import warnings

import numpy as np
import torch
import torch.distributed as dist
from torch.optim.lr_scheduler import CosineAnnealingWarmRestarts

from ._capi import LunarisB200Error
from .data import PixelArtDataset, SpriteLoader, SyntheticSprites, split_indices
from .lunar_evaluator import LunarMoETeacher
from .lunar_generate import LunarisCoreVAE, vae_losses
from .optim import ClipAdamW

METRIC_KEYS = ("recon_loss", "kl_loss", "quality_loss", "pg_loss", "semantic_reward", "quality_reward", "baseline",
               "advantage", "vae_loss", "teacher_loss", "total_loss", "quality_scores")


def build_arg_parser():
    """Same 35 flags as the reference (train_hybrid.py:1076-1133); --data_dir may be 'synthetic'."""
    p = argparse.ArgumentParser(description="Hybrid Training for Lunaris (B200-native): Generator and Evaluator")
    p.add_argument('--data_dir', type=str, required=True)
    p.add_argument('--output_dir', type=str, default='output')
    p.add_argument('--resume_from', type=str)
    p.add_argument('--batch_size', type=int, default=16)
    p.add_argument('--gradient_accumulation_steps', type=int, default=2)
    p.add_argument('--chunk_size', type=int, default=32)
    p.add_argument('--num_epochs', type=int, default=100)
    p.add_argument('--num_workers', type=int, default=4)
    p.add_argument('--seed', type=int, default=42)
    p.add_argument('--compile', action='store_true')
    p.add_argument('--mixed_precision', action='store_true')
    p.add_argument('--latent_dim', type=int, default=256)
    p.add_argument('--embedding_dim', type=int, default=64)
    p.add_argument('--feature_dim', type=int, default=128)
    p.add_argument('--num_experts', type=int, default=4)
    p.add_argument('--vae_lr', type=float, default=1e-4)
    p.add_argument('--teacher_lr', type=float, default=1e-4)
    p.add_argument('--min_lr', type=float, default=1e-6)
    p.add_argument('--weight_decay', type=float, default=0.01)
    p.add_argument('--max_grad_norm', type=float, default=1.0)
    p.add_argument('--scheduler_t0', type=int, default=10)
    p.add_argument('--recon_weight', type=float, default=1.0)
    p.add_argument('--kl_weight', type=float, default=0.1)
    p.add_argument('--quality_weight', type=float, default=0.5)
    p.add_argument('--log_every', type=int, default=100)
    p.add_argument('--save_every', type=int, default=1000)
    p.add_argument('--sample_every', type=int, default=500)
    p.add_argument('--keep_n_checkpoints', type=int, default=5)
    p.add_argument('--early_stopping_patience', type=int, default=7)
    p.add_argument('--eval_save_freq', type=int, default=500)
    p.add_argument('--reward_scale', type=float, default=0.1)
    p.add_argument('--semantic_weight', type=float, default=0.5)
    p.add_argument('--baseline_momentum', type=float, default=0.9)
    p.add_argument('--force_cpu', action='store_true')
    p.add_argument('--memory_efficient', action='store_true')
    return p


# ====================================================================================================== data parallel
class GradBucketReducer:
    """Bucketed gradient all-reduce launched from grad-ready hooks on a side stream (the reference has no data
    parallelism; SURVEY.md 8e defines it: batch-sharded ranks, one gradient average per optimizer step).

    `params` is the STATIC live set in the order gradients become ready (for the hybrid step: every VAE tensor, then
    the Teacher tensors of the reference's executed gradient set). They are packed into ~bucket_mb buckets, each with a
    persistent flat fp32 buffer. When the last member of a bucket gets its gradient, one multi-tensor copy moves the
    members' gradients into the flat buffer, `p.grad` is re-pointed at its slice (so the optimizer reads the reduced
    values in place: no concatenation, no write-back copies), and the bucket's all-reduce (SUM) starts on the side
    stream, overlapping the rest of the backward. The 1/world average is folded into the optimizer
    (`ClipAdamW.grad_scale`); `finish(average=True)` scales here instead for callers without it."""

    def __init__(self, params, group=None, bucket_mb=32):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        for p in self.params:
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_mb * (1 << 20):
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self.flat, self.views = [], {}
        for b in self.buckets:
            flat = torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=b[0].device)
            off = 0
            for p in b:
                self.views[id(p)] = flat[off:off + p.numel()].view(p.shape)
                off += p.numel()
            self.flat.append(flat)
        self.ready = [0] * len(self.buckets)
        self.inflight = {}
        self.enabled = True
        self.keep_local = False            # dp_check: keep a copy of every bucket's local gradients
        self.local = {}
        self.cuda = bool(self.params) and self.params[0].is_cuda
        self.stream = torch.cuda.Stream() if self.cuda else None
        self.hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params] \
            if self.world > 1 else []

    def _on_grad(self, p):
        if not self.enabled:
            return
        i = self.bucket_of[id(p)]
        self.ready[i] += 1
        if self.ready[i] == len(self.buckets[i]):
            self._launch(i)

    def _launch(self, i):
        if i in self.inflight:
            return
        have = [p for p in self.buckets[i] if p.grad is not None]
        if not have:
            return
        flat = self.flat[i]
        if len(have) < len(self.buckets[i]):
            flat.zero_()                   # members without a gradient on this rank contribute zeros
        torch._foreach_copy_([self.views[id(p)] for p in have], [p.grad for p in have])
        for p in have:
            p.grad = self.views[id(p)]
        if self.keep_local:
            self.local[i] = flat.clone()
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.inflight[i] = work

    def finish(self, average=False):
        """Start whatever did not fire from a hook, wait for every bucket. After this `p.grad` of every live parameter
        is a view of the summed (average=True: averaged) gradients."""
        if self.world == 1:
            return
        for i in range(len(self.buckets)):
            self._launch(i)
        for i, work in sorted(self.inflight.items()):
            work.wait()
        if self.cuda and self.inflight:
            torch.cuda.current_stream().wait_stream(self.stream)
        if average:
            for i in self.inflight:
                self.flat[i].mul_(1.0 / self.world)
        self.inflight.clear()
        self.ready = [0] * len(self.buckets)


def live_parameters(vae, teacher):
    """The tensors that receive gradients in the hybrid step, in the order they become ready: all 72 VAE tensors
    (vae_loss.backward()), then - from teacher_loss.backward() - the gate and quality heads and the trunk tensors of
    the reference's executed gradient set."""
    from .lunar_evaluator import _trunk_grad_params
    heads = list(teacher.gate.parameters()) + list(teacher.quality_heads.parameters())
    return list(vae.parameters()) + heads + _trunk_grad_params(teacher)


class EarlyStopping:
    """Patience counter on the epoch loss (train_hybrid.py:206-225)."""

    def __init__(self, patience=7, min_delta=0.0):
        self.patience, self.min_delta = patience, min_delta
        self.counter, self.best_loss, self.early_stop = 0, None, False

    def __call__(self, loss):
        if self.best_loss is None:
            self.best_loss = loss
        elif loss > self.best_loss + self.min_delta:
            self.counter += 1
            if self.counter >= self.patience:
                self.early_stop = True
        else:
            self.best_loss = loss
            self.counter = 0


# flags of the reference CLI that this trainer parses (same surface) but that cannot change its behaviour
_INERT_FLAGS = {
    "mixed_precision": "the B200 path always computes in bf16 with fp32 accumulation (the reference's switch selects "
                       "fp16 autocast + GradScaler)",
    "compile": "the step already runs hand-written kernels; there is nothing for torch.compile to do",
    "memory_efficient": "parsed and ignored by the reference too (train_hybrid.py:1132)",
    "num_workers": "batches are gathered by one prefetch thread, not by DataLoader worker processes",
    "chunk_size": "parsed and ignored by the reference too; the attention chunk is fixed at 32 (lunar_evaluator.py:148)",
    "save_every": "parsed and ignored by the reference too", "sample_every": "parsed and ignored by the reference too",
    "keep_n_checkpoints": "parsed and ignored by the reference too",
    "eval_save_freq": "the PNG eval-sample dump (train_hybrid.py:718-789) is outside the accelerated path",
}


# ====================================================================================================== trainer
class TrainingManager:
    """Drop-in for the reference's TrainingManager on the hot path: models, optimizers, schedulers, data, one step,
    checkpoint save / load (train_hybrid.py:382-404, 502-585, 594-615, 791-836, 838-954)."""

    def __init__(self, args, device=None):
        self.args = args
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        if getattr(args, "force_cpu", False):
            raise LunarisB200Error("--force_cpu: lunaris_orion_b200 has no CPU path (use the reference trainer on CPU)")
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("lunaris_orion_b200 needs a CUDA device (B200, sm_100a): there is no CPU path")
            device = torch.device("cuda", self.local_rank)
            torch.cuda.set_device(device)
        self.device = device
        if self.rank == 0:
            defaults = build_arg_parser().parse_args(["--data_dir", "x"])
            for flag, why in _INERT_FLAGS.items():
                if getattr(args, flag, None) != getattr(defaults, flag):
                    warnings.warn(f"--{flag} has no effect in lunaris_orion_b200: {why}", stacklevel=2)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=device)
        # identical initial weights on every rank: seed before construction, VAE first (train_hybrid.py:1138, 394, 400)
        torch.manual_seed(args.seed)
        np.random.seed(args.seed)
        torch.cuda.manual_seed_all(args.seed)
        self.vae = LunarisCoreVAE(latent_dim=args.latent_dim).to(device).train()
        self.teacher = LunarMoETeacher(num_experts=args.num_experts, feature_dim=args.feature_dim,
                                       embedding_dim=args.embedding_dim).to(device).train()
        # clip_grad_norm_ + AdamW fused into two multi-tensor launches per model (state layout == torch.optim.AdamW)
        mk = dict(weight_decay=args.weight_decay, betas=(0.9, 0.999), max_grad_norm=args.max_grad_norm)
        self.vae_optimizer = ClipAdamW(self.vae.parameters(), lr=args.vae_lr, **mk)
        self.teacher_optimizer = ClipAdamW(self.teacher.parameters(), lr=args.teacher_lr, **mk)
        sk = dict(T_0=args.scheduler_t0, T_mult=2, eta_min=args.min_lr)
        self.vae_scheduler = CosineAnnealingWarmRestarts(self.vae_optimizer, **sk)
        self.teacher_scheduler = CosineAnnealingWarmRestarts(self.teacher_optimizer, **sk)
        # data after the models, like the reference (train_hybrid.py:273-275): the 90/10 split continues the seeded
        # CPU stream right after the weight init, so one process sees the reference's split (and first-epoch order)
        self._setup_data()
        # per-rank dropout / epsilon streams (rank 0 keeps the reference's stream position)
        if self.rank > 0:
            torch.manual_seed(args.seed + self.rank)
        torch.cuda.manual_seed_all(args.seed + self.rank)
        self.reducer = None
        if self.world > 1:
            self.reducer = GradBucketReducer(live_parameters(self.vae, self.teacher))
            self.vae_optimizer.grad_scale = self.teacher_optimizer.grad_scale = 1.0 / self.world
        self._after_reduce = None     # bench.py dp_check: called between the gradient all-reduce and the optimizer
        self.baseline = None          # float64 device scalar (reference: python float, train_hybrid.py:875-879)
        self.early_stopping = EarlyStopping(patience=args.early_stopping_patience)
        self.global_step = 0
        self.best_loss = float("inf")
        self.reward_scale = args.reward_scale
        self.semantic_weight = args.semantic_weight
        self.baseline_momentum = args.baseline_momentum
        self.checkpoint_dir = os.path.join(args.output_dir, "checkpoints")
        if getattr(args, "resume_from", None):
            self._load_checkpoint(args.resume_from)

    # ------------------------------------------------------------------ data (train_hybrid.py:529-585)
    def _setup_data(self):
        a = self.args
        if a.data_dir == "synthetic":
            self.dataset = SyntheticSprites(max(8 * a.batch_size * self.world, 64))
        else:
            self.dataset = PixelArtDataset(a.data_dir)
        train_idx, val_idx = split_indices(len(self.dataset))            # consumes the default generator (:555)
        gen = None
        if self.world > 1:                # every rank must draw the same epoch permutations: a shared private stream
            gen = torch.Generator()
            gen.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
        mk = dict(batch_size=a.batch_size, device=self.device, rank=self.rank, world=self.world)
        self.train_loader = SpriteLoader(self.dataset, train_idx, shuffle=True, generator=gen, **mk)
        self.val_loader = SpriteLoader(self.dataset, val_idx, shuffle=False, **mk)

    # ------------------------------------------------------------------ one step (train_hybrid.py:838-954)
    def _process_batch(self, images, batch_idx, return_tensor=False):
        a = self.args
        accum = a.gradient_accumulation_steps
        self.vae_optimizer.zero_grad(set_to_none=True)
        self.teacher_optimizer.zero_grad(set_to_none=True)
        images = images.detach()

        recon, mu, logvar = self.vae(images)
        with torch.no_grad():
            self.teacher(images)                                        # pass A: BN statistics / RNG position only
        recon_loss, kl_loss = vae_losses(recon, images, mu, logvar)     # fused MSE + KL (train_hybrid.py:859-862)
        teacher_eval = self.teacher(recon.detach())                     # pass B
        quality_scores = teacher_eval['quality_scores']
        semantic_score = teacher_eval['semantic_score']
        quality_reward = quality_scores.mean(dim=1, keepdim=True)
        total_reward = quality_reward + self.semantic_weight * semantic_score
        tr = total_reward.mean().detach().double()                      # the reference's EMA is a python float (f64)
        if self.baseline is None:
            self.baseline = tr
        else:
            self.baseline = self.baseline_momentum * self.baseline + (1 - self.baseline_momentum) * tr
        advantage = (total_reward - self.baseline.float()).detach() * self.reward_scale
        pg_loss = -(advantage * recon_loss).mean()
        vae_loss = (a.recon_weight * recon_loss + a.kl_weight * kl_loss + pg_loss) / accum
        quality_loss = -torch.mean(quality_scores)
        teacher_loss = a.quality_weight * quality_loss / accum

        boundary = (batch_idx + 1) % accum == 0
        if self.reducer is not None:
            self.reducer.enabled = boundary
        vae_loss.backward()
        teacher_loss.backward()

        if boundary:
            if self.reducer is not None:
                self.reducer.finish()
                if self._after_reduce is not None:
                    self._after_reduce(self)
            self.vae_optimizer.step()                    # clip (train_hybrid.py:913-915) happens inside the fused step
            self.teacher_optimizer.step()
            self.vae_scheduler.step()
            self.teacher_scheduler.step()

        packed = torch.stack([recon_loss, kl_loss, quality_loss, pg_loss, semantic_score.mean(), quality_reward.mean(),
                              self.baseline.float(), advantage.mean(), vae_loss, teacher_loss, vae_loss + teacher_loss,
                              quality_scores.mean()]).detach().float()
        self.global_step += 1
        self._last_recon = recon.detach()
        if return_tensor:
            return packed
        return dict(zip(METRIC_KEYS, packed.cpu().tolist()))

    # ------------------------------------------------------------------ checkpoints (train_hybrid.py:594-615, 791-836)
    def _save_checkpoint(self, is_best=False):
        if self.rank != 0:
            return None
        os.makedirs(self.checkpoint_dir, exist_ok=True)
        ckpt = {
            'global_step': self.global_step,
            'vae_state_dict': self.vae.state_dict(),
            'teacher_state_dict': self.teacher.state_dict(),
            'vae_optimizer': self.vae_optimizer.state_dict(),
            'teacher_optimizer': self.teacher_optimizer.state_dict(),
            'vae_scheduler': self.vae_scheduler.state_dict(),
            'teacher_scheduler': self.teacher_scheduler.state_dict(),
            'best_loss': self.best_loss,
            'args': vars(self.args),
        }
        path = os.path.join(self.checkpoint_dir, 'latest.pt')
        torch.save(ckpt, path)
        if is_best:
            torch.save(ckpt, os.path.join(self.checkpoint_dir, 'best.pt'))
        return path

    def _load_checkpoint(self, path):
        ckpt = torch.load(path, map_location=self.device, weights_only=True)
        self.vae.load_state_dict(ckpt['vae_state_dict'], strict=False)
        self.teacher.load_state_dict(ckpt['teacher_state_dict'], strict=False)
        self.vae_optimizer.load_state_dict(ckpt['vae_optimizer'])
        self.teacher_optimizer.load_state_dict(ckpt['teacher_optimizer'])
        self.vae_scheduler.load_state_dict(ckpt['vae_scheduler'])
        self.teacher_scheduler.load_state_dict(ckpt['teacher_scheduler'])
        self.global_step = ckpt['global_step']
        self.best_loss = ckpt['best_loss']

    # ------------------------------------------------------------------ loop (train_hybrid.py:956-1070, hot part only)
    def train(self):
        """Epoch loop. Deviation from the reference, on purpose: the reference never appends to `epoch_losses`
        (train_hybrid.py:987), so its epoch mean is NaN, `best.pt` is never written and early stopping never fires
        (SURVEY.md 0.8); here the epoch mean of `total_loss` drives both, as the code evidently intends."""
        a = self.args
        for epoch in range(a.num_epochs):
            t0, losses = time.time(), []
            for batch_idx, images in enumerate(self.train_loader.epoch()):
                m = self._process_batch(images, batch_idx)
                losses.append(m['total_loss'])
                if self.rank == 0 and self.global_step % a.log_every == 0:
                    print(f"step {self.global_step} " + " ".join(f"{k}={v:.4f}" for k, v in m.items()), flush=True)
            epoch_loss = float(np.mean(losses)) if losses else float("nan")
            if self.rank == 0:
                n_img = len(losses) * a.batch_size * self.world
                print(f"epoch {epoch} loss {epoch_loss:.4f} {n_img / max(time.time() - t0, 1e-9):.1f} img/s",
                      flush=True)
            self.early_stopping(epoch_loss)
            if self.early_stopping.early_stop:
                if self.rank == 0:
                    print("Early stopping triggered", flush=True)
                break
            if epoch_loss < self.best_loss:
                self.best_loss = epoch_loss
                self._save_checkpoint(is_best=True)
        self._save_checkpoint()


def main(argv=None):
    args = build_arg_parser().parse_args(argv)
    trainer = TrainingManager(args)
    trainer.train()
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
