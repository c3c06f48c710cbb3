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

import numpy as np
import torch
import torch.distributed as dist
from torch.optim.lr_scheduler import CosineAnnealingWarmRestarts

from .lunar_evaluator import LunarMoETeacher
from .lunar_generate import LunarisCoreVAE, sprites_to_tensor, vae_losses
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
    """Bucketed gradient all-reduce (average) launched from grad-ready hooks on a side stream.

    Parameters are packed into ~bucket_mb buckets in the given order; a bucket is flattened and all-reduced as soon as
    every member has its gradient, so the VAE's buckets travel over NVLink while the Teacher backward still runs.
    Parameters that never get a gradient (the reference's None-set) are simply never waited for: finish() reduces
    whatever is ready, identically on every rank."""

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
        self.ready = [0] * len(self.buckets)
        self.inflight = {}
        self.enabled = True
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
        members = [p for p in self.buckets[i] if p.grad is not None]
        if not members or i in self.inflight:
            return
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                flat = torch.cat([p.grad.reshape(-1).float() for p in members])
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            flat = torch.cat([p.grad.reshape(-1).float() for p in members])
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.inflight[i] = (members, flat, work)

    def finish(self):
        """Reduce any bucket that did not fill (None-set members), wait, and write the averaged grads back."""
        if self.world == 1:
            return
        for i in range(len(self.buckets)):
            self._launch(i)
        for i, (members, flat, work) in sorted(self.inflight.items()):
            work.wait()
            if self.cuda:
                torch.cuda.current_stream().wait_stream(self.stream)
            off = 0
            flat.mul_(1.0 / self.world)
            for p in members:
                n = p.numel()
                p.grad.copy_(flat[off:off + n].view_as(p.grad))
                off += n
        self.inflight.clear()
        self.ready = [0] * len(self.buckets)


# ====================================================================================================== data
class SpriteData:
    """sprites*.npy (uint8 [N,128,128,3]) reader compatible with the reference's PixelArtDataset normalisation
    (train_hybrid.py:181-182: x/127.5 - 1, HWC -> CHW), sharded by rank; 'synthetic' generates SURVEY.md §8d data."""

    def __init__(self, data_dir, batch_size, rank=0, world=1, seed=1234):
        if data_dir == "synthetic":
            rng = np.random.default_rng(seed + rank)
            self.arr = rng.integers(0, 256, (max(4 * batch_size, 64), 128, 128, 3), dtype=np.uint8)
        else:
            files = sorted(f for f in os.listdir(data_dir) if f.startswith("sprites") and f.endswith(".npy"))
            if not files:
                raise FileNotFoundError(f"no sprites*.npy under {data_dir}")
            arrs = [np.load(os.path.join(data_dir, f), mmap_mode="r") for f in files]
            self.arr = arrs[0] if len(arrs) == 1 else np.concatenate(arrs)
        self.bs, self.rank, self.world = batch_size, rank, world

    def __len__(self):
        return len(self.arr) // (self.bs * self.world)

    def batches(self, epoch, device):
        n = len(self)
        order = np.random.default_rng(epoch).permutation(len(self.arr))[: n * self.bs * self.world]
        order = order.reshape(n, self.world, self.bs)[:, self.rank]
        for idx in order:
            u8 = torch.from_numpy(np.ascontiguousarray(self.arr[np.sort(idx)]))
            if device.type == "cuda":
                yield sprites_to_tensor(u8.pin_memory().to(device, non_blocking=True))
            else:
                yield u8.permute(0, 3, 1, 2).float().div_(127.5).sub_(1.0)


# ====================================================================================================== trainer
class TrainingManager:
    """Drop-in for the reference's TrainingManager on the hot path: models, optimizers, schedulers, one step,
    checkpoint save / load (train_hybrid.py:382-404, 502-527, 594-615, 791-836, 838-954)."""

    def __init__(self, args, device=None):
        self.args = args
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("lunaris_orion_b200 needs a CUDA device (B200, sm_100a): there is no CPU path")
            device = torch.device("cuda", self.local_rank)
            torch.cuda.set_device(device)
        self.device = device
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=device)
        # identical initial weights on every rank: seed before construction, VAE first (train_hybrid.py:1138, 394, 400)
        torch.manual_seed(args.seed)
        np.random.seed(args.seed)
        torch.cuda.manual_seed_all(args.seed)
        self.vae = LunarisCoreVAE(latent_dim=args.latent_dim).to(device).train()
        self.teacher = LunarMoETeacher(num_experts=args.num_experts, feature_dim=args.feature_dim,
                                       embedding_dim=args.embedding_dim).to(device).train()
        # per-rank dropout / epsilon streams
        torch.manual_seed(args.seed + self.rank)
        torch.cuda.manual_seed_all(args.seed + self.rank)
        # clip_grad_norm_ + AdamW fused into two multi-tensor launches per model (state layout == torch.optim.AdamW)
        mk = dict(weight_decay=args.weight_decay, betas=(0.9, 0.999), max_grad_norm=args.max_grad_norm)
        self.vae_optimizer = ClipAdamW(self.vae.parameters(), lr=args.vae_lr, **mk)
        self.teacher_optimizer = ClipAdamW(self.teacher.parameters(), lr=args.teacher_lr, **mk)
        sk = dict(T_0=args.scheduler_t0, T_mult=2, eta_min=args.min_lr)
        self.vae_scheduler = CosineAnnealingWarmRestarts(self.vae_optimizer, **sk)
        self.teacher_scheduler = CosineAnnealingWarmRestarts(self.teacher_optimizer, **sk)
        self.reducer = GradBucketReducer(list(self.vae.parameters()) + list(self.teacher.parameters())) \
            if self.world > 1 else None
        self.baseline = None          # device scalar (reference: python float, train_hybrid.py:875-879)
        self.global_step = 0
        self.best_loss = float("inf")
        self.reward_scale = args.reward_scale
        self.semantic_weight = args.semantic_weight
        self.baseline_momentum = args.baseline_momentum
        self.checkpoint_dir = os.path.join(args.output_dir, "checkpoints")
        if getattr(args, "resume_from", None):
            self._load_checkpoint(args.resume_from)

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
        tr = total_reward.mean().detach()
        if self.baseline is None:
            self.baseline = tr
        else:
            self.baseline = self.baseline_momentum * self.baseline + (1 - self.baseline_momentum) * tr
        advantage = (total_reward - self.baseline).detach() * self.reward_scale
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
            self.vae_optimizer.step()                    # clip (train_hybrid.py:913-915) happens inside the fused step
            self.teacher_optimizer.step()
            self.vae_scheduler.step()
            self.teacher_scheduler.step()

        packed = torch.stack([recon_loss, kl_loss, quality_loss, pg_loss, semantic_score.mean(), quality_reward.mean(),
                              self.baseline, advantage.mean(), vae_loss, teacher_loss, vae_loss + teacher_loss,
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
        a = self.args
        data = SpriteData(a.data_dir, a.batch_size, self.rank, self.world)
        for epoch in range(a.num_epochs):
            t0, losses = time.time(), []
            for batch_idx, images in enumerate(data.batches(epoch, self.device)):
                m = self._process_batch(images, batch_idx)
                losses.append(m['total_loss'])
                if self.rank == 0 and self.global_step % a.log_every == 0:
                    print(f"step {self.global_step} " + " ".join(f"{k}={v:.4f}" for k, v in m.items()), flush=True)
            epoch_loss = float(np.mean(losses)) if losses else float("nan")
            if self.rank == 0:
                n_img = len(losses) * a.batch_size * self.world
                print(f"epoch {epoch} loss {epoch_loss:.4f} {n_img / max(time.time() - t0, 1e-9):.1f} img/s",
                      flush=True)
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
