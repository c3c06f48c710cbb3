"""Input pipeline of the hybrid trainer: the reference's on-disk format and dataset semantics
(MeryylleA/Lunaris-Orion train_hybrid.py:100-201 PixelArtDataset, :529-585 _setup_data) feeding the B200 step.

Kept from the reference: `sprites*.npy` (uint8 [N,128,128,3], memory-mapped, any number of files, global index ->
(file, local index) through cumulative sizes :131,153-157), `labels*.csv` (8 columns :106, one row per sprite,
length checked :147-149), `x / 127.5 - 1` + HWC->CHW (:181-182), the 90 / 10 `random_split` drawn from torch's default
generator (:551-555) and RandomSampler's epoch permutation, so a single process sees the reference's split and batch
order for the same seed.

Different by design (the reference uses spawn worker processes and a pandas `iloc` per item, :185,561-570): batches
are gathered as uint8 NHWC - the layout the kernels want - straight from the memory maps into a ring of pinned host
buffers by one prefetch thread, copied on a copy stream, and normalised on the GPU by `lun_sprites_u8_to_f32`; no
per-item Python, no fp32 on the host, no `np.concatenate` of the files. Ranks take disjoint slices of every global
batch."""
import csv
import os
import queue
import threading

import numpy as np
import torch

LABEL_COLUMNS = ("filename", "category", "prompt", "seed", "pixel_size", "guidance_scale", "pag_scale", "num_steps")


def _infer(col):
    """Column typing like pandas.read_csv: int if every field parses as int, else float, else str."""
    for cast in (int, float):
        try:
            return [cast(v) for v in col]
        except ValueError:
            continue
    return col


class PixelArtDataset:
    """Drop-in for train_hybrid.py:100-201 (without the unused teacher_model hook): same files, same length check,
    same `__getitem__` dictionary; plus batched uint8 gathers for the GPU loader."""

    def __init__(self, data_dir):
        names = sorted(os.listdir(data_dir)) if os.path.isdir(data_dir) else []
        self.sprites_files = [os.path.join(data_dir, f) for f in names if f.startswith("sprites") and f.endswith(".npy")]
        self.labels_files = [os.path.join(data_dir, f) for f in names if f.startswith("labels") and f.endswith(".csv")]
        if not self.sprites_files or not self.labels_files:
            raise ValueError(f"No sprites or labels files found in {data_dir}")
        self.sprites = []
        for f in self.sprites_files:
            arr = np.load(f, mmap_mode="r")
            if arr.shape[1:] != (128, 128, 3):
                raise ValueError(f"Expected 128x128x3 images in {f}, got {arr.shape[1:]}")
            if arr.dtype != np.uint8:
                raise ValueError(f"Expected uint8 sprites in {f}, got {arr.dtype}")
            self.sprites.append(arr)
        self.cumulative_sizes = np.cumsum([0] + [len(a) for a in self.sprites])
        cols = {c: [] for c in LABEL_COLUMNS}
        for f in self.labels_files:
            with open(f, newline="") as fh:
                rd = csv.DictReader(fh)
                missing = [c for c in LABEL_COLUMNS if c not in (rd.fieldnames or [])]
                if missing:
                    raise ValueError(f"{f}: missing label columns {missing}")
                for row in rd:
                    for c in LABEL_COLUMNS:
                        cols[c].append(row[c])
        self.labels = {c: _infer(v) for c, v in cols.items()}
        total = int(self.cumulative_sizes[-1])
        assert len(self.labels["filename"]) == total, \
            f"Mismatch between total sprites ({total}) and labels ({len(self.labels['filename'])})"

    def __len__(self):
        return int(self.cumulative_sizes[-1])

    def _get_sprite_index(self, idx):
        file_idx = int(np.searchsorted(self.cumulative_sizes, idx, side="right") - 1)
        return file_idx, int(idx - self.cumulative_sizes[file_idx])

    def metadata(self, idx):
        return {c: self.labels[c][idx] for c in LABEL_COLUMNS}

    def __getitem__(self, idx):
        f, i = self._get_sprite_index(idx)
        image = torch.from_numpy(self.sprites[f][i].astype(np.float32) / 127.5 - 1.0).permute(2, 0, 1)
        return {"image": image, "metadata": self.metadata(idx)}

    def gather_u8(self, indices, out):
        """out[j] = sprite indices[j] (uint8 NHWC). Reads are grouped per file and issued in ascending local order
        (sequential pages of the memory map); `out` is typically a pinned buffer."""
        indices = np.asarray(indices, dtype=np.int64)
        files = np.searchsorted(self.cumulative_sizes, indices, side="right") - 1
        local = indices - self.cumulative_sizes[files]
        for f in np.unique(files):
            sel = np.nonzero(files == f)[0]
            order = sel[np.argsort(local[sel], kind="stable")]
            out[order] = self.sprites[f][local[order]]
        return out


class SyntheticSprites:
    """`--data_dir synthetic`: SURVEY.md 8(d) uniform-random uint8 sprites held in memory (no files, no labels)."""

    def __init__(self, n, seed=1234):
        self.arr = np.random.default_rng(seed).integers(0, 256, (n, 128, 128, 3), dtype=np.uint8)

    def __len__(self):
        return len(self.arr)

    def metadata(self, idx):
        return {c: None for c in LABEL_COLUMNS}

    def gather_u8(self, indices, out):
        np.take(self.arr, np.asarray(indices, dtype=np.int64), axis=0, out=out)
        return out


def split_indices(n, generator=None):
    """The reference's `random_split(dataset, [int(0.9 n), n - int(0.9 n)])` (train_hybrid.py:551-555): one
    `randperm(n)` from the given (default: torch's global) generator, first 90 % train, rest validation."""
    train_size = int(0.9 * n)
    perm = torch.randperm(n, generator=generator).tolist()
    return perm[:train_size], perm[train_size:]


def epoch_permutation(n, generator=None):
    """The order torch's DataLoader(shuffle=True) visits a dataset in for one epoch: creating the loader iterator draws
    one 64-bit base seed from `generator` (default: the global one), then RandomSampler draws a second one that seeds
    a fresh generator for `randperm(n)`."""
    torch.empty((), dtype=torch.int64).random_(generator=generator)              # DataLoader's per-iterator base seed
    seed = int(torch.empty((), dtype=torch.int64).random_(generator=generator).item())
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g).tolist()


class SpriteLoader:
    """Batches of `batch_size` sprites per rank as fp32 NCHW tensors in [-1,1] on `device` (drop_last, like the
    reference's loaders). One prefetch thread fills a ring of pinned uint8 buffers; the host->device copy of batch k+1
    is issued on a copy stream before batch k is handed out."""

    def __init__(self, dataset, indices, batch_size, device, rank=0, world=1, shuffle=True, generator=None, depth=3):
        self.ds, self.indices = dataset, list(indices)
        self.bs, self.rank, self.world = batch_size, rank, world
        self.device = torch.device(device)
        self.shuffle, self.generator, self.depth = shuffle, generator, max(2, depth)
        self.h2d_bytes_per_batch = batch_size * 128 * 128 * 3
        self._host = self._dev = None

    def __len__(self):
        return len(self.indices) // (self.bs * self.world)

    def batch_indices(self):
        """Global dataset indices of every batch of one epoch for this rank: [len(self), batch_size]."""
        n = len(self)
        order = epoch_permutation(len(self.indices), self.generator) if self.shuffle else range(len(self.indices))
        idx = np.asarray([self.indices[i] for i in order][: n * self.bs * self.world], dtype=np.int64)
        return idx.reshape(n, self.world, self.bs)[:, self.rank]

    def _buffers(self):
        if self._host is None:
            shape = (self.bs, 128, 128, 3)
            cuda = self.device.type == "cuda"
            self._host = [torch.empty(shape, dtype=torch.uint8, pin_memory=cuda) for _ in range(self.depth)]
            if cuda:
                self._dev = [torch.empty(shape, dtype=torch.uint8, device=self.device) for _ in range(self.depth)]
                self._copy_stream = torch.cuda.Stream(self.device)
        return self._host

    def epoch(self, with_indices=False):
        """Iterator over one epoch."""
        return self._iterate([self.batch_indices()], with_indices)

    def forever(self, with_indices=False):
        """Iterator chaining epochs without end (benchmarks)."""
        if self.generator is None and self.shuffle:          # epochs are drawn on the prefetch thread: private stream
            self.generator = torch.Generator()
            self.generator.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))

        def epochs():
            while True:
                yield self.batch_indices()
        return self._iterate(epochs(), with_indices)

    def _iterate(self, epochs, with_indices):
        host = self._buffers()
        ready, free = queue.Queue(), queue.Queue()
        for s in range(self.depth):
            free.put((s, None))
        stop = threading.Event()

        def produce():
            try:
                for batches in epochs:
                    for idx in batches:
                        slot, ev = free.get()
                        if stop.is_set():
                            return
                        if ev is not None:
                            ev.synchronize()                  # the previous copy out of this pinned buffer finished
                        self.ds.gather_u8(idx, host[slot].numpy())
                        ready.put((slot, idx))
                ready.put(None)
            except BaseException as e:                        # surface loader errors in the training thread
                ready.put(e)
        th = threading.Thread(target=produce, daemon=True, name="lunaris-sprite-prefetch")
        th.start()

        def take():
            item = ready.get()
            if isinstance(item, BaseException):
                raise item
            return item

        def issue(item):
            """Start the host->device copy of a ready batch."""
            slot, idx = item
            if self.device.type != "cuda":
                return slot, idx, None
            cs = self._copy_stream
            with torch.cuda.stream(cs):
                self._dev[slot].copy_(host[slot], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
            return slot, idx, ev

        def finish(pending):
            slot, idx, ev = pending
            if self.device.type == "cuda":
                from .lunar_generate import sprites_to_tensor
                main = torch.cuda.current_stream(self.device)
                main.wait_event(ev)
                x = sprites_to_tensor(self._dev[slot])
                done = torch.cuda.Event()
                done.record(main)
                self._copy_stream.wait_event(done)            # the device slot is rewritten only after this read
                free.put((slot, ev))
            else:
                x = host[slot].permute(0, 3, 1, 2).float().div_(127.5).sub_(1.0)
                free.put((slot, None))
            return (x, idx) if with_indices else x

        try:
            item = take()
            pending = issue(item) if item is not None else None
            while pending is not None:
                item = take()
                nxt = issue(item) if item is not None else None
                yield finish(pending)
                pending = nxt
        finally:
            stop.set()
            free.put((0, None))
