"""Host-side input pipeline (lunaris_orion_b200/data.py) on CPU: dataset / labels / 90-10 split / epoch
order against the UNMODIFIED reference dataset class and torch DataLoader (train_hybrid.py:100-201, 529-585), multi-
file indexing, rank sharding, the prefetching iterator and the reference normalisation."""
import os

import numpy as np
import pytest
import torch

from lunaris_orion_b200 import data as D
from oracle import ref_harness, reference_loader


def _write(dirname, sizes, seed=5):
    rng = np.random.default_rng(seed)
    arrs = []
    for k, n in enumerate(sizes):
        a = rng.integers(0, 256, (n, 128, 128, 3), dtype=np.uint8)
        arrs.append(a)
        np.save(os.path.join(dirname, f"sprites_{k:03d}.npy"), a)
        with open(os.path.join(dirname, f"labels_{k:03d}.csv"), "w") as f:
            f.write(ref_harness.LABEL_HEADER + "\n")
            for i in range(n):
                f.write(f'f{k}_{i}.png,cat{i % 3},"a prompt, with a comma {i}",{1000 + i},8,7.5,3.0,{20 + i}\n')
    return np.concatenate(arrs)


@pytest.mark.skipif(not reference_loader.available(), reason="reference not present")
def test_dataset_split_and_epoch_order_match_the_reference(tmp_path):
    _write(str(tmp_path), (13, 17))
    th = ref_harness.import_reference_trainer()
    ref_ds, mine = th.PixelArtDataset(str(tmp_path)), D.PixelArtDataset(str(tmp_path))
    assert len(ref_ds) == len(mine) == 30
    assert list(mine.cumulative_sizes) == list(ref_ds.cumulative_sizes)
    for idx in (0, 12, 13, 29):                              # both sides of the file boundary
        a, b = ref_ds[idx], mine[idx]
        assert torch.equal(a["image"], b["image"])
        assert mine._get_sprite_index(idx) == tuple(int(v) for v in ref_ds._get_sprite_index(idx))
        for c in D.LABEL_COLUMNS:                            # same values, same python types as pandas' inference
            assert a["metadata"][c] == b["metadata"][c] and type(b["metadata"][c])(a["metadata"][c]) == b["metadata"][c]
    # seeded 90/10 split (train_hybrid.py:551-555) and the first epoch's batch order of DataLoader(shuffle=True)
    torch.manual_seed(42)
    tr, va = torch.utils.data.random_split(ref_ds, [27, 3])
    order = []
    orig = type(ref_ds).__getitem__
    type(ref_ds).__getitem__ = lambda self, i: (order.append(i), orig(self, i))[1]
    try:
        for _ in th.DataLoader(tr, shuffle=True, batch_size=4, num_workers=0, drop_last=True):
            pass
    finally:
        type(ref_ds).__getitem__ = orig
    torch.manual_seed(42)
    t2, v2 = D.split_indices(30)
    assert t2 == list(tr.indices) and v2 == list(va.indices)
    ld = D.SpriteLoader(mine, t2, 4, "cpu")
    assert len(ld) == 6
    assert ld.batch_indices().reshape(-1).tolist() == order


def test_loader_batches_are_sharded_prefetched_and_normalised(tmp_path):
    allx = _write(str(tmp_path), (12, 9, 11), seed=0)
    ds = D.PixelArtDataset(str(tmp_path))
    train = list(range(len(ds)))
    seen = []
    for rank in range(2):
        g = torch.Generator().manual_seed(7)                # same generator state on every rank -> same permutation
        ld = D.SpriteLoader(ds, train, 4, "cpu", rank=rank, world=2, generator=g, depth=2)
        assert len(ld) == 4                                 # 32 sprites / (4 per rank * 2 ranks)
        n = 0
        for x, idx in ld.epoch(with_indices=True):
            n += 1
            assert x.shape == (4, 3, 128, 128) and x.dtype == torch.float32
            want = torch.from_numpy(allx[idx].astype(np.float32) / 127.5 - 1.0).permute(0, 3, 1, 2)
            assert torch.equal(x, want)                     # x / 127.5 - 1, HWC -> CHW (train_hybrid.py:181-182)
            seen += idx.tolist()
        assert n == 4
    assert len(seen) == len(set(seen)) == 32                # ranks see disjoint sprites, nothing twice in an epoch


def test_gather_crosses_file_boundaries_in_request_order(tmp_path):
    allx = _write(str(tmp_path), (5, 7, 3), seed=2)
    ds = D.PixelArtDataset(str(tmp_path))
    idx = np.array([14, 0, 5, 4, 12, 11, 6, 13])
    out = np.empty((8, 128, 128, 3), dtype=np.uint8)
    ds.gather_u8(idx, out)
    assert np.array_equal(out, allx[idx])
    assert ds.metadata(12)["filename"] == "f2_0.png" and ds.metadata(4)["num_steps"] == 24


def test_label_mismatch_and_missing_files_are_rejected(tmp_path):
    with pytest.raises(ValueError):
        D.PixelArtDataset(str(tmp_path))
    _write(str(tmp_path), (4,))
    with open(tmp_path / "labels_000.csv", "a") as f:
        f.write("extra.png,c,p,1,8,7.5,3.0,20\n")
    with pytest.raises(AssertionError):
        D.PixelArtDataset(str(tmp_path))


def test_forever_iterator_chains_epochs_and_synthetic_data():
    ds = D.SyntheticSprites(24)
    ld = D.SpriteLoader(ds, list(range(24)), 8, "cpu")
    it = ld.forever()
    xs = [next(it) for _ in range(7)]                       # more than two epochs of three batches
    assert all(x.shape == (8, 3, 128, 128) for x in xs)
    assert float(min(x.min() for x in xs)) >= -1.0 and float(max(x.max() for x in xs)) <= 1.0
    it.close()
