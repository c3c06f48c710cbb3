"""Host-side input pipeline: sprites_*.npy reader, rank sharding and the reference normalisation."""
import numpy as np
import torch

from lunaris_orion_b200.train_hybrid import SpriteData


def test_sprite_files_are_read_sharded_and_normalised(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (12, 128, 128, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (8, 128, 128, 3), dtype=np.uint8)
    np.save(tmp_path / "sprites_000.npy", a)
    np.save(tmp_path / "sprites_001.npy", b)
    allx = np.concatenate([a, b])
    seen = []
    for rank in range(2):
        d = SpriteData(str(tmp_path), 4, rank=rank, world=2)
        assert len(d) == 2                                    # 20 sprites / (4 per rank * 2 ranks)
        for x in d.batches(0, torch.device("cpu")):
            assert x.shape == (4, 3, 128, 128) and x.dtype == torch.float32
            u8 = torch.round((x + 1.0) * 127.5).to(torch.uint8).permute(0, 2, 3, 1).numpy()
            for img in u8:                                    # x/127.5 - 1 of an actual sprite (train_hybrid.py:181)
                idx = np.where((allx == img).all(axis=(1, 2, 3)))[0]
                assert len(idx) == 1
                seen.append(int(idx[0]))
    assert len(seen) == len(set(seen)) == 16                  # ranks see disjoint sprites


def test_synthetic_data_follows_the_survey_recipe():
    d = SpriteData("synthetic", 8)
    x = next(d.batches(0, torch.device("cpu")))
    assert x.shape == (8, 3, 128, 128) and float(x.min()) >= -1.0 and float(x.max()) <= 1.0
