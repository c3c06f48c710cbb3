"""Tiny driver for ncu: `steps` hybrid training steps at C3 (or the shapes given) and nothing else."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lunaris_orion_b200.train_hybrid import TrainingManager

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B, L, E, F = (int(v) for v in sys.argv[2:6]) if len(sys.argv) > 5 else (64, 512, 256, 512)
dev = torch.device("cuda:0")
tm = TrainingManager(bench._args_ns(B, L, E, F), device=dev)
x = torch.rand(B, 3, 128, 128, device=dev) * 2 - 1
for i in range(steps):
    tm._process_batch(x, i, return_tensor=True)
torch.cuda.synchronize()
print("done")
