"""Smallest end-to-end exercise of every kernel for compute-sanitizer (one tool per call)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from lunaris_orion_b200.train_hybrid import TrainingManager, build_arg_parser
from lunaris_orion_b200 import lunar_generate as lg
import teacher_cases as tc
dev = torch.device("cuda:0")
args = build_arg_parser().parse_args(["--data_dir", "synthetic", "--batch_size", "2", "--gradient_accumulation_steps", "1",
                                      "--latent_dim", "64", "--embedding_dim", "32", "--feature_dim", "64"])
tm = TrainingManager(args, device=dev)
m = tm._process_batch(tc.images(2, 3).to(dev), 0)
print("step ok", {k: round(v, 4) for k, v in list(m.items())[:3]})
att = lg.SelfAttention2d(64).to(dev)
with torch.no_grad():
    y = att(torch.randn(1, 64, 16, 16, device=dev))
    s = tm.vae.sample(2)
torch.cuda.synchronize()
print("done", float(y.abs().mean()), float(s.abs().mean()))
