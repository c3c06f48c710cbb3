"""Is the step CPU-(launch-)bound? Host enqueue time vs device time of K steps, at C2 and C3 shapes."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200.train_hybrid import TrainingManager, build_arg_parser
dev = torch.device("cuda:0")
for name, B, L, E, F in (("C2", 16, 256, 128, 256), ("C3", 64, 512, 256, 512)):
    args = build_arg_parser().parse_args(["--data_dir", "synthetic", "--batch_size", str(B),
        "--gradient_accumulation_steps", "1", "--latent_dim", str(L), "--embedding_dim", str(E), "--feature_dim", str(F)])
    tm = TrainingManager(args, device=dev)
    x = torch.rand(B, 3, 128, 128, device=dev) * 2 - 1
    for i in range(3): tm._process_batch(x, i, return_tensor=True)
    torch.cuda.synchronize()
    K = 6
    t0 = time.time()
    for i in range(K): tm._process_batch(x, i, return_tensor=True)
    t1 = time.time()
    torch.cuda.synchronize()
    t2 = time.time()
    print(f"{name}: host enqueue {1e3 * (t1 - t0) / K:.1f} ms/step, wall to completion {1e3 * (t2 - t0) / K:.1f} ms/step")
    del tm
    torch.cuda.empty_cache()
