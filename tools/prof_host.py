"""cProfile of the host side of one step at the C2 shapes (where the step is launch-bound)."""
import cProfile, os, pstats, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200.train_hybrid import TrainingManager, build_arg_parser
dev = torch.device("cuda:0")
args = build_arg_parser().parse_args(["--data_dir", "synthetic", "--batch_size", "16", "--gradient_accumulation_steps", "1",
                                      "--latent_dim", "256", "--embedding_dim", "128", "--feature_dim", "256"])
tm = TrainingManager(args, device=dev)
x = torch.rand(16, 3, 128, 128, device=dev) * 2 - 1
for i in range(3): tm._process_batch(x, i, return_tensor=True)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(5): tm._process_batch(x, i, return_tensor=True)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
