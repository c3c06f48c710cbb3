"""Which torch ops launch a given ATen kernel inside one training step (C2 shapes)? Prints the ops, input shapes and
total device time behind every kernel whose name contains `elementwise_kernel<128, {2,4}` - the strided-copy / strided-fill
kernels of the host glue (found: the fc weight re-layouts after an optimizer step, the padding-row fills)."""
import os, sys, collections
import torch
sys.path.insert(0, os.getcwd())
import bench
from lunaris_orion_b200.train_hybrid import TrainingManager
B, L, E, F = 16, 256, 128, 256
dev = torch.device("cuda:0")
tm = TrainingManager(bench._args_ns(B, L, E, F), device=dev)
x = torch.rand(B, 3, 128, 128, device=dev) * 2 - 1
for i in range(2): tm._process_batch(x, i, return_tensor=True)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
    tm._process_batch(x, 2, return_tensor=True); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0, set()])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CPU and getattr(e, "kernels", None):
        for k in e.kernels:
            if "elementwise_kernel<128, 2" in k.name or "elementwise_kernel<128, 4" in k.name:
                st = [s for s in (e.stack or []) if "lunaris_orion_b200" in s][:2]
                key = (e.name, str(e.input_shapes)[:80], tuple(st))
                agg[key][0] += 1; agg[key][1] += k.duration
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{v[0]:4d} x  {v[1]:8.1f} us total  {k[0]}  {k[1]}\n        {k[2]}")
