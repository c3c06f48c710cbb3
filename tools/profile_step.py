"""In-situ kernel time breakdown of one C3 training step (CUPTI via torch.profiler; no ncu serialisation)."""
import os, sys, collections, re
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lunaris_orion_b200.train_hybrid import TrainingManager

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L, E, F = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (512, 256, 512)    # latent, emb, feat
dev = torch.device("cuda:0")
tm = TrainingManager(bench._args_ns(B, L, E, F), device=dev)
x = torch.rand(B, 3, 128, 128, device=dev) * 2 - 1
for i in range(2):
    tm._process_batch(x, i, return_tensor=True)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tm._process_batch(x, 2, return_tensor=True)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\(.*", "", e.name)[:70]
        agg[name][0] += 1
        agg[name][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"total kernel time {tot/1e3:.1f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("PROF_TOP", "22"))]:
    print(f"{v[1]/tot*100:6.2f}%  {v[1]/1e3:8.2f} ms  n={v[0]:4d}  avg={v[1]/v[0]:8.1f} us  {k}")
# conv_fprop by duration bucket
b = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and ("conv_fprop" in e.name or "conv_wgrad" in e.name):
        key = ("W" if "wgrad" in e.name else "F", int(round(e.device_time / 100.0)) * 100)
        b[key][0] += 1
        b[key][1] += e.device_time
for k, v in sorted(b.items(), key=lambda kv: -kv[1][1])[:12]:
    print(k, v[0], f"{v[1]/1e3:.2f} ms")

# ---- idle gaps on the GPU timeline
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0],
            key=lambda e: e.time_range.start)
gaps = []
for a, b_ in zip(ev[:-1], ev[1:]):
    gap = b_.time_range.start - a.time_range.end
    if gap > 30:
        gaps.append((gap, a.name[:50], b_.name[:50]))
span = ev[-1].time_range.end - ev[0].time_range.start
print(f"timeline span {span/1e3:.1f} ms, sum of gaps > 30us: {sum(g[0] for g in gaps)/1e3:.1f} ms in {len(gaps)} gaps")
for g in sorted(gaps, reverse=True)[:14]:
    print(f"  {g[0]:8.0f} us  after {g[1]}  before {g[2]}")
