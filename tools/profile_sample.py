import os, sys, collections, re
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200.lunar_generate import LunarisCoreVAE
dev = torch.device("cuda:0")
torch.manual_seed(42)
vae = LunarisCoreVAE(512).to(dev).eval()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for _ in range(3): vae.sample(N)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    vae.sample(N); torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
for e in ev:
    print(f"{e.device_time:8.1f} us  {re.sub(r'\(.*','',e.name)[:60]}")
print("span", (ev[-1].time_range.end - ev[0].time_range.start), "us; sum", sum(e.device_time for e in ev))
