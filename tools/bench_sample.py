"""Config C5 (BASELINE.json configs[4]): decoder-only sampling, batch 256, latent 512 -> images/s on one B200."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200.lunar_generate import LunarisCoreVAE
dev = torch.device("cuda:0")
torch.manual_seed(42)
vae = LunarisCoreVAE(512).to(dev).eval()
B = 256
for _ in range(3): vae.sample(B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): vae.sample(B)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(json.dumps({"workload": "C5 decoder-only sampling B=256 latent 512", "ms_per_batch": round(ms, 3),
                  "images_per_s": round(B / ms * 1e3, 1), "tflops": round(B * 1.136e9 / ms / 1e9, 1),
                  "hbm_gbs_at_4.3MB_per_img": round(B * 4.3e6 / ms / 1e6, 1)}))
