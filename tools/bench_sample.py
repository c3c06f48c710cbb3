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

# ---- generate-and-rank: decode + Teacher eval scoring, C3 Teacher (feat 512 / emb 256), batch 64
from lunaris_orion_b200.lunar_evaluator import LunarMoETeacher
from lunaris_orion_b200 import scoring
teacher = LunarMoETeacher(feature_dim=512, embedding_dim=256).to(dev).eval()
Bs = 64


def round_trip():
    z = torch.randn(Bs, 512, device=dev)
    return scoring.assess_quality(teacher, scoring.decode(vae, z))


for _ in range(2): round_trip()
torch.cuda.synchronize()
e0.record()
for _ in range(5): round_trip()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps({"workload": "generate-and-rank: decode + Teacher eval quality score, B=64, latent 512 / feat 512",
                  "ms_per_batch": round(ms, 2), "images_per_s": round(Bs / ms * 1e3, 1),
                  "teacher_fwd_tflops": round(Bs * 2058.4e9 / ms / 1e9, 1)}))
