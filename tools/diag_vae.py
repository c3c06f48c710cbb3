import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import vae_cases as vc
print("VAE", json.dumps(vc.vae_report(torch.device("cuda:0")), indent=1, default=str), flush=True)
