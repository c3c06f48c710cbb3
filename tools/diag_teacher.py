import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import teacher_cases as tc
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("eval", "both"):
    print("EVAL", json.dumps(tc.eval_report(dev), indent=1, default=str), flush=True)
if which in ("train", "both"):
    print("TRAIN", json.dumps(tc.train_report(dev), indent=1, default=str), flush=True)
if which in ("trunk",):
    print("TRUNK", json.dumps(tc.trunk_report(dev), indent=1, default=str), flush=True)
