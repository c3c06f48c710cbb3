"""Run every conv parity case in its own process (a trapping kernel then cannot poison the others)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if len(sys.argv) > 1 and sys.argv[1] == "--one":
    import torch
    from conv_cases import CASES
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    fn, kw = CASES[sys.argv[2]]
    err, tol = fn(torch.device("cuda:0"), **kw)
    print(f"RESULT {sys.argv[2]} err={err:.5g} tol={tol:.5g} {'PASS' if err <= tol else 'FAIL'}")
    sys.exit(0)

from conv_cases import CASES
names = sys.argv[1:] or sorted(CASES)
for n in names:
    try:
        p = subprocess.run([sys.executable, __file__, "--one", n], capture_output=True, text=True, timeout=180)
        lines = [l for l in (p.stdout + p.stderr).splitlines() if l.strip()]
        res = [l for l in lines if l.startswith("RESULT")]
        print(res[0] if res else f"RESULT {n} CRASH rc={p.returncode} :: " + " | ".join(lines[-4:]))
    except subprocess.TimeoutExpired:
        print(f"RESULT {n} TIMEOUT")
    sys.stdout.flush()
