"""bench.py contract on the CPU: the reference arm (`--impl reference`: the UNMODIFIED reference trainer from
/root/reference or the oracle/_ref bytecode, timed on the host cores; the oracle port only if neither exists) prints
ONE JSON line with the keys the driver reads; under torchrun only rank 0 prints; our arm refuses to run without a GPU
(no CPU fallback on the product path)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--steps", "1", "--warmup", "0", "--cpu-batch", "1", "--feat", "64", "--latent", "64", "--emb", "32"]


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    e["CUDA_VISIBLE_DEVICES"] = ""
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          env=e, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference"] + SMALL)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["unit"] == "images/s" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    from oracle import reference_loader
    assert cb["kind"] == ("reference" if reference_loader.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb and cb["timed_steps"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_from_the_staged_bytecode_alone():
    """What the GPU box has: no /root/reference, only oracle/_ref/*.pyc (oracle/make_ref.py)."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isfile(os.path.join(ref, "train_hybrid.rbc")):
        import pytest
        pytest.skip("oracle/_ref not built")
    r = _run(["--impl", "reference"] + SMALL, env={"LUNARIS_REFERENCE": ref})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["kind"] == "reference" and "bytecode" in d["cpu_baseline"]["sample"]


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(["--impl", "reference", "--gpus", "2"] + SMALL, env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_our_arm_fails_loudly_without_a_gpu():
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
