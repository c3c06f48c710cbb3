"""Locate and import the UNMODIFIED reference (MeryylleA/Lunaris-Orion) for oracle pinning, golden generation and the
reference arm of bench.py. Resolution order: $LUNARIS_REFERENCE, /root/reference (build container: sources), then
oracle/_ref (byte-compiled by oracle/make_ref.py; the only form that exists on the GPU box).
TEST / BENCH INFRASTRUCTURE ONLY."""
import importlib
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get("LUNARIS_REFERENCE"), "/root/reference", os.path.join(_HERE, "_ref")]


def _module_file(d, name):
    for ext in (".py", ".pyc"):
        p = os.path.join(d, name + ext)
        if os.path.isfile(p):
            return p
    return None


def _resolve():
    for d in _CANDIDATES:
        if d and all(_module_file(d, m) for m in ("lunar_generate", "lunar_evaluator", "train_hybrid")):
            return d
    return _CANDIDATES[1]


REF = _resolve()


def available():
    return _module_file(REF, "lunar_evaluator") is not None


def kind():
    """'source' (build container) or 'bytecode' (oracle/_ref)."""
    f = _module_file(REF, "lunar_evaluator")
    return None if f is None else ("source" if f.endswith(".py") else "bytecode")


def load():
    """Returns (lunar_generate, lunar_evaluator) reference modules under private names (no sys.path pollution
    for the same-named drop-in modules of this repo)."""
    if not available():
        raise RuntimeError(f"reference not found at {REF} (run oracle/make_ref.py in the build container)")
    mods = []
    for name in ("lunar_generate", "lunar_evaluator"):
        spec = importlib.util.spec_from_file_location("_ref_" + name, _module_file(REF, name))
        m = importlib.util.module_from_spec(spec)
        sys.modules["_ref_" + name] = m
        spec.loader.exec_module(m)
        mods.append(m)
    return tuple(mods)
