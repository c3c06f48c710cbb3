"""Import the UNMODIFIED reference (MeryylleA/Lunaris-Orion) from /root/reference for oracle pinning and golden
generation. Only available in the build container; the GPU box has no /root/reference (use tests/golden there)."""
import importlib
import os
import sys

REF = os.environ.get("LUNARIS_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF, "lunar_evaluator.py"))


def load():
    """Returns (lunar_generate, lunar_evaluator) reference modules under private names (no sys.path pollution
    for the same-named drop-in modules of this repo)."""
    if not available():
        raise RuntimeError(f"reference not found at {REF}")
    mods = []
    for name in ("lunar_generate", "lunar_evaluator"):
        spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(REF, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["_ref_" + name] = m
        spec.loader.exec_module(m)
        mods.append(m)
    return tuple(mods)
