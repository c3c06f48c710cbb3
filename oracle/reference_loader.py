"""Locate and import the UNMODIFIED reference (MeryylleA/Lunaris-Orion) for oracle pinning, golden generation and the
reference arm of bench.py. Resolution order: $LUNARIS_REFERENCE, /root/reference (build container: sources), then
oracle/_ref (CPython bytecode written by oracle/make_ref.py as *.rbc; the only form that exists on the GPU box).
TEST / BENCH INFRASTRUCTURE ONLY."""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get("LUNARIS_REFERENCE"), "/root/reference", os.path.join(_HERE, "_ref")]
MODULES = ("lunar_generate", "lunar_evaluator", "train_hybrid")


def _module_file(d, name):
    for ext in (".py", ".rbc"):
        p = os.path.join(d, name + ext)
        if os.path.isfile(p):
            return p
    return None


def _resolve():
    for d in _CANDIDATES:
        if d and all(_module_file(d, m) for m in MODULES):
            return d
    return _CANDIDATES[1]


REF = _resolve()


def available():
    return _module_file(REF, "lunar_evaluator") is not None


def kind():
    """'source' (build container) or 'bytecode' (oracle/_ref)."""
    f = _module_file(REF, "lunar_evaluator")
    return None if f is None else ("source" if f.endswith(".py") else "bytecode")


def _spec(name, fullname=None):
    path = _module_file(REF, name)
    fullname = fullname or name
    if path.endswith(".py"):
        return importlib.util.spec_from_file_location(fullname, path)
    loader = importlib.machinery.SourcelessFileLoader(fullname, path)
    return importlib.util.spec_from_file_location(fullname, path, loader=loader)


class _Finder(importlib.abc.MetaPathFinder):
    """Serves `import train_hybrid` / `lunar_generate` / `lunar_evaluator` from the reference directory. It sits at the
    END of sys.meta_path: anything importable through sys.path wins, which is how lunaris_orion_b200/dropin shadows
    the two model modules for the drop-in seam."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname in MODULES and path is None and _module_file(REF, fullname):
            return _spec(fullname)
        return None


_finder = None


def install_finder():
    global _finder
    if _finder is None:
        _finder = _Finder()
        sys.meta_path.append(_finder)


def load():
    """Returns (lunar_generate, lunar_evaluator) reference modules under private names (no sys.path pollution
    for the same-named drop-in modules of this repo)."""
    if not available():
        raise RuntimeError(f"reference not found at {REF} (run oracle/make_ref.py in the build container)")
    mods = []
    for name in ("lunar_generate", "lunar_evaluator"):
        spec = _spec(name, "_ref_" + name)
        m = importlib.util.module_from_spec(spec)
        sys.modules["_ref_" + name] = m
        spec.loader.exec_module(m)
        mods.append(m)
    return tuple(mods)
