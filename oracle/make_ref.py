"""Build oracle/_ref/: the UNMODIFIED reference (MeryylleA/Lunaris-Orion) byte-compiled from the sources where they
lie under /root/reference, so the real reference trainer travels to the GPU box (which has no /root/reference).

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

Outputs only compiled artefacts - sourceless CPython bytecode of the three modules on the hot path (lunar_generate,
lunar_evaluator, train_hybrid), stored as `*.rbc` (plain .pyc contents; the .pyc suffix is filtered out of gpurun
snapshots) - into the git-ignored oracle/_ref/. No reference source is copied into the repository or its history. The
GPU box runs the same image (same CPython), so the bytecode imports there through oracle/reference_loader.py.
TEST / BENCH INFRASTRUCTURE ONLY: consumers are tests/, bench.py's `--impl reference` and `cpu_baseline` legs.
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("LUNARIS_REFERENCE_SRC", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = ("lunar_generate", "lunar_evaluator", "train_hybrid")


def stage(verbose=False):
    """Returns the output directory, or None when the reference sources are not present (GPU box: prebuilt files)."""
    if not all(os.path.isfile(os.path.join(SRC, m + ".py")) for m in MODULES):
        return None
    os.makedirs(OUT, exist_ok=True)
    for m in MODULES:
        dst = os.path.join(OUT, m + ".rbc")
        py_compile.compile(os.path.join(SRC, m + ".py"), cfile=dst, dfile=f"<reference>/{m}.py", doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        if verbose:
            print("compiled", dst)
    with open(os.path.join(OUT, "PYTHON_VERSION"), "w") as f:
        f.write(sys.version.split()[0] + " " + sys.implementation.cache_tag + "\n")
    return OUT


if __name__ == "__main__":
    print(stage(verbose=True) or f"reference sources not found under {SRC}")
