"""Install the UNMODIFIED reference (PatrickTUM/UnCRtainTS) under git-ignored ``baseline/_ref/``.

The reference is a pure-Python source tree without packaging metadata (no setup.py / pyproject.toml), so the
``pip install --target baseline/_ref /root/reference`` of the bench contract has nothing to build: "installing" it
is copying its three Python packages (``model/`` incl. ``model/src``, ``data/``, ``util/``) verbatim.  The copy is
git-ignored (the history holds none of the reference's sources) but NOT gpurun-ignored, so it travels to the GPU box,
where /root/reference does not exist.  ``bench.py --impl reference`` and the unmodified-loop GPU test import it from there.

    python baseline/install_ref.py            # in the build container; no-op if /root/reference is absent
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("UNCRTAINTS_REFERENCE", "/root/reference")


def install(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(SRC, "model", "src")):
        if verbose:
            print(f"[install_ref] {SRC} not present; keeping whatever is under {DEST}")
        return os.path.isdir(os.path.join(DEST, "model", "src"))
    for sub in ("model", "data", "util"):
        dst = os.path.join(DEST, sub)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, sub), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(DEST))
        print(f"[install_ref] copied {n} files of the unmodified reference into {DEST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
