"""Launcher: ``python -m uncrtaints_b200.run train_reconstruct.py --flags...`` (cwd = the reference's model/ dir).

Installs the B200 path into the reference's ``src`` package, then executes the unmodified script with runpy.
"""
from __future__ import annotations

import os
import runpy
import sys


def main() -> None:
    if len(sys.argv) < 2:
        print(__doc__)
        raise SystemExit(2)
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)) or os.getcwd())
    from .install import install
    install()
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
