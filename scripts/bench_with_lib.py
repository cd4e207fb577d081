"""Development A/B helper: run bench.py against another build of the library on the SAME GPU box.

    python scripts/bench_with_lib.py uncrtaints_b200/_ab_base.so --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-parity

The product never does this: uncrtaints_b200/_lib.py loads libuncrtaints_b200.so next to itself and nothing else."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uncrtaints_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
