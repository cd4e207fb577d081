"""Print the key fields of the last JSON line of a bench log:  python scripts/show_bench.py gpurun_out/bench_n2.log"""
import json
import sys

for path in sys.argv[1:]:
    lines = [x for x in open(path) if x.startswith("{")]
    if not lines:
        print(path, "no JSON line; tail:", open(path).read()[-800:])
        continue
    j = json.loads(lines[-1])
    print(path, {k: j.get(k) for k in ("value", "n_gpus", "ms_per_step", "e2e", "clocks", "gpu_launches")})
    if "roofline" in j:
        print("   roofline", {k: j["roofline"].get(k) for k in ("kernel", "achieved", "frac", "traffic", "share_of_step")})
    if "kernels_ms_per_step" in j:
        k = j["kernels_ms_per_step"]
        print("   ", " ".join(f"{a}={b['ms']:.2f}" for a, b in sorted(k.items(), key=lambda kv: -kv[1]["ms"])))
