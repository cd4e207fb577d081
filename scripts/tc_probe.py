"""Probe the tcgen05 GEMM with alternative descriptor encodings (each variant in its own process, with a timeout, so a
faulting or hanging variant cannot take the others down).  Usage: python scripts/tc_probe.py  -> gpurun_out/tc_probe.txt"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from uncrtaints_b200 import _lib
L = _lib.lib()
desc_hi, lbo, idesc = (int(a, 0) for a in sys.argv[1:4])
_lib.check(L.ub200_tc_debug_set(desc_hi, lbo, idesc), "set")
N, P = 1, 512
g = torch.Generator("cpu").manual_seed(31)
x = torch.randn(N, P, 128, generator=g)
coef = torch.stack([torch.ones(N, 128), torch.zeros(N, 128)], dim=-1).contiguous()
w1 = torch.randn(256, 128, generator=g) * 0.1
ref = x.double() @ w1.double().t()
xd, cd, wd = x.cuda(), coef.cuda(), w1.cuda()
h1 = torch.zeros(N, P, 256, device="cuda")
stats = torch.zeros(N, 256, 2, dtype=torch.float64, device="cuda")
scratch = torch.empty(256 * 1024, dtype=torch.uint8, device="cuda")
_lib.check(L.ub200_gemm1_forward(1, xd.data_ptr(), cd.data_ptr(), wd.data_ptr(), h1.data_ptr(), stats.data_ptr(), N, P,
                                 scratch.data_ptr(), torch.cuda.current_stream().cuda_stream), "gemm1")
torch.cuda.synchronize()
err = float((h1.double().cpu() - ref).norm() / ref.norm())
print("rel_l2=%%.3e  max|h1|=%%.3f  nan=%%d" %% (err, float(h1.abs().max()), int(torch.isnan(h1).any())))
''' % ROOT

SBO, V1, SW128 = 64, 1 << 14, 2 << 29
IDESC = (1 << 4) | (1 << 7) | (1 << 10) | (16 << 17) | (8 << 24)
variants = [
    ("default (SBO=64, version=1, SW128, LBO=1)", SBO | V1 | SW128, 1, IDESC),
    ("LBO=0", SBO | V1 | SW128, 0, IDESC),
    ("version=0", SBO | SW128, 1, IDESC),
    ("LBO=64", SBO | V1 | SW128, 64, IDESC),
]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "tc_probe.txt"), "w") as out:
    for name, hi, lbo, idesc in variants:
        try:
            r = subprocess.run([sys.executable, "-c", CHILD, hex(hi), hex(lbo), hex(idesc)], capture_output=True, text=True, timeout=90)
            msg = (r.stdout.strip().splitlines() or ["<no output>"])[-1] + (f"  [rc={r.returncode}] " + r.stderr.strip()[-300:] if r.returncode else "")
        except subprocess.TimeoutExpired:
            msg = "TIMEOUT (hang)"
        line = f"{name}: {msg}"
        print(line)
        out.write(line + "\n")
        out.flush()
