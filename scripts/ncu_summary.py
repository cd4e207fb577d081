"""Summarise ncu captures brought back in gpurun_out/ into tracked files under profiles/.

    python scripts/ncu_summary.py r01 [--full-batch 4]

Reads gpurun_out/launches.csv (`ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 3`,
every launch of the process) and gpurun_out/prof_*.ncu-rep (`ncu --set full --clock-control none --import-source on`, captured
with `bench.py --batch <full-batch>`), writes
  profiles/<tag>_launches.md / .csv   one training step (the launches between two consecutive in_conv moment kernels, i.e. from
                                      the first kernel of a forward pass to the last kernel of its backward pass)
  profiles/<tag>_ncu_full_summary.md  per-kernel table of the --set full captures
  profiles/<tag>_traffic.json         dram__bytes_read.sum + dram__bytes_write.sum per launch and per frame (what bench.py's
                                      roofline.traffic is derived from)
"""
import csv
import glob
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
full_batch = int(sys.argv[sys.argv.index("--full-batch") + 1]) if "--full-batch" in sys.argv else 4
T = 3
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
KEYS = [
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct"),
    ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
]
traffic = {}
lines = [f"# ncu --set full summaries ({tag})", "",
         f"Captured with `scripts/gpu_ncu2.sh` (`ncu --set full --clock-control none --import-source on`, `bench.py --batch {full_batch} "
         f"--steps 1`: decoder launches cover {full_batch} frames, encoder launches {full_batch * T});",
         "times are cold-cache single launches under the profiler: compare shares and percentages, not absolutes. Units as printed by "
         "ncu.", ""]
reps = {os.path.splitext(p)[0] for p in glob.glob(os.path.join(ROOT, "gpurun_out", "prof_*.ncu-rep"))}
reps |= {os.path.splitext(p)[0] for p in glob.glob(os.path.join(ROOT, "gpurun_out", "prof_*.csv"))}
for stem in sorted(reps):
    rep = stem + ".ncu-rep"
    if os.path.exists(stem + ".csv") and os.path.getsize(stem + ".csv") > 0:      # raw page exported on the GPU box
        raw = open(stem + ".csv").read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    lines += [f"## {os.path.basename(rep)}", "", "| kernel | " + " | ".join(k for _, k in KEYS) + " | top stalls |", "|---|" + "---|" * (len(KEYS) + 1)]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?").split("(")[0][:70].strip()
        try:
            def tobytes(key):
                v, un = float(d[key]), u.get(key, "").lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(un, 1)
            traffic.setdefault(name, []).append({"dram_bytes": tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"),
                                                 "grid": d.get("launch__grid_size", "")})
        except (KeyError, ValueError):
            pass
        vals = []
        for k, _ in KEYS:
            v = d.get(k, "")
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            vals.append(f"{v} {u.get(k, '')}".strip() if k.startswith(("gpu__time", "dram__bytes")) else v)
        st = [(float(v), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h, v in d.items()
              if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and v not in ("", "n/a")]
        tot = sum(x for x, _ in st) or 1.0
        top = ", ".join(f"{h} {100 * x / tot:.0f}%" for x, h in sorted(st, reverse=True)[:4])
        lines.append(f"| {name} | " + " | ".join(vals) + f" | {top} |")
    lines.append("")
if len(lines) > 6:
    open(os.path.join(out_dir, f"{tag}_ncu_full_summary.md"), "w").write("\n".join(lines))
    # per-frame traffic of the row-streaming depthwise kernels and the GEMMs: a decoder launch covers `full_batch` frames
    # A launch covers either the decoder batch (full_batch frames) or the encoder batch (full_batch * T frames); which one a
    # captured launch was is decided by the algorithmic bytes per frame where the kernel has an entry.
    A, Hh = 128 * 65536 * 4, 256 * 65536 * 4
    ALG = {"dwrows_bwd2_kernel": 4 * Hh, "dwrows_fwd_kernel": 2 * Hh, "gemm_tc_kernel<128, 256, TLoadNormed": A + Hh,
           "gemm_tc_kernel<256, 128, TLoadGeluGate": Hh + A, "gemm_tc_kernel<128, 256, TLoadNormBwd": 2 * A + 2 * Hh,
           "gemm_tc_kernel<256, 128, TLoadNormBwd": 2 * Hh + 2 * A, "bwd_tc_kernel<1": 2 * Hh + 2 * A, "bwd_tc_kernel<2": 2 * Hh + 2 * A,
           "wgrad_tc_kernel<TLoadNormBwd, TLoadGeluGate>": 2 * A + Hh,
           "wgrad_tc_kernel<TLoadNormed, TLoadNormBwd>": A + 2 * Hh, "se_pool_kernel": Hh, "residual_bwd_kernel<1>": 5 * A, "residual_bwd_kernel": 4 * A,
           "residual_fwd_kernel": 3 * A, "norm_bwd_stats_kernel": 2 * A}
    per_frame = {}
    for name, ls in traffic.items():
        small = min(x["dram_bytes"] for x in ls)
        alg = next((v for k, v in ALG.items() if k in name), None)
        frames = full_batch
        if alg is not None and abs(small / (full_batch * T) - alg) < abs(small / full_batch - alg):
            frames = full_batch * T
        per_frame[name] = small / frames
    json.dump({"note": f"per-launch dram__bytes_read.sum + dram__bytes_write.sum from ncu --set full (bench.py --batch {full_batch}: "
                       f"decoder launches = {full_batch} frames, encoder launches = {full_batch * T} frames); per_frame = smallest launch / "
                       f"{full_batch}", "per_frame_bytes": per_frame, "kernels": traffic},
              open(os.path.join(out_dir, f"{tag}_traffic.json"), "w"), indent=1)

src = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(src):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    body = rows[hi + 1:]
    marks = [i for i, r in enumerate(body) if "inconv_moments_kernel" in r[ki]]
    # bench.py --steps 1 --warmup 3: steps 0-2 warm-up, step 3 the device-resident timed step
    lo, hi2 = (marks[3], marks[4]) if len(marks) >= 5 else (marks[-2], marks[-1])
    step = body[lo:hi2]
    agg = {}
    for r in step:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)   # -> us
        name = r[ki].split("(")[0][:80]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 3`\n\n")
        f.write(f"ONE training step at B=16, T=3 (launches {lo}..{hi2 - 1} of the process: from the first kernel of the 4th forward pass to the "
                f"last kernel of its backward pass; {len(step)} launches, {total / 1e3:.2f} ms summed).  Per-launch times are cold-cache and "
                "serialised by the profiler; the SHARE column is what is comparable with bench.py's CUDA-event breakdown "
                "(`kernels_ms_per_step`).\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {name} | {cnt} | {us:.1f} | {100 * us / total:.1f}% |\n")
    with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(hdr)
        w.writerows(step)
print("written", sorted(os.listdir(out_dir)))
