#!/bin/bash
# ncu evidence: (1) per-launch device times of every launch of `bench.py --steps 1 --warmup 3` at the full batch,
# (2) --set full captures of selected kernels (batch 4).  The raw page of every capture is exported to CSV on the box; the
# .ncu-rep files themselves are kept only for the kernels listed in NCU_KEEP (gpurun_out/ is limited to 64 MiB).
mkdir -p gpurun_out
B=${NCU_BATCH:-16}
if [ "${NCU_LIST:-1}" = "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${NCU_SKIP:-0} -c ${NCU_COUNT:-1400} --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
fi
for k in ${NCU_KERNELS}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip ${NCU_FULL_SKIP:-0} -c ${NCU_FULL_COUNT:-2} -f \
      -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch ${NCU_FULL_BATCH:-4} > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?"
  ncu -i gpurun_out/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_$k.csv 2>/dev/null
  case " ${NCU_KEEP:-} " in *" $k "*) ;; *) rm -f gpurun_out/prof_$k.ncu-rep ;; esac
done
du -sh gpurun_out; ls gpurun_out | head -60
