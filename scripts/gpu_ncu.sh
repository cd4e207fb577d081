#!/bin/bash
# ncu evidence (B200_PROFILING.md): per-launch device times of one bench step, and --set full captures of the top kernels.
mkdir -p gpurun_out
B=${NCU_BATCH:-16}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-840} -c ${NCU_COUNT:-420} --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch $B > gpurun_out/ncu_launches.log 2>&1
for k in ${NCU_KERNELS:-dwconv_bwd gemm_tc_kernel wgrad_tc dwconv_fwd dh2_kernel}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_FULL_SKIP:-6} -c ${NCU_FULL_COUNT:-4} -f \
      -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 4 > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
ls -la gpurun_out/
