#!/bin/bash
# One gpurun call of round 2: box facts, GPU tests, smoke, the bench line with its comparators, the reference arm, the loss bench.
# Sections are selected with STEPS (default: all).  Everything of interest lands in gpurun_out/.
STEPS=${STEPS:-"box tests smoke bench refarm loss"}
mkdir -p gpurun_out
for s in $STEPS; do
case $s in
box)
  { nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv; nproc; free -g; grep -m1 "model name" /proc/cpuinfo; } > gpurun_out/box.txt 2>&1 ;;
tests)
  rm -f gpurun_out/parity_report.txt
  timeout ${TEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 --durations=15 ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log ;;
smoke)
  timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log ;;
bench)
  timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_main.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_main.log
  tail -c 6000 gpurun_out/bench_main.log ;;
benchq)
  timeout 600 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 --no-cpu-baseline --no-eager-baseline --no-parity ${BENCH_ARGS} > gpurun_out/bench_quick.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_quick.log
  python scripts/show_bench.py gpurun_out/bench_quick.log ;;
refarm)
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_cpu.log 2>&1; tail -c 1500 gpurun_out/bench_reference_cpu.log ;;
loss)
  timeout 300 python scripts/bench_loss.py > gpurun_out/bench_loss.log 2>&1; tail -5 gpurun_out/bench_loss.log ;;
bench3)
  BENCH_ARGS="--backend 3" ; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-parity --backend 3 > gpurun_out/bench_b3.log 2>&1; python scripts/show_bench.py gpurun_out/bench_b3.log ;;
bench11)
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --backend 3 > gpurun_out/bench_b11.log 2>&1; python scripts/show_bench.py gpurun_out/bench_b11.log; grep -o '"parity": {[^}]*}' gpurun_out/bench_b11.log ;;
bench27)
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-parity --backend 3 > gpurun_out/bench_b27.log 2>&1; python scripts/show_bench.py gpurun_out/bench_b27.log ;;
ncu)
  for k in ${NCU_KERNELS}; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip ${NCU_SKIP:-6} -c ${NCU_COUNT:-1} -f \
        -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-parity --batch ${NCU_BATCH:-4} ${BENCH_ARGS} > gpurun_out/ncu_$k.log 2>&1
    echo "ncu $k exit $?"
    ncu -i gpurun_out/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_$k.csv 2>/dev/null
    ncu -i gpurun_out/prof_$k.ncu-rep --page details > gpurun_out/prof_$k.txt 2>/dev/null
    [ "${NCU_SOURCE:-0}" = "1" ] && ncu -i gpurun_out/prof_$k.ncu-rep --page source --csv > gpurun_out/prof_${k}_source.csv 2>/dev/null
    [ "${NCU_KEEP:-0}" = "1" ] || rm -f gpurun_out/prof_$k.ncu-rep
  done ;;
aux)
  timeout 600 python scripts/bench_aux.py > gpurun_out/bench_aux.log 2>&1; tail -c 1500 gpurun_out/bench_aux.log ;;
cfg3)
  timeout 900 python bench.py --steps 5 --warmup 3 --batch 32 --t 5 --no-cpu-baseline --no-eager-baseline --no-parity ${CFG3_ARGS} > gpurun_out/bench_cfg3.log 2>&1; python scripts/show_bench.py gpurun_out/bench_cfg3.log ;;
esac
done
