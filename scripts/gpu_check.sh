#!/bin/bash
# One gpurun call: GPU parity tests + smoke + a short bench; everything of interest lands in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.txt 2>&1
nproc >> gpurun_out/nvsmi.txt
timeout 120 python scripts/tc_probe.py > gpurun_out/tc_probe.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 240 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
for b in ${BENCH_BACKENDS:-0 1}; do
  timeout 900 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 --backend $b ${BENCH_ARGS} > gpurun_out/bench_b$b.log 2>&1
  echo "bench exit $?" >> gpurun_out/bench_b$b.log
done
cat gpurun_out/tc_probe.txt; tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; for b in ${BENCH_BACKENDS:-0 1}; do tail -c 1500 gpurun_out/bench_b$b.log; done
