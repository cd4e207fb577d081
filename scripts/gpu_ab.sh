#!/bin/bash
# A/B loop of one gpurun call: for every environment assignment in VARIANTS (';'-separated) run the selected parity tests (PYTEST_K,
# optional) and a short bench, print samples/s and the per-kernel-class milliseconds.
mkdir -p gpurun_out
IFS=';' read -ra VARS <<< "${VARIANTS:-X=1}"
for v in "${VARS[@]}"; do
  tag=$(echo $v | tr '= ' '__')
  if [ -n "$PYTEST_K" ]; then
    env $v timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -k "$PYTEST_K" > gpurun_out/pytest_ab_$tag.log 2>&1
    echo "$v pytest exit $?: $(tail -1 gpurun_out/pytest_ab_$tag.log)"
  fi
  env $v timeout 600 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 --no-cpu-baseline --no-eager-baseline --no-parity ${BENCH_ARGS} > gpurun_out/bench_ab_$tag.log 2>&1
  echo "== $v"; python scripts/show_bench.py gpurun_out/bench_ab_$tag.log
done
