#!/bin/bash
# Same-box A/B of two builds of the library: LIBS="path1 path2 ..." (default: the in-tree build and uncrtaints_b200/_ab_base.so), ROUNDS alternations.
mkdir -p gpurun_out
LIBS=${LIBS:-"uncrtaints_b200/_ab_base.so uncrtaints_b200/libuncrtaints_b200.so"}
for r in $(seq 1 ${ROUNDS:-2}); do
  for l in $LIBS; do
    tag=$(basename $l .so)_$r
    timeout 600 python scripts/bench_with_lib.py $l --steps ${BENCH_STEPS:-10} --warmup 3 --no-cpu-baseline --no-eager-baseline --no-parity ${BENCH_ARGS} > gpurun_out/bench_lib_$tag.log 2>&1
    echo "== $l (round $r)"; python scripts/show_bench.py gpurun_out/bench_lib_$tag.log | grep -v roofline | cut -c1-360
  done
done
