#!/bin/bash
# quick GPU loop: selected parity tests + bench under env-selected variants
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log; tail -12 gpurun_out/pytest_quick.log
IFS=';' read -ra VARS <<< "${VARIANTS:-X=1}"
for v in "${VARS[@]}"; do
  env $v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
  python - "$v" <<'PY'
import json, sys
l=[x for x in open('gpurun_out/bench_quick.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); k=j['kernels_ms_per_step']
    print(sys.argv[1], 'samples/s', j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'], 'loss', j['e2e']['last_loss'])
    print('   ', ' '.join(f"{a}={b['ms']:.2f}" for a,b in sorted(k.items(), key=lambda kv:-kv[1]['ms'])))
else:
    print(open('gpurun_out/bench_quick.log').read()[-2000:])
PY
  cp gpurun_out/bench_quick.log "gpurun_out/bench_quick_$(echo $v | tr '= ' '__').log"
done
