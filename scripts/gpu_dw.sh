#!/bin/bash
# dwconv row-streaming kernels: parity tests, per-kernel timing of one block, and full bench under each mode
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 240 -k "row_streaming" > gpurun_out/pytest_rows.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_rows.log
tail -5 gpurun_out/pytest_rows.log
timeout 300 python scripts/bench_block.py --n 16 --groups 0 --modes ${MODES:-0,1,2,3,7} > gpurun_out/block_dec.log 2>&1; cat gpurun_out/block_dec.log
timeout 300 python scripts/bench_block.py --n 48 --groups 4 --modes ${MODES2:-0,3,7} > gpurun_out/block_enc.log 2>&1; cat gpurun_out/block_enc.log
for m in ${BENCH_MODES:-7}; do
  UB200_DWCONV_MODE=$m timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dw$m.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_dw$m.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('mode $m', j['value'], j['ms_per_step'], {k:v['ms'] for k,v in j['kernels_ms_per_step'].items() if 'dwconv' in k})
else:
    print(open('gpurun_out/bench_dw$m.log').read()[-1500:])
PY
done
