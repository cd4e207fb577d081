#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 240 -k "calibration or mgnll" > gpurun_out/pytest_cfg5.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_cfg5.log; tail -15 gpurun_out/pytest_cfg5.log
timeout 300 python scripts/bench_loss.py > gpurun_out/bench_loss.log 2>&1; cat gpurun_out/bench_loss.log
