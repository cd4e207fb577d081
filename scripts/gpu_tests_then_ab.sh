#!/bin/bash
# One gpurun call: the full GPU test suite on the in-tree build, then the same-box A/B against uncrtaints_b200/_ab_base.so.
mkdir -p gpurun_out
STEPS="tests smoke" bash scripts/gpu_r2.sh
ROUNDS=${ROUNDS:-2} bash scripts/gpu_ab_lib.sh
