#!/bin/bash
# Round-end evidence in one gpurun call: box facts, full GPU test suite, smoke, the driver's bench command with comparators, the
# reference arm, the loss bench, the ncu launch list of one step and --set full captures of the kernels named in NCU_KERNELS.
STEPS="${STEPS:-box tests smoke bench refarm loss}" BENCH_STEPS=20 bash scripts/gpu_r2.sh
NCU_LIST=1 NCU_KERNELS="${NCU_KERNELS:-residual_bwd_kernel dwrows_bwd2_kernel}" bash scripts/gpu_ncu2.sh
