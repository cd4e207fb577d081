#!/bin/bash
# ncu --set full of selected kernels inside one standalone MBConv block (bench_block.py), N=16 decoder shape
mkdir -p gpurun_out
K=${NCU_K:-dwrows}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-2} -f \
   -o gpurun_out/prof_blk_$K python scripts/bench_block.py --n ${NCU_N:-16} --modes ${NCU_MODE:-7} --reps 1 > gpurun_out/ncu_blk_$K.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_blk_$K.log; ls -la gpurun_out/*.ncu-rep
