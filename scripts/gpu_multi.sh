#!/bin/bash
# N-GPU bench exactly as the driver launches it (one rank per GPU over NCCL), plus the N=1 line for comparison.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
echo "exit $?" >> gpurun_out/bench_n$N.log
tail -c 1200 gpurun_out/bench_n$N.log
