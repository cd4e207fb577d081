#!/bin/bash
# Multi-GPU evidence (run with `gpurun --gpus N`): NCCL parity test, the bench at BASELINE config #2's per-GPU batch and at
# config #4's (32 per GPU), exactly as the driver launches it (one rank per GPU under torchrun).
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
[ "${SKIP_TEST:-0}" = "1" ] || { timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi.log; }
tail -3 gpurun_out/pytest_multi.log; cat gpurun_out/nccl_parity.txt
for B in ${BATCHES:-16 32}; do
  NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 --batch $B > gpurun_out/bench_n${N}_b$B.log 2>&1
  echo "exit $?" >> gpurun_out/bench_n${N}_b$B.log
  python scripts/show_bench.py gpurun_out/bench_n${N}_b$B.log | head -2
  grep -m3 -i "NVLS\|via P2P\|Connected all" gpurun_out/bench_n${N}_b$B.log
done
