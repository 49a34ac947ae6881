#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== h2d + layernorm tests"; timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "h2d or layernorm" 2>&1 | tail -3
echo "== nccl 2-rank test"; timeout -k 10 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s 2>&1 | tail -30 > gpurun_out/e_multigpu_test.log; tail -3 gpurun_out/e_multigpu_test.log
echo "== bench 2 gpus"; timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/e_bench_2gpu.log 2>&1; tail -c 600 gpurun_out/e_bench_2gpu.log
