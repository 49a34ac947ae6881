#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
echo "== bench cfg2 $N gpus"; timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/k_bench_cfg2_${N}gpu.log 2>&1; tail -c 3000 gpurun_out/k_bench_cfg2_${N}gpu.log
echo "== bench cfg4 $N gpus"; timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload cfg4 --steps 3 --warmup 3 > gpurun_out/k_bench_cfg4_${N}gpu.log 2>&1; tail -c 3000 gpurun_out/k_bench_cfg4_${N}gpu.log
