#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== bench cfg4 1 gpu"; timeout -k 10 1200 python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_cfg4_1gpu.log 2>&1; tail -c 1500 gpurun_out/m_bench_cfg4_1gpu.log
