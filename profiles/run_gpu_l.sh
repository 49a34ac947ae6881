#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== bench tiled 1 gpu"; timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tiled > gpurun_out/l_bench_tiled.log 2>&1; tail -c 300 gpurun_out/l_bench_tiled.log
