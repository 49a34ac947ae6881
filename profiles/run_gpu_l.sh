#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== tests"; timeout -k 10 1500 python -m pytest tests -m gpu -q -s 2>&1 > gpurun_out/l_tests.log; grep -E "passed|failed" gpurun_out/l_tests.log | tail -3; grep -E "^FAILED|^ERROR|^E  " gpurun_out/l_tests.log | head -20
echo "== bench tiled 1 gpu"; timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tiled > gpurun_out/l_bench_tiled.log 2>&1; tail -c 300 gpurun_out/l_bench_tiled.log
echo "== bench"; DOVE_BENCH_CLASSES=gpurun_out/l_bench_classes.json timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/l_bench.log 2>&1; tail -c 600 gpurun_out/l_bench.log
