#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== tests"; timeout -k 10 1500 python -m pytest tests -m gpu -q -s 2>&1 > $O/final_tests.log; grep -E "passed|failed" $O/final_tests.log | tail -2; grep -E "^FAILED|^ERROR|^E  " $O/final_tests.log | head -10
echo "== bench (default, with CPU baseline)"; DOVE_BENCH_CLASSES=$O/final_bench_classes.json timeout -k 10 1500 python bench.py > $O/final_bench_1gpu.log 2>&1; tail -c 300 $O/final_bench_1gpu.log
