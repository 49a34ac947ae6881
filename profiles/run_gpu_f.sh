#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"; timeout -k 10 1200 python -m pytest tests -m gpu -q -s -k "not cfg1_full and not fullsize" 2>&1 > gpurun_out/f_tests.log; grep -E "passed|failed" gpurun_out/f_tests.log | tail -3; grep -E "^FAILED|^ERROR|conv narrow|^E  " gpurun_out/f_tests.log | head -40
echo "== bench"; DOVE_BENCH_CLASSES=gpurun_out/f_bench_classes.json timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench.log 2>&1; tail -c 3500 gpurun_out/f_bench.log
echo "== gpu_library"; timeout -k 10 900 python bench.py --impl gpu_library --steps 2 --warmup 1 > gpurun_out/f_gpu_library.log 2>&1; tail -c 1500 gpurun_out/f_gpu_library.log
echo "== gpu_library channels_last"; timeout -k 10 900 python bench.py --impl gpu_library --steps 2 --warmup 1 --channels-last > gpurun_out/f_gpu_library_cl.log 2>&1; tail -c 1500 gpurun_out/f_gpu_library_cl.log
