#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== attention tests"; timeout -k 10 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" -s 2>&1 | tail -60 > gpurun_out/g_attn_tests.log; tail -2 gpurun_out/g_attn_tests.log
echo "== attention microbench"; timeout -k 10 500 python profiles/microbench_attn.py 7 8 9 10 13 > gpurun_out/g_microbench_attn.txt 2>&1; head -7 gpurun_out/g_microbench_attn.txt
