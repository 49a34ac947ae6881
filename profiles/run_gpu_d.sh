#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== attention tests"; timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" -s 2>&1 | tail -40 > gpurun_out/d_attn_tests.log; tail -4 gpurun_out/d_attn_tests.log
echo "== attention microbench"; timeout -k 10 500 python profiles/microbench_attn.py > gpurun_out/d_microbench_attn.txt 2>&1; cat gpurun_out/d_microbench_attn.txt
