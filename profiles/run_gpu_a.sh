#!/bin/bash
# round-2 GPU pass A: attention variants first (short timeouts), then the full GPU suite, microbenches, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
nproc >> gpurun_out/a_smi.txt; free -g >> gpurun_out/a_smi.txt
echo "== attention tests"; timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" -s 2>&1 | tail -60 > gpurun_out/a_attn_tests.log; tail -5 gpurun_out/a_attn_tests.log
echo "== attention microbench"; timeout -k 10 400 python profiles/microbench_attn.py > gpurun_out/a_microbench_attn.txt 2>&1; cat gpurun_out/a_microbench_attn.txt
echo "== gn microbench"; timeout -k 10 200 python profiles/microbench_gn.py > gpurun_out/a_microbench_gn.txt 2>&1; cat gpurun_out/a_microbench_gn.txt
echo "== full gpu suite"; timeout -k 10 1500 python -m pytest tests -m gpu -q -s 2>&1 > gpurun_out/a_tests_full.log; grep -E "passed|failed|error" gpurun_out/a_tests_full.log | tail -5; grep -E "^FAILED|^ERROR" gpurun_out/a_tests_full.log | head -30
echo "== bench"; timeout -k 10 900 python bench.py --steps 3 --warmup 3 > gpurun_out/a_bench.log 2>&1; tail -c 6000 gpurun_out/a_bench.log
