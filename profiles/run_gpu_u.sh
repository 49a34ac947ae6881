#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 100 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention_variants and 8 and (4000 or 300 or 65)" > gpurun_out/u_memcheck_attn.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/u_memcheck_attn.log | head -5
