#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== targeted tests"; timeout -k 10 900 python -m pytest tests -m gpu -q -s -k "reference_process or per_frame or qkv or groupnorm or attention or one_step or surface or test_dit or fused_post" 2>&1 > gpurun_out/b_tests.log; grep -E "passed|failed" gpurun_out/b_tests.log | tail -3; grep -E "^FAILED|^ERROR|gemm_qkv|groupnorm C" gpurun_out/b_tests.log | head -30
echo "== ncu attention"; timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 6 -o gpurun_out/r02_attn_v123 -f python profiles/ncu_targets.py attn 1 2 3 > gpurun_out/b_ncu_attn.log 2>&1; tail -3 gpurun_out/b_ncu_attn.log
echo "== ncu conv"; timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"umma_gemm_kernel|conv2cta_kernel" -c 6 -o gpurun_out/r02_conv -f python profiles/ncu_targets.py conv > gpurun_out/b_ncu_conv.log 2>&1; tail -3 gpurun_out/b_ncu_conv.log
echo "== ncu gn"; timeout -k 10 300 ncu --set full --clock-control none -k regex:gn_apply -c 8 -o gpurun_out/r02_gn -f python profiles/ncu_targets.py gn > gpurun_out/b_ncu_gn.log 2>&1; tail -2 gpurun_out/b_ncu_gn.log
echo "== microbench"; timeout -k 10 300 python profiles/microbench_gn.py > gpurun_out/b_microbench_gn.txt 2>&1; cat gpurun_out/b_microbench_gn.txt
echo "== bench"; timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench.log 2>&1; tail -c 5000 gpurun_out/b_bench.log
ls -la gpurun_out/*.ncu-rep
