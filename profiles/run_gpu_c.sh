#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== targeted tests"; timeout -k 10 900 python -m pytest tests -m gpu -q -s -k "reference_process or per_frame or qkv or groupnorm or one_step or surface or test_dit or fused_post or same_seed" 2>&1 > gpurun_out/c_tests.log; grep -E "passed|failed" gpurun_out/c_tests.log | tail -3; grep -E "^FAILED|^ERROR|gemm_qkv|groupnorm C|^E  " gpurun_out/c_tests.log | head -40
echo "== ncu attention"; timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/r02_attn_v2emu1 -f python profiles/ncu_targets.py attn 2 > gpurun_out/c_ncu_attn.log 2>&1; tail -2 gpurun_out/c_ncu_attn.log
timeout -k 10 600 ncu --set full --clock-control none -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/r02_attn_v3emu2 -f python profiles/ncu_targets.py attn 3 > gpurun_out/c_ncu_attn3.log 2>&1; tail -2 gpurun_out/c_ncu_attn3.log
echo "== ncu conv"; timeout -k 10 600 ncu --set full --clock-control none -k regex:"umma_gemm_kernel|conv2cta_kernel" -s 1 -c 1 -o gpurun_out/r02_conv_trans -f python profiles/ncu_targets.py conv > gpurun_out/c_ncu_conv.log 2>&1; tail -2 gpurun_out/c_ncu_conv.log
timeout -k 10 600 ncu --set full --clock-control none -k regex:"conv2cta_kernel" -s 1 -c 1 -o gpurun_out/r02_conv2cta -f python profiles/ncu_targets.py conv > gpurun_out/c_ncu_conv2.log 2>&1; tail -2 gpurun_out/c_ncu_conv2.log
echo "== bench"; DOVE_BENCH_CLASSES=gpurun_out/c_bench_classes.json timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench.log 2>&1; tail -c 4500 gpurun_out/c_bench.log
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
