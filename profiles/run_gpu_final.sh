#!/bin/bash
# Round-2 final evidence run (1 GPU): tests, bench arms, launch list, ncu captures, attention sweep.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/final_smoke.log
echo "== tests"; timeout -k 10 1500 python -m pytest tests -m gpu -q -s 2>&1 > $O/final_tests.log; grep -E "passed|failed" $O/final_tests.log | tail -3; grep -E "^FAILED|^ERROR|^E  " $O/final_tests.log | head -20
echo "== bench (default, with CPU baseline)"; DOVE_BENCH_CLASSES=$O/final_bench_classes.json timeout -k 10 1500 python bench.py > $O/final_bench_1gpu.log 2>&1; tail -c 400 $O/final_bench_1gpu.log
echo "== bench --impl reference"; timeout -k 10 900 python bench.py --impl reference --steps 1 --warmup 0 > $O/final_bench_reference.log 2>&1; tail -c 600 $O/final_bench_reference.log
echo "== bench tiled"; timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tiled > $O/final_bench_tiled.log 2>&1; tail -c 200 $O/final_bench_tiled.log
echo "== gpu_library"; timeout -k 10 900 python bench.py --impl gpu_library --steps 2 --warmup 1 > $O/final_gpu_library.log 2>&1; tail -c 700 $O/final_gpu_library.log
echo "== gpu_library channels_last"; timeout -k 10 900 python bench.py --impl gpu_library --steps 2 --warmup 1 --channels-last > $O/final_gpu_library_cl.log 2>&1; tail -c 700 $O/final_gpu_library_cl.log
echo "== attention sweep"; DOVE_ATTN_NS=19426,4978,3216,32866 timeout -k 10 600 python profiles/microbench_attn.py 0 2 8 > $O/final_microbench_attn.txt 2>&1; cat $O/final_microbench_attn.txt
K='regex:attn_fwd|conv2cta|conv_narrow|conv_trans_halo|gaussian_sample|gemm2cta|gemv_kernel|gn_apply|gn_finalize|gn_stats|gn_partial|layernorm_mod|ncthw_to_cl|cl_to_ncthw|patchify|pixels_to_cl|time_pool|umma_gemm|unpatchify|upsample_nearest|post_scale|velocity|blend|qk_norm'
echo "== ncu launch list (library kernels; warm-up step + timed step)"; timeout -k 10 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 4000 --csv --log-file $O/final_launches.csv python bench.py --steps 1 --warmup 1 --profile --no-cpu-baseline > $O/final_launches_bench.log 2>&1; tail -c 200 $O/final_launches_bench.log; wc -l $O/final_launches.csv
echo "== ncu conv"; timeout -k 10 600 ncu --set full --clock-control none -k regex:"conv_trans_halo_kernel" -s 2 -c 1 -o $O/r02f_conv_trans -f python profiles/ncu_targets.py conv > $O/final_ncu_conv.log 2>&1; tail -1 $O/final_ncu_conv.log
timeout -k 10 600 ncu --set full --clock-control none -k regex:"conv2cta_kernel" -s 2 -c 1 -o $O/r02f_conv2cta -f python profiles/ncu_targets.py conv > $O/final_ncu_conv2.log 2>&1; tail -1 $O/final_ncu_conv2.log
ls -la $O/*.ncu-rep | tail -4
