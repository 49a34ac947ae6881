#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== graph + latent tests"; timeout -k 10 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -x -k "graph_replay or latent_frames" 2>&1 | tail -3
echo "== memcheck (halo conv, attention v4 at 4000 rows)"; timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "(conv_cta_pair and (512 or 496 or 368) and 1-) or (attention_variants and 4000 and 8-)" > $O/s_memcheck.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" $O/s_memcheck.log | head -8
K='regex:attn_fwd|conv2cta|conv_narrow|conv_trans_halo|gaussian_sample|gemm2cta|gemv_kernel|gn_apply|gn_finalize|gn_stats|gn_partial|layernorm_mod|ncthw_to_cl|cl_to_ncthw|patchify|pixels_to_cl|time_pool|umma_gemm|unpatchify|upsample_nearest|post_scale|velocity|blend|qk_norm'
echo "== ncu launch list (library kernels; warm-up step + timed step)"; timeout -k 10 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 4000 --csv --log-file $O/final_launches.csv python bench.py --steps 1 --warmup 1 --profile --no-cpu-baseline > $O/final_launches_bench.log 2>&1; tail -c 200 $O/final_launches_bench.log; wc -l $O/final_launches.csv
