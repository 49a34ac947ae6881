#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
K='regex:attn_fwd|conv2cta|conv_narrow|conv_trans_halo|gaussian_sample|gemm2cta|gemv_kernel|gn_apply|gn_finalize|gn_stats|gn_partial|layernorm_mod|ncthw_to_cl|cl_to_ncthw|patchify|pixels_to_cl|time_pool|umma_gemm|unpatchify|upsample_nearest|post_scale|velocity|blend|qk_norm'
echo "== ncu launch list (library kernels; warm-up step + timed step)"; timeout -k 10 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 4000 --csv --log-file $O/final_launches.csv python bench.py --steps 1 --warmup 1 --profile --no-cpu-baseline > $O/final_launches_bench.log 2>&1; tail -c 200 $O/final_launches_bench.log; wc -l $O/final_launches.csv
