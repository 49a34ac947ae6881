#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 13; do
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/r02_attn_v5_var$v -f python profiles/ncu_targets.py attn $v > gpurun_out/h_ncu_attn.log 2>&1; tail -1 gpurun_out/h_ncu_attn.log
done
