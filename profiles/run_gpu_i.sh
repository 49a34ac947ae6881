#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 7 8; do timeout 120 profiles/bin/trace_attn $v > gpurun_out/i_trace_v$v.txt 2>&1; head -3 gpurun_out/i_trace_v$v.txt; done
bash profiles/run_gpu_g.sh
