#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 3 75 python -u profiles/microbench.py > gpurun_out/v_microbench.txt 2>&1; echo "rc=$?"; wc -l gpurun_out/v_microbench.txt
