#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 3 50 python -u profiles/microbench_gn.py > gpurun_out/w_microbench_gn.txt 2>&1; echo "rc=$?"
