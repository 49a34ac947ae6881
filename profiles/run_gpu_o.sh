#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== conv tests"; timeout -k 10 900 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x -k "conv or vae or one_step" 2>&1 | tail -5
echo "== bench"; timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/o_bench.log 2>&1; python - <<PY
import json
for l in open("gpurun_out/o_bench.log"):
    if l.startswith("{"):
        d=json.loads(l); f=d["families"]
        print(d["value"], d["ms_per_step"], d["clocks"], {k:(round(x["ms_per_step"],1)) for k,x in f.items() if isinstance(x,dict)})
        print([ (c["class"],round(c["ms_per_step"],1),c["launches_per_step"], c.get("tflops") and round(c["tflops"])) for c in d["top_classes"]])
PY
echo "== tiled bench"; timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tiled > gpurun_out/o_bench_tiled.log 2>&1; python - <<PY
import json
for l in open("gpurun_out/o_bench_tiled.log"):
    if l.startswith("{"):
        d=json.loads(l); f=d["families"]
        print(d["value"], d["ms_per_step"], d["clocks"], {k:(round(x["ms_per_step"],1)) for k,x in f.items() if isinstance(x,dict)})
        print([ (c["class"],round(c["ms_per_step"],1),c["launches_per_step"], c.get("tflops") and round(c["tflops"])) for c in d["top_classes"]])
PY
