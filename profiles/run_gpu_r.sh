#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tiled bench (graph value)"; timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tiled > gpurun_out/r_bench_tiled.log 2>&1; python - <<PY
import json
ok=False
for l in open("gpurun_out/r_bench_tiled.log"):
    if l.startswith("{"):
        ok=True
        d=json.loads(l); f=d["families"]
        print(d["value"], d["ms_per_step"], "e2e", d["e2e"], d["clocks"], d["peak_mem_gb_rank0"], d["gpu_launches"])
        print(d["families_source"]); print({k:(round(x["ms_per_step"],1)) for k,x in f.items() if isinstance(x,dict)})
if not ok: print(open("gpurun_out/r_bench_tiled.log").read()[-2500:])
PY
echo "== default bench"; timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench.log 2>&1; tail -c 300 gpurun_out/r_bench.log
