#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== attention variants"; DOVE_ATTN_NS=19426 timeout -k 10 300 python profiles/microbench_attn.py 8 12 13 14 8 12 13 14 2>&1 | tail -9
echo "== tiled bench"; timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tiled > gpurun_out/q_bench_tiled.log 2>&1; python - <<PY
import json
for l in open("gpurun_out/q_bench_tiled.log"):
    if l.startswith("{"):
        d=json.loads(l); f=d["families"]
        print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], {k:(round(x["ms_per_step"],1)) for k,x in f.items() if isinstance(x,dict)})
PY
