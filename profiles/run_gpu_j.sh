#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 12 2 8; do
echo "== bench attn_variant $v"; DOVE_ATTN_VARIANT=$v timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench_v$v.log 2>&1; python - <<PY
import json
for l in open("gpurun_out/j_bench_v$v.log"):
    if l.startswith("{"):
        d=json.loads(l); f=d["families"]
        print(d["value"], d["ms_per_step"], d["clocks"], {k:(round(x["ms_per_step"],1)) for k,x in f.items() if isinstance(x,dict)})
PY
done
