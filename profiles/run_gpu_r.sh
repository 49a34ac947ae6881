#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out

for g in 1; do
echo "== tiled bench graph=$g"; DOVE_BENCH_GRAPH=$g timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tiled > gpurun_out/r_bench_tiled$g.log 2>&1; python - <<PY
import json
ok=False
for l in open("gpurun_out/r_bench_tiled$g.log"):
    if l.startswith("{"):
        ok=True
        d=json.loads(l); f=d["families"]
        print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"], d["peak_mem_gb_rank0"])
if not ok: print(open("gpurun_out/r_bench_tiled$g.log").read()[-1500:])
PY
done
