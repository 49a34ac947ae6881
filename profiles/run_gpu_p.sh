#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"conv_trans_halo_kernel" -s 2 -c 1 -o $O/r02f_conv_halo -f python profiles/ncu_targets.py conv > $O/p_ncu_conv.log 2>&1; tail -1 $O/p_ncu_conv.log
echo "== full tests"; timeout -k 10 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== tiled bench"; timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tiled > $O/p_bench_tiled.log 2>&1; python - <<PY
import json
for l in open("gpurun_out/p_bench_tiled.log"):
    if l.startswith("{"):
        d=json.loads(l); f=d["families"]
        print(d["value"], d["ms_per_step"], d["clocks"], {k:(round(x["ms_per_step"],1)) for k,x in f.items() if isinstance(x,dict)})
PY
