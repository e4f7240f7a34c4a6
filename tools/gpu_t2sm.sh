#!/bin/bash
mkdir -p gpurun_out
for v in 1 2 1 2; do
LWB_2SM=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_2sm$v.log 2>&1; echo "2sm=$v rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_2sm$v.log").read().splitlines() if l.startswith("{")][-1])
print("2sm=$v fps", round(d["value"],1), "ms", round(d["ms_per_step"],4), {k:v for k,v in d["layers"].items() if k.startswith("T")})
PY
done
