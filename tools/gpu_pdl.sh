#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pdl_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pdl_tests.log
for p in 0 1 0 1; do
LWB_PDL=$p timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_pdl$p.log 2>&1; echo "pdl=$p rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_pdl$p.log").read().splitlines() if l.startswith("{")][-1])
print("pdl=$p fps", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1))
PY
done
