#!/bin/bash
mkdir -p gpurun_out
LWB_NORM_R4=1 timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py tests/test_imitator_gpu.py tests/test_configs_gpu.py -m gpu -q -p no:cacheprovider --timeout 500 > gpurun_out/r4_tests.log 2>&1; echo "pytest(R4) rc=$?"; tail -2 gpurun_out/r4_tests.log
for v in 0 1 0 1; do
LWB_NORM_R4=$v timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r4_$v.log 2>&1; echo "r4=$v rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r4_$v.log").read().splitlines() if l.startswith("{")][-1])
print("r4=$v fps", round(d["value"],1), "ms", round(d["ms_per_step"],4), "norm", round(d["breakdown_ms_per_step"]["norm"],4), {k:v["ms"] for k,v in d["layers"].items() if k.startswith("norm") and "+" not in k})
PY
done
