#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py -m gpu -q -s -p no:cacheprovider --timeout 600 -k "fp16f8 or conv2d or inference_matches" > gpurun_out/f8_tests.log 2>&1; echo "pytest rc=$?"
grep -h "fp16f8\|passed\|failed\|Error\|error" gpurun_out/f8_tests.log | head -60
LWB_PRECISION=fp16f8 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_f8.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_f8.log").read().splitlines() if l.startswith("{")][-1])
    print("f8 fps", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d.get("breakdown_ms_per_step"), d.get("roofline",{}).get("frac"), d.get("fast_mode"))
    for k,v in d.get("layers",{}).items(): print("   ", k, v)
except Exception as e: print("parse error", e); print(open("gpurun_out/bench_f8.log").read()[-3000:])
PY
