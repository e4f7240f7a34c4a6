#!/bin/bash
# GPU run 5: validate halo conv variant + new norm/raster kernels, then A/B the halo modes.
mkdir -p gpurun_out
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 900"
timeout 900 $PYT tests/test_conv_gpu.py -k "halo or stem" -s > gpurun_out/halo.log 2>&1; echo "halo rc=$?"
grep -E "passed|failed|halo|rowk" gpurun_out/halo.log | cut -c1-160 | tail -50
timeout 900 $PYT tests/test_raster_gpu.py tests/test_conv_gpu.py -k "not halo" > gpurun_out/kern.log 2>&1; echo "kern rc=$?"; tail -3 gpurun_out/kern.log
for mode in 0 auto all; do
  LWB_HALO=$mode timeout 600 $PYT tests/test_generator_gpu.py -s > gpurun_out/gen_$mode.log 2>&1; echo "gen[$mode] rc=$?"; grep -E "golden|oracle|passed|failed" gpurun_out/gen_$mode.log | head -12
  LWB_HALO=$mode timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$mode.log 2>&1; echo "bench[$mode] rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$mode.log").read().strip().splitlines()[-1])
    print("$mode", "fps", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d.get("breakdown_ms_per_step"), "conv frac", round(d["roofline"]["frac"],3))
    for k,v in d.get("layers",{}).items(): print("   ", k, v)
except Exception as e: print("$mode", "parse error", e); print(open("gpurun_out/bench_$mode.log").read()[-2000:])
PY
done
