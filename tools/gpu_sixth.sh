#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 --deselect tests/test_stock_gpu_compare.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
grep -h "vs reference golden\|vs oracle\|max-abs" gpurun_out/pytest_gpu.log | head -20
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench1.log 2>&1; echo "bench1 rc=$?"
python - <<PY
import json
for f in ("bench1",):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.log"%f).read().splitlines() if l.startswith("{")][-1])
        print(f, "fps", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "u8", round(d["e2e"]["uint8_frames"]["value"],1), d.get("breakdown_ms_per_step"), d.get("clocks"), "cpu", d.get("cpu_baseline",{}).get("value"), "launches", d["gpu_launches"])
    except Exception as e: print(f, "parse error", e); print(open("gpurun_out/%s.log"%f).read()[-3000:])
PY

