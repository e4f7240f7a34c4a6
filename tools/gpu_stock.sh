#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench1.log 2>&1; echo "bench1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/bench2.log 2>&1; echo "bench2 rc=$?"
python - <<PY
import json
for f in ("bench1","bench2"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.log"%f).read().splitlines() if l.startswith("{")][-1])
        print(f, "fps", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d.get("breakdown_ms_per_step"), d.get("clocks"), "cpu", d.get("cpu_baseline",{}).get("value"))
        for k,v in d.get("layers",{}).items(): print("   ", k, v)
    except Exception as e: print(f, "parse error", e); print(open("gpurun_out/%s.log"%f).read()[-3000:])
PY
