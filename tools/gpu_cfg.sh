#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --size 512 --batch 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_512_b8.json 2> gpurun_out/bench_512.err; echo "512 rc=$?"
timeout 400 python bench.py --size 256 --batch 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_256_b8.json 2> gpurun_out/bench_256b8.err; echo "b8 rc=$?"
python - <<PY
import json
for f in ("bench_512_b8","bench_256_b8"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "fps", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["breakdown_ms_per_step"], d["roofline"]["frac"])
    except Exception as e: print(f, "err", e)
PY
tail -3 gpurun_out/bench_512.err
