#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"; tail -5 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('n2 fps', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'c3', d['config3_stream64'], 'bcast', d['init_broadcast'], d.get('streams'))
PY
