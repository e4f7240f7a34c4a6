#!/bin/bash
# 8-GPU bench (one process per GPU, the driver's launch line); charged 8x -- keep it short
mkdir -p gpurun_out
N=${1:-8}
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N rc=$?"; tail -5 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open('gpurun_out/bench_n%s.json' % n).read().strip().splitlines()[-1])
print('fps', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'c3', d['config3_stream64'], 'bcast', d['init_broadcast'], d.get('streams'), d.get('clocks'))
PY
