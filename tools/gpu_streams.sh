#!/bin/bash
mkdir -p gpurun_out
for v in "LWB_STREAMS=2" "LWB_STREAMS=4" "LWB_STREAMS=1"; do
  timeout 300 env $v python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras --steady-steps 200 > gpurun_out/s_$v.json 2> gpurun_out/s_$v.err; echo "$v rc=$?"; tail -2 gpurun_out/s_$v.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/s_$v.json').read().strip().splitlines()[-1]); print('$v', 'fps', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'steady', d['steady_state']['ms_per_step'], d['clocks'])"
done
