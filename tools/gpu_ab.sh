#!/bin/bash
# A/B of env variants, alternating, on one box: usage gpu_ab.sh "VAR=a" "VAR=b" [rounds]
mkdir -p gpurun_out
A=$1; B=$2; R=${3:-3}
for i in $(seq 1 $R); do for v in "$A" "$B"; do
  timeout 300 env $v python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-extras --no-parity --steady-steps 300 > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python -c "
import json
d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); print('$v', 'ms', round(d['ms_per_step'],3), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'steady_ms', round(d['steady_state']['ms_per_step'],3), 'mhz', d['clocks']['sm_mhz'], d['steady_state']['clocks']['sm_mhz'])"
done; done
