#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -q -x -p no:cacheprovider --timeout 600 > gpurun_out/pytest_yhalo_conv.log 2>&1; echo "conv tests (yhalo on) rc=$?"; tail -15 gpurun_out/pytest_yhalo_conv.log
timeout 900 python -m pytest tests/test_generator_gpu.py tests/test_hmr_gpu.py tests/test_inpaintor.py -m gpu -q -x -p no:cacheprovider --timeout 600 > gpurun_out/pytest_yhalo_gen.log 2>&1; echo "generator/hmr/inpaintor tests (yhalo on) rc=$?"; tail -8 gpurun_out/pytest_yhalo_gen.log
for v in "LWB_YHALO=1" "LWB_YHALO_NA=2" "LWB_YHALO=0"; do
  timeout 400 env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --steady-steps 100 > gpurun_out/y_$v.json 2> gpurun_out/y_$v.err; echo "bench $v rc=$?"; tail -2 gpurun_out/y_$v.err
  python - "gpurun_out/y_$v.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'fps', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), d.get('breakdown_ms_per_step'), 'frac', d['roofline']['frac'], 'parity', d['parity']['max_abs'], 'steady', d['steady_state']['ms_per_step'])
    for k, v in sorted(d.get('layers', {}).items(), key=lambda kv: -kv[1]['ms']):
        if not k.startswith('norm'): print('  %-52s %.4f ms x%d  %s' % (k, v['ms'], v['n'], v.get('tflops_algorithmic')))
except Exception as e:
    print('no line', e)
PY
done
