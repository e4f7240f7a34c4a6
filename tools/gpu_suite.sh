#!/bin/bash
# GPU suite + smoke + one bench line (development loop).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('fps', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), d.get('breakdown_ms_per_step'),
          'frac', d.get('roofline', {}).get('frac'), 'parity', d.get('parity', {}).get('max_abs'), 'steady', (d.get('steady_state') or {}).get('ms_per_step'))
    for k, v in sorted(d.get('layers', {}).items(), key=lambda kv: -kv[1]['ms']):
        print('  %-44s %.4f ms x%d  %s' % (k, v['ms'], v['n'], v.get('tflops_algorithmic', v.get('gbs'))))
except Exception as e:
    print('no bench line', e)
PY
