#!/bin/bash
# GPU suite + smoke + bench variants (development loop; one gpurun call).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 --deselect tests/test_run_imitator_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 600 python -m pytest tests/test_run_imitator_gpu.py -m gpu -q -s -p no:cacheprovider --timeout 500 > gpurun_out/pytest_run_imitator.log 2>&1; echo "run_imitator rc=$?"; tail -6 gpurun_out/pytest_run_imitator.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
summ() {
python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'fps', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), d.get('breakdown_ms_per_step'),
          'frac', d.get('roofline', {}).get('frac'), 'parity', (d.get('parity') or {}).get('max_abs'), 'steady', (d.get('steady_state') or {}).get('ms_per_step'),
          'c3', (d.get('config3_stream64') or {}).get('ms_per_64_frames'), 'hmr', (d['e2e'].get('hmr') or {}).get('ms_per_frame'), 'pers', d.get('personalize'))
    for k, v in sorted(d.get('layers', {}).items(), key=lambda kv: -kv[1]['ms']):
        print('  %-52s %.4f ms x%d  %s' % (k, v['ms'], v['n'], v.get('tflops_algorithmic', v.get('gbs'))))
except Exception as e:
    print(f, 'no bench line', e)
PY
}
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; summ gpurun_out/bench.json
for v in $BENCH_VARIANTS; do
  timeout 400 env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --steady-steps 100 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?"; tail -2 gpurun_out/bench_$v.err; summ gpurun_out/bench_$v.json
done
