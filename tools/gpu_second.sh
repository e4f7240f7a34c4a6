#!/bin/bash
# Second GPU run: end-to-end generator parity, smoke(), bench (ours + reference arm), launch list, ncu capture.
mkdir -p gpurun_out
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 900 -s"
timeout 900 $PYT tests/test_generator_gpu.py > gpurun_out/gen.log 2>&1; echo "gen rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench_ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 30 -c 3 -o gpurun_out/prof_conv -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
for f in gen smoke bench bench_ref; do echo "=== $f"; tail -25 gpurun_out/$f.log; done
