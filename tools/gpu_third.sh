#!/bin/bash
# GPU run 3: full -m gpu suite, bench with per-layer breakdown, steady-state launch list, ncu captures of the top kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
NCU="ncu --clock-control none --profile-from-start off"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_steady.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --profile-range > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 1500 $NCU --set full --import-source on -k regex:"k_conv_tc|k_norm_act|k_face_raster|k_heads7x7|k_resolve" -c 90 -o gpurun_out/prof_step -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --profile-range > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -c 6000 gpurun_out/bench.log
