#!/bin/bash
# only the ncu passes of tools/gpu_final.sh (launch list, conv DRAM traffic, --set full captures)
mkdir -p gpurun_out
sed -n '/^NCU=/,/^cap raster/p' tools/gpu_final.sh > /tmp/ncu_part.sh
bash /tmp/ncu_part.sh
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 --durations=12 > gpurun_out/pytest_durations.log 2>&1; tail -20 gpurun_out/pytest_durations.log
ls -la gpurun_out | tail -20
