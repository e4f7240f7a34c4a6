#!/bin/bash
# First GPU bring-up: kernel parity tests, each group in its own process (a trap in one group
# must not poison the others).  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 900 -s"
timeout 900 $PYT tests/test_raster_gpu.py > gpurun_out/raster.log 2>&1; echo "raster rc=$?"
timeout 600 $PYT tests/test_conv_gpu.py -k "heads or norm_act or warp or direct" > gpurun_out/glue.log 2>&1; echo "glue rc=$?"
timeout 900 $PYT tests/test_conv_gpu.py -k "conv2d or transpose or concat" > gpurun_out/conv.log 2>&1; echo "conv rc=$?"
timeout 600 $PYT tests/test_conv_gpu.py -k "stem" > gpurun_out/stem.log 2>&1; echo "stem rc=$?"
for f in raster glue conv stem; do echo "=== $f"; grep -E "passed|failed|error|PASS|FAIL|Error|mismatch|max-abs|stats rel|timeout" gpurun_out/$f.log | tail -60; done
