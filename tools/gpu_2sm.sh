#!/bin/bash
mkdir -p gpurun_out
LWB_2SM=1 timeout 600 python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 tests/test_conv_gpu.py -k "not halo and not stem" -x > gpurun_out/conv_2sm.log 2>&1; echo "conv tests 2SM rc=$?"; tail -15 gpurun_out/conv_2sm.log | cut -c1-200
for m in 0 1; do echo "=== LWB_2SM=$m"; LWB_2SM=$m timeout 300 python tools/epi_probe.py 2>&1 | tail -5; done
