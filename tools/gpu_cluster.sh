#!/bin/bash
mkdir -p gpurun_out
for cl in 2 4; do
  LWB_CLUSTER=$cl timeout 600 python -m pytest -m gpu -q -p no:cacheprovider --timeout 600 tests/test_conv_gpu.py -k "not halo" > gpurun_out/conv_cl$cl.log 2>&1; echo "conv tests CL=$cl rc=$?"; tail -3 gpurun_out/conv_cl$cl.log
done
for cl in 1 2 4; do echo "=== LWB_CLUSTER=$cl"; LWB_CLUSTER=$cl timeout 300 python tools/conv_microbench.py 2>&1 | grep -E "halo=0" | grep -v "stats=0"; done
