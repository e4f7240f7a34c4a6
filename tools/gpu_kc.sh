#!/bin/bash
mkdir -p gpurun_out
LWB_KC=32 timeout 600 python -m pytest -m gpu -q -p no:cacheprovider --timeout 600 tests/test_conv_gpu.py -k "not halo" > gpurun_out/conv_kc32.log 2>&1; echo "conv tests KC=32 rc=$?"; tail -4 gpurun_out/conv_kc32.log
for kc in 64 32; do echo "=== LWB_KC=$kc"; LWB_KC=$kc timeout 300 python tools/conv_microbench.py 2>&1 | grep -E "halo=0"; done
