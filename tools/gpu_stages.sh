#!/bin/bash
for st in 0 2 3 4; do echo "=== LWB_STAGES=$st (0 = max)"; LWB_CLUSTER=1 LWB_STAGES=$st timeout 300 python tools/conv_microbench.py 2>&1 | grep -E "res 512|skipper 64"; done
