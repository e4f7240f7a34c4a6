#!/bin/bash
mkdir -p gpurun_out
LWB_PRECISION=fp16f8 timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider --timeout 900 --deselect tests/test_stock_gpu_compare.py > gpurun_out/f8_suite.log 2>&1; echo "pytest rc=$?"
grep -h "vs reference golden\|vs oracle\|max-abs %\|oracle loop\|passed\|failed" gpurun_out/f8_suite.log | grep -v "fp16f8 vs\|/split\|/fast\|/halo" | head -40
