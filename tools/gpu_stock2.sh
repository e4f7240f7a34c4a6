#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stock_gpu_compare.py -m gpu -q -s -p no:cacheprovider --timeout 600 > gpurun_out/stock.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/stock.log
