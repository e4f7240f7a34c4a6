#!/bin/bash
# gpurun_out/ (scratch, written by tools/gpu_final.sh on the GPU box) -> profiles/$1/ (tracked).  usage: summarize_evidence.sh r02
set -e
R=${1:-r02}; D=profiles/$R; G=gpurun_out
mkdir -p $D
last() { tail -n 1 "$1"; }
[ -s $G/bench.json ] && last $G/bench.json > $D/bench_n1.json
[ -s $G/bench_ref.json ] && last $G/bench_ref.json > $D/bench_reference_arm.json
for f in stock_compare.json sass_summary.txt smoke.log pytest_gpu.log launches_steady.csv conv_traffic.csv; do [ -s $G/$f ] && cp $G/$f $D/$f; done
[ -s $G/launches_steady.csv ] && python tools/launch_shares.py $G/launches_steady.csv 1 > $D/launch_shares.md
[ -s $G/conv_traffic.csv ] && python tools/traffic_from_csv.py $G/conv_traffic.csv profiles/conv_traffic.json
for n in conv_res conv_stem conv_convt_merged conv_skip256 conv_heads norm heads raster; do
  [ -s $G/$n.raw.csv ] && python tools/ncu_summary.py $G/$n.raw.csv > $D/ncu_$n.txt
done
ls -la $D
