#!/bin/bash
# Targeted ncu captures inside the profiler range (one steady-state step); reports exported to CSV on the box.
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --profile-range"
cap() {  # name, kernel regex, skip, count
  timeout 900 $NCU --set full --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/$1 -f $BENCH > gpurun_out/$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
}
cap raster "k_face_raster" 0 1
cap norm "k_norm_act" 0 3
cap conv_a "k_conv_tc" 0 6
cap conv_b "k_conv_tc" 24 7
cap heads "k_heads7x7" 0 1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/launches_steady.csv $BENCH > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 600 python -m pytest tests/test_inpaintor.py -m gpu -q -p no:cacheprovider -s > gpurun_out/inpaintor.log 2>&1; echo "inpaintor rc=$?"; tail -5 gpurun_out/inpaintor.log
rm -f gpurun_out/conv_b.source.csv gpurun_out/conv_a.source.csv
du -sh gpurun_out
