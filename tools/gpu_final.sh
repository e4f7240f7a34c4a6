#!/bin/bash
# Round evidence: full GPU suite, smoke, bench (both arms), steady-state launch list, targeted ncu captures -> CSV.
# Output lands in gpurun_out/; summarise into profiles/rNN/ with tools/summarize_evidence.sh afterwards.
mkdir -p gpurun_out
git_head=$(cat .git_head 2>/dev/null || echo unknown)
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
LWB_RUN_STOCK=1 timeout 900 python -m pytest tests/test_stock_gpu_compare.py -m gpu -q -s -p no:cacheprovider --timeout 800 > gpurun_out/stock.log 2>&1; echo "stock rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2>gpurun_out/bench.err; echo "bench rc=$?"
NCU="ncu --clock-control none --profile-from-start off"
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-parity --steady-steps 0 --profile-range"
export LWB_GRAPH_PROFILE_NOTE=1   # ncu passes: eager launches on one stream (LWB_GRAPH=0 LWB_STREAMS=1), same kernels
timeout 600 env LWB_STREAMS=1 LWB_GRAPH=0 $NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/launches_steady.csv $BENCH > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
cap() {  # name, kernel regex, skip, count
  timeout 900 env LWB_STREAMS=1 LWB_GRAPH=0 $NCU --set full --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/$1 -f $BENCH > gpurun_out/$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
# DRAM traffic of every conv launch of one step (roofline.traffic): cheap metrics pass
timeout 600 env LWB_STREAMS=1 LWB_GRAPH=0 $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:k_conv_tc -c 64 --csv --log-file gpurun_out/conv_traffic.csv $BENCH > gpurun_out/ncu_traffic.log 2>&1; echo "traffic rc=$?"
# k_conv_tc2y (y-halo, default) launches of one step, in order: 0 stem, 1-3 encoders, 4-15 residual convs, 16 skipper @64,
# 17 merged T256->128, 18 skipper @128, 19 merged T128->64, 20 skipper @256, 21 folded heads (N = 32); the four phase launches
# of T512->256 are the only k_conv_tc2 (non-halo) launches
cap conv_res "^k_conv_tc2y$" 8 2
cap conv_stem "^k_conv_tc2y$" 0 1
cap conv_convt_merged "^k_conv_tc2y$" 19 1
cap conv_skip256 "^k_conv_tc2y$" 20 1
cap conv_heads "^k_conv_tc2y$" 21 1
cap norm "k_norm_act" 0 2
cap heads "k_heads" 0 1
cap raster "k_face_raster|k_resolve" 0 2
cuobjdump -sass impersonator_b200/liblwb_b200.so > /tmp/sass.txt 2>/dev/null
{ echo "SASS instruction counts of impersonator_b200/liblwb_b200.so (cuobjdump -sass | grep -c):";
  for pat in UTCHMMA UTCQMMA "\.2CTA" LDTM UTMALDG UTCBAR "SYNCS" ; do printf "%-10s %s\n" "$pat" "$(grep -c "$pat" /tmp/sass.txt)"; done; } > gpurun_out/sass_summary.txt
du -sh gpurun_out; python -c "
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]); print('fps',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['breakdown_ms_per_step'],d['roofline']['frac'],d['roofline']['issued_frac'],d['clocks'])
r=json.loads(open('gpurun_out/bench_ref.json').read().strip().splitlines()[-1]); print('ref fps',r['value'],r['cpu_baseline']['cores'])
"
