#!/bin/bash
# only the ncu passes of tools/gpu_final.sh (launch list, conv DRAM traffic, --set full captures)
mkdir -p gpurun_out
sed -n '/^NCU=/,/^cap raster/p' tools/gpu_final.sh > /tmp/ncu_part.sh
bash /tmp/ncu_part.sh
ls -la gpurun_out | tail -20
