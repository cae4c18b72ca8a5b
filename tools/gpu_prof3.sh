#!/bin/bash
set -u
mkdir -p gpurun_out
SB=./simd-radix-sort_b200/sortbench
CMD="$SB --n 67108864 --key u64 --pay 8 --iters 1 --noverify --opt algo=2 --opt tile_cfg=1"
$CMD > gpurun_out/prof3_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:onesweep -s 12 -c 1 -o gpurun_out/prof3_sweep $CMD > gpurun_out/prof3_ncu.log 2>&1
gzip -9 gpurun_out/prof3_sweep.ncu-rep
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:segfix -s 1 -c 1 -o gpurun_out/prof3_segfix $CMD >> gpurun_out/prof3_ncu.log 2>&1
gzip -9 gpurun_out/prof3_segfix.ncu-rep
cat gpurun_out/prof3_plain.log; tail -3 gpurun_out/prof3_ncu.log; ls -la gpurun_out
