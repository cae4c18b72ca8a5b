#!/bin/bash
set -u
mkdir -p gpurun_out
SB=./simd-radix-sort_b200/sortbench
CMD="$SB --n 67108864 --key u64 --pay 8 --iters 1 --noverify --opt use_match=0 --opt hist_match=0"
$CMD > gpurun_out/prof1_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:onesweep -s 8 -c 2 -o gpurun_out/prof1_sweep $CMD > gpurun_out/prof1_ncu.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hist_kernel -s 1 -c 1 -o gpurun_out/prof1_hist $CMD >> gpurun_out/prof1_ncu.log 2>&1
tail -3 gpurun_out/prof1_ncu.log
