#!/bin/bash
set -u
mkdir -p gpurun_out
SB=./simd-radix-sort_b200/sortbench
CMD="$SB --n 268435456 --key u64 --pay 8 --iters 1 --noverify --dist 6 --opt margin_bits=1"
$SB --n 268435456 --key u64 --pay 8 --iters 2 --dist 6 --opt margin_bits=1 --prof > gpurun_out/prof4_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:onesweep -s 7 -c 1 -o gpurun_out/prof4_fix $CMD > gpurun_out/prof4_ncu.log 2>&1
gzip -9 gpurun_out/prof4_fix.ncu-rep
cat gpurun_out/prof4_plain.log; tail -3 gpurun_out/prof4_ncu.log; ls -la gpurun_out | tail -5
