#!/bin/bash
set -u
mkdir -p gpurun_out
SB=./simd-radix-sort_b200/sortbench
{
timeout 300 $SB --n 1000000000 --key u64 --pay 8 --iters 2 --noverify --prof
timeout 300 $SB --n 268435456 --key u32 --pay 4 --iters 2 --noverify --prof
timeout 300 $SB --n 1000000000 --key i64 --aos 16 --iters 2 --dist 1 --noverify --prof
timeout 5 nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 50 -i 0 | head -5
} > gpurun_out/quick.log 2>&1
cat gpurun_out/quick.log
