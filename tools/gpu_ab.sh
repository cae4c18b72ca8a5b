#!/bin/bash
SB=./simd-radix-sort_b200/sortbench
timeout 300 $SB --n 1000000000 --key u64 --pay 8 --iters 3 --prof
timeout 300 $SB --n 268435456 --key u32 --pay 4 --iters 3 --prof
