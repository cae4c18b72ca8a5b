#!/bin/bash
SB=./simd-radix-sort_b200/sortbench
timeout 300 $SB --n 4400000000 --key u8 --iters 1 --prof
timeout 300 $SB --n 4300000000 --key u16 --pay 1 --iters 1 --prof --desc
timeout 300 $SB --n 2200000000 --key i64 --pay 4 --iters 1 --prof
