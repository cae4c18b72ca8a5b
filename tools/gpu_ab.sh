#!/bin/bash
SB=./simd-radix-sort_b200/sortbench
timeout 300 $SB --n 1000000000 --key u64 --pay 8 --iters 2 --prof
timeout 300 $SB --n 1000000000 --key u64 --pay 8 --iters 2 --prof --opt junction_table=0
timeout 300 $SB --n 500000000 --key f64 --pay 8,4 --iters 1 --dist 2 --prof --desc
timeout 300 $SB --n 500000000 --key i64 --aos 32 --iters 1 --prof
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
