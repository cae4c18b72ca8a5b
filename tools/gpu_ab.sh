#!/bin/bash
SB=./simd-radix-sort_b200/sortbench
timeout 300 $SB --n 1000000000 --key u64 --pay 8 --iters 3 --prof
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
