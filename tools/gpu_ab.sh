#!/bin/bash
timeout 60 ./simd-radix-sort_b200/sortbench --n 1000000000 --key u64 --pay 8 --iters 2 --prof
timeout 60 ./simd-radix-sort_b200/sortbench --n 268435456 --key u32 --pay 4 --iters 2 --prof | tail -1
timeout 100 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
