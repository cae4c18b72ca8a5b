#!/bin/bash
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 100 ./simd-radix-sort_b200/sortbench --n 1000000000 --key u64 --pay 8 --iters 2 --prof
