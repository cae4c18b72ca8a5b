#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out/sweep10.log
SB=./simd-radix-sort_b200/sortbench
{
echo "== correctness"
for key in u8 u32 f32 u64 f64; do timeout 120 $SB --n 1000003 --key $key --pay 4 --iters 1 || echo "FAIL $key"; done
timeout 120 $SB --n 20000003 --key u64 --pay 8,1,2 --iters 1 --desc
timeout 120 $SB --n 20000003 --key i64 --aos 16 --iters 1 --dist 2
timeout 120 $SB --n 20000003 --key i64 --aos 16 --iters 1 --dist 1
timeout 120 $SB --n 20000003 --key u64 --pay 8 --iters 1 --dist 3
echo "== timing"
timeout 300 $SB --n 1000000000 --key u64 --pay 8 --iters 3 --noverify --prof
timeout 300 $SB --n 268435456 --key u32 --pay 4 --iters 3 --noverify --prof
timeout 300 $SB --n 500000000 --key f32 --pay 4,8,2 --iters 2 --desc --dist 4 --noverify --prof
timeout 300 $SB --n 2000000000 --key i64 --aos 16 --iters 1 --dist 2 --noverify --prof
timeout 300 $SB --n 2000000000 --key i64 --aos 16 --iters 1 --dist 1 --noverify --prof
} > $OUT 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest10.log 2>&1; echo "pytest exit $?" >> $OUT; tail -5 gpurun_out/pytest10.log >> $OUT
cat $OUT
