#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out/sweep7.log
SB=./simd-radix-sort_b200/sortbench
{
echo "== correctness"
for key in u8 i16 u32 f32 u64 i64 f64; do timeout 120 $SB --n 1000003 --key $key --pay 4 --iters 1 || echo "FAIL $key"; done
timeout 120 $SB --n 5000003 --key u64 --pay 8 --iters 1 --opt algo=2
timeout 120 $SB --n 5000003 --key u64 --pay 8,1,2 --iters 1 --desc --opt algo=2
timeout 120 $SB --n 5000003 --key i64 --aos 16 --iters 1 --dist 1 --opt algo=2
timeout 120 $SB --n 5000003 --key i64 --aos 16 --iters 1 --dist 2 --opt algo=2
timeout 120 $SB --n 5000003 --key f64 --pay 8 --iters 1 --dist 4 --opt algo=2
timeout 120 $SB --n 5000003 --key u64 --pay 8 --iters 1 --dist 3
timeout 120 $SB --n 5000003 --key u64 --pay 8 --iters 1 --dist 5
echo "== 2^28 / 1e9 hybrid"
timeout 300 $SB --n 268435456 --key u64 --pay 8 --iters 3 --noverify
timeout 300 $SB --n 1000000000 --key u64 --pay 8 --iters 3 --noverify
echo "== other shapes"
timeout 300 $SB --n 268435456 --key u32 --pay 4 --iters 3 --noverify
timeout 300 $SB --n 1000000 --key u32 --pay 4 --iters 20 --noverify
timeout 300 $SB --n 500000000 --key f32 --pay 4,8,2 --iters 2 --desc --dist 4 --noverify
timeout 300 $SB --n 1000000000 --key i64 --aos 16 --iters 2 --dist 1 --noverify
timeout 300 $SB --n 1000000000 --key i64 --aos 16 --iters 2 --dist 2 --noverify
} > $OUT 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest7.log 2>&1; echo "pytest exit $?" >> $OUT; tail -5 gpurun_out/pytest7.log >> $OUT
timeout 900 python bench.py --n 200000000 --steps 2 --warmup 1 --e2e-steps 1 --cpu-sample 4194304 >> $OUT 2>&1
cat $OUT
