#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out/hyb1.log
SB=./simd-radix-sort_b200/sortbench
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest3.log 2>&1; echo "pytest exit $?" > $OUT; tail -25 gpurun_out/pytest3.log >> $OUT
{
echo "== hybrid vs lsd, u64+u64"
for n in 16777216 268435456 1000000000; do
  timeout 300 $SB --n $n --key u64 --pay 8 --iters 3 --opt algo=2 --opt tile_cfg=1
  timeout 300 $SB --n $n --key u64 --pay 8 --iters 3 --opt algo=1 --opt tile_cfg=1
done
echo "== hybrid other dists (1e9 aos i64, f64)"
timeout 300 $SB --n 1000000000 --key i64 --aos 16 --iters 2 --dist 1 --opt algo=2
timeout 300 $SB --n 1000000000 --key i64 --aos 16 --iters 2 --dist 2 --opt algo=2
timeout 300 $SB --n 500000000 --key f64 --pay 8 --iters 2 --dist 4 --opt algo=2
timeout 300 $SB --n 500000000 --key f64 --pay 8 --iters 2 --dist 4 --opt algo=1
timeout 300 $SB --n 500000000 --key u64 --pay 8 --iters 2 --dist 3 --opt algo=2
echo "== margin sweep 1e9"
for m in 1 2 4 6; do timeout 300 $SB --n 1000000000 --key u64 --pay 8 --iters 2 --noverify --opt algo=2 --opt margin_bits=$m; done
echo "== hist alone (u32 key only n=1e9 -> 4 passes)"; timeout 300 $SB --n 1000000000 --key u32 --pay 4 --iters 2 --noverify
} >> $OUT 2>&1
timeout 900 python bench.py --n 200000000 --steps 2 --warmup 1 --e2e-steps 1 --cpu-sample 4194304 >> $OUT 2>&1
cat $OUT
