#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out/sweep2.log
SB=./simd-radix-sort_b200/sortbench
{
echo "== correctness"
for key in u8 i16 u32 f32 u64 f64; do timeout 120 $SB --n 1000003 --key $key --pay 4 --iters 1 --opt use_match=0 || echo "FAIL $key"; done
timeout 120 $SB --n 5000000 --key u64 --pay 8,1,2 --iters 1 --desc --opt use_match=0
timeout 120 $SB --n 5000000 --key i64 --aos 16 --iters 1 --dist 1 --opt use_match=0
echo "== sweep u64+u64 n=2^28 (ballot, plain hist)"
for cfg in 0 1 2 3; do
  timeout 300 $SB --n 268435456 --key u64 --pay 8 --iters 3 --noverify --opt tile_cfg=$cfg --opt use_match=0 --opt hist_match=0
done
echo "== u32+u32"
for cfg in 0 1 2 3; do timeout 300 $SB --n 268435456 --key u32 --pay 4 --iters 3 --noverify --opt tile_cfg=$cfg --opt use_match=0 --opt hist_match=0; done
echo "== 1e9"
timeout 300 $SB --n 1000000000 --key u64 --pay 8 --iters 2 --noverify --opt tile_cfg=2 --opt use_match=0 --opt hist_match=0
timeout 300 $SB --n 1000000000 --key u64 --pay 8 --iters 2 --noverify --opt tile_cfg=3 --opt use_match=0 --opt hist_match=0
} > $OUT 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest2.log 2>&1; echo "pytest exit $?" >> $OUT; tail -3 gpurun_out/pytest2.log >> $OUT
g++ -std=c++20 -O2 -Iinclude tests/cpp/dropin_test.cpp -o /tmp/dropin_test -L simd-radix-sort_b200 -lb200sort -Wl,-rpath,$PWD/simd-radix-sort_b200 && /tmp/dropin_test >> $OUT 2>&1
timeout 600 python bench.py --n 100000000 --steps 2 --warmup 1 --e2e-steps 1 --cpu-sample 4194304 >> $OUT 2>&1
cat $OUT
