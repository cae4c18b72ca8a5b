#!/bin/bash
# first GPU contact: environment facts, native self-checks, parity tests, variant sweep
set -u
mkdir -p gpurun_out
OUT=gpurun_out/first.log
{
echo "== env"; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv
grep -m1 "model name" /proc/cpuinfo; nproc; grep -c avx512_vbmi2 /proc/cpuinfo
SB=./simd-radix-sort_b200/sortbench
echo "== native self-checks (small)"
for key in u8 i16 u32 f32 u64 i64 f64; do
  timeout 120 $SB --n 1000003 --key $key --pay 4 --iters 1 || echo "FAIL $key"
done
timeout 120 $SB --n 5000000 --key u64 --pay 8,1,2 --iters 1 --desc
timeout 120 $SB --n 5000000 --key i64 --aos 16 --iters 1 --dist 1
timeout 120 $SB --n 3000000 --key f32 --pay 4,8,2 --iters 1 --desc --dist 4
timeout 120 $SB --n 3000000 --key u32 --aos 64 --iters 1
} > $OUT 2>&1
echo "== pytest" >> $OUT
timeout 1500 python -m pytest tests -x -q -m gpu >> gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> $OUT
tail -15 gpurun_out/pytest.log >> $OUT
{
SB=./simd-radix-sort_b200/sortbench
echo "== sweep u64+u64 n=2^28"
for cfg in 0 1 2 3; do for m in 1 0; do
  timeout 300 $SB --n 268435456 --key u64 --pay 8 --iters 3 --noverify --opt tile_cfg=$cfg --opt use_match=$m
done; done
echo "== hist variants"
timeout 300 $SB --n 268435456 --key u64 --pay 8 --iters 3 --noverify --opt hist_match=0
echo "== u32+u32"
for cfg in 0 1 2 3; do timeout 300 $SB --n 268435456 --key u32 --pay 4 --iters 3 --noverify --opt tile_cfg=$cfg; done
timeout 300 $SB --n 1000000 --key u32 --pay 4 --iters 10
echo "== 1e9 u64+u64"
timeout 600 $SB --n 1000000000 --key u64 --pay 8 --iters 2
echo "== c3 / c4 shapes"
timeout 600 $SB --n 500000000 --key f32 --pay 4,8,2 --iters 2 --desc --dist 4
timeout 600 $SB --n 1000000000 --key i64 --aos 16 --iters 2 --dist 1
timeout 600 $SB --n 1000000000 --key i64 --aos 16 --iters 2 --dist 2
} >> $OUT 2>&1
echo done >> $OUT
