#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out/sweep9.log
SB=./simd-radix-sort_b200/sortbench
{
echo "== correctness"
for key in u8 u32 f32 u64 f64; do timeout 120 $SB --n 1000003 --key $key --pay 4 --iters 1 || echo "FAIL $key"; done
timeout 120 $SB --n 5000003 --key u64 --pay 8,1,2 --iters 1 --desc --opt algo=2
timeout 120 $SB --n 5000003 --key i64 --aos 16 --iters 1 --dist 2 --opt algo=2
timeout 120 $SB --n 777 --key u64 --pay 8 --iters 1
echo "== timing"
timeout 300 $SB --n 1000000000 --key u64 --pay 8 --iters 3 --noverify --prof
for cfg in 0 2 3; do timeout 300 $SB --n 1000000000 --key u64 --pay 8 --iters 2 --noverify --opt tile_cfg=$cfg; done
timeout 300 $SB --n 268435456 --key u32 --pay 4 --iters 3 --noverify --prof
timeout 300 $SB --n 1000000 --key u32 --pay 4 --iters 20 --noverify --prof
} > $OUT 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest9.log 2>&1; echo "pytest exit $?" >> $OUT; tail -5 gpurun_out/pytest9.log >> $OUT
cat $OUT
CMD="$SB --n 67108864 --key u64 --pay 8 --iters 1 --noverify"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:segfix -s 1 -c 1 -o gpurun_out/prof4_segfix $CMD > gpurun_out/prof4_ncu.log 2>&1
gzip -9 gpurun_out/prof4_segfix.ncu-rep
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:onesweep -s 12 -c 1 -o gpurun_out/prof4_sweep $CMD >> gpurun_out/prof4_ncu.log 2>&1
gzip -9 gpurun_out/prof4_sweep.ncu-rep
ls -la gpurun_out
