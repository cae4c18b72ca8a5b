#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out/sweep4.log
SB=./simd-radix-sort_b200/sortbench
{
echo "== correctness direct modes"
for m in 1 2; do
timeout 120 $SB --n 5000003 --key u64 --pay 8 --iters 1 --opt algo=2 --opt scatter_mode=$m
timeout 120 $SB --n 5000003 --key u64 --pay 8,1,2 --iters 1 --desc --opt algo=1 --opt scatter_mode=$m
timeout 120 $SB --n 5000003 --key i64 --aos 16 --iters 1 --dist 1 --opt algo=2 --opt scatter_mode=$m
timeout 120 $SB --n 5000003 --key u32 --pay 4 --iters 1 --opt scatter_mode=$m
timeout 120 $SB --n 1000003 --key u16 --pay 1 --iters 1 --opt scatter_mode=$m
done
echo "== modes x cfg, u64+u64 n=2^28 hybrid"
for m in 0 1 2; do for cfg in 1 2 3 4; do echo "mode $m cfg $cfg"; timeout 300 $SB --n 268435456 --key u64 --pay 8 --iters 3 --noverify --opt algo=2 --opt tile_cfg=$cfg --opt scatter_mode=$m; done; done
echo "== 1e9 hybrid"
for m in 0 1 2; do for cfg in 1 3; do echo "mode $m cfg $cfg"; timeout 300 $SB --n 1000000000 --key u64 --pay 8 --iters 2 --noverify --opt algo=2 --opt tile_cfg=$cfg --opt scatter_mode=$m; done; done
echo "== u32+u32 2^28"
for m in 0 1 2; do echo "mode $m"; timeout 300 $SB --n 268435456 --key u32 --pay 4 --iters 3 --noverify --opt tile_cfg=1 --opt scatter_mode=$m; done
echo "== f32 + 3 payloads 5e8 desc"
for m in 0 2; do echo "mode $m"; timeout 300 $SB --n 500000000 --key f32 --pay 4,8,2 --iters 2 --desc --dist 4 --noverify --opt scatter_mode=$m; done
} > $OUT 2>&1
cat $OUT
