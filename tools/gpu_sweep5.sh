#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out/sweep5.log
SB=./simd-radix-sort_b200/sortbench
{
echo "== correctness"
for ns in 1 2; do
timeout 120 $SB --n 5000003 --key u64 --pay 8 --iters 1 --opt algo=2 --opt nstage=$ns
timeout 120 $SB --n 5000003 --key u64 --pay 8,1,2 --iters 1 --desc --opt algo=1 --opt nstage=$ns
timeout 120 $SB --n 5000003 --key i64 --aos 16 --iters 1 --dist 1 --opt algo=2 --opt nstage=$ns
timeout 120 $SB --n 3000003 --key u32 --aos 64 --iters 1 --opt nstage=$ns
timeout 120 $SB --n 5000003 --key u32 --pay 4 --iters 1 --opt nstage=$ns
timeout 120 $SB --n 1000003 --key u16 --pay 1 --iters 1 --opt nstage=$ns
done
echo "== nstage x cfg, u64+u64 n=2^28 hybrid"
for ns in 1 2; do for cfg in 0 1 2 3 4 5; do echo "nstage $ns cfg $cfg"; timeout 300 $SB --n 268435456 --key u64 --pay 8 --iters 3 --noverify --opt algo=2 --opt tile_cfg=$cfg --opt nstage=$ns; done; done
echo "== 1e9 hybrid"
for ns in 1 2; do for cfg in 1 3 5; do echo "nstage $ns cfg $cfg"; timeout 300 $SB --n 1000000000 --key u64 --pay 8 --iters 2 --noverify --opt algo=2 --opt tile_cfg=$cfg --opt nstage=$ns; done; done
echo "== u32+u32 2^28"
for cfg in 1 3 5; do echo "cfg $cfg"; timeout 300 $SB --n 268435456 --key u32 --pay 4 --iters 3 --noverify --opt tile_cfg=$cfg; done
} > $OUT 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest5.log 2>&1; echo "pytest exit $?" >> $OUT; tail -5 gpurun_out/pytest5.log >> $OUT
timeout 900 python bench.py --n 200000000 --steps 2 --warmup 1 --e2e-steps 1 --cpu-sample 4194304 >> $OUT 2>&1
cat $OUT
