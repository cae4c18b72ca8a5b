#!/bin/bash
# round-end evidence: parity tests, smoke, the bench line, the reference arm, the ncu launch list of the
# bench command and one full capture of the dominant kernel
set -u
mkdir -p gpurun_out
OUT=gpurun_out/final.log
{
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (b200)"; timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_b200.json
echo "== bench (reference arm)"; timeout 1500 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
} > $OUT 2>&1
# launch list of the same bench command (metric-only pass; numbers printed under ncu are not bench values)
BCMD="python bench.py --gpus 1 --steps 2 --warmup 3 --e2e-steps 0 --cpu-sample 0"
$BCMD > gpurun_out/bench_plain_for_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BCMD > gpurun_out/ncu_launches.log 2>&1
SB=./simd-radix-sort_b200/sortbench
CMD="$SB --n 67108864 --key u64 --pay 8 --iters 1 --noverify"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:onesweep -s 4 -c 1 -o gpurun_out/final_sweep $CMD > gpurun_out/final_ncu.log 2>&1
gzip -9 gpurun_out/final_sweep.ncu-rep
# the FIX instantiation (last pass) at the run density of the 1e9 headline: 2^28 62-bit keys, cut at bit 32
CMD2="$SB --n 268435456 --key u64 --pay 8 --iters 1 --noverify --dist 6 --opt margin_bits=1 --opt allow_lshift=0"
$CMD2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:onesweep -s 7 -c 1 -o gpurun_out/final_fix $CMD2 >> gpurun_out/final_ncu.log 2>&1
gzip -9 gpurun_out/final_fix.ncu-rep
$SB --n 1000000000 --key u64 --pay 8 --iters 3 --prof > gpurun_out/final_sortbench.log 2>&1
$SB --n 268435456 --key u32 --pay 4 --iters 3 --prof >> gpurun_out/final_sortbench.log 2>&1
$SB --n 500000000 --key f32 --pay 4,8,2 --iters 2 --prof --desc --dist 4 >> gpurun_out/final_sortbench.log 2>&1
$SB --n 2000000000 --key i64 --aos 16 --iters 1 --prof --dist 1 >> gpurun_out/final_sortbench.log 2>&1
$SB --n 2000000000 --key i64 --aos 16 --iters 1 --prof --dist 2 >> gpurun_out/final_sortbench.log 2>&1
$SB --n 1000000 --key u32 --pay 4 --iters 10 >> gpurun_out/final_sortbench.log 2>&1
cat $OUT; ls -la gpurun_out
