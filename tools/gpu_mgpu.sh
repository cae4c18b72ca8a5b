#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out/mgpu.log
nvidia-smi -L > $OUT 2>&1
NG=$(nvidia-smi -L | wc -l)
REC=${REC:-1000000000}
B200SORT_MGPU_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 2 --warmup 1 --records $REC --e2e-steps 1 >> $OUT 2>&1; echo "bench exit $?" >> $OUT
grep -v "^W\|^\*\*\*" $OUT | tail -30
