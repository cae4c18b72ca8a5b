#!/bin/bash
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
B200SORT_MGPU_TRACE=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/mgpu8.log 2>&1; echo "bench exit $?" >> gpurun_out/mgpu8.log
grep -v "^W\|^\*\*\*" gpurun_out/mgpu8.log | grep "rank 0\]\|rank 7\]\|metric\|exit\|Error\|error" | tail -16
