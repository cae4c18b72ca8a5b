#!/bin/bash
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_mgpu.py -x -q -m gpu 2>&1 | tail -25
REC=${REC:-1000000000}
B200SORT_MGPU_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 3 --warmup 3 --records $REC --e2e-steps 0 > gpurun_out/mgpu$NG.log 2>&1; echo "bench exit $?" >> gpurun_out/mgpu$NG.log
grep -v "^W\|^\*\*\*" gpurun_out/mgpu$NG.log | grep "rank 0\]\|metric\|exit\|rror" | tail -8
