#!/bin/bash
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
REC=${REC:-1000000000}
run() {
  echo "=== $*"
  env "$@" B200SORT_MGPU_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 1 --warmup 1 --records $REC --e2e-steps 0 2>&1 | grep "rank 0\]" | tail -1
}
run NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
run NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=16
run NCCL_P2P_USE_CUDA_MEMCPY=1
run NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 NCCL_BUFFSIZE=16777216
