#!/bin/bash
# tools/gpu_submit.sh <timeout> [--gpus N] -- <command>: gpurun with retries while the pod answers "busy" (exit 3)
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 150
done
exit 3
