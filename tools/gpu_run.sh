#!/bin/bash
# tools/gpu_run.sh -- the one parameterised GPU session script (run under gpurun from the repo root).
#   tools/gpu_run.sh <name> [stage ...]      stages are run in order; all output -> gpurun_out/<name>.log
# stages:
#   env                      nvidia-smi, host CPU flags
#   tests                    python -m pytest tests -m gpu -x -q
#   smoke                    __graft_entry__.smoke()
#   bench[:ARGS]             python bench.py ARGS                     (':' separates, '+' stands for a blank)
#   sb:ARGS                  simd-radix-sort_b200/sortbench ARGS
#   cub[:N]                  tools/cub_compare N   (comparison row only)
#   pytest:ARGS              python -m pytest ARGS -x -q
#   py:SECONDS:ARGS          timeout SECONDS python ARGS   (B200SORT_MGPU_DEBUG=1)
#   mbench:N[:ARGS]          torchrun --nproc-per-node N bench.py --gpus N ARGS  (with B200SORT_MGPU_TRACE=1)
#   launches:ARGS            ncu launch list of bench.py ARGS -> gpurun_out/<name>_launches.csv
#   ncu:KERNEL:SKIP:ARGS     ncu --set full of launch #SKIP of KERNEL in sortbench ARGS -> gpurun_out/<name>_KERNEL.ncu-rep
set -u
name=$1; shift
mkdir -p gpurun_out
log=gpurun_out/$name.log
: > "$log"
SB=simd-radix-sort_b200/sortbench
for st in "$@"; do
  kind=${st%%:*}; rest=""; [[ "$st" == *:* ]] && rest=${st#*:}; rest=${rest//+/ }
  echo "=== $st" >> "$log"
  case $kind in
    env) nvidia-smi >> "$log" 2>&1; grep -o -m1 'avx512_vbmi2' /proc/cpuinfo >> "$log"; nproc >> "$log" ;;
    tests) timeout 1500 python -m pytest tests -m gpu -x -q >> "$log" 2>&1 ;;
    smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> "$log" 2>&1 ;;
    bench) timeout 900 python bench.py $rest >> "$log" 2>&1 ;;
    sb) timeout 600 $SB $rest >> "$log" 2>&1 ;;
    sbq) timeout 25 $SB $rest >> "$log" 2>&1 ;;
    cub) timeout 600 tools/cub_compare $rest >> "$log" 2>&1 ;;
    pytest) timeout 1500 python -m pytest $rest -x -q >> "$log" 2>&1 ;;
    py) t=${rest%%:*}; args=${rest#*:}; timeout $t python $args >> "$log" 2>&1 ;;
    mbench) n=${rest%%:*}; args=""; [[ "$rest" == *:* ]] && args=${rest#*:}
            B200SORT_MGPU_TRACE=1 timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n $args >> "$log" 2>&1 ;;
    launches) timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${name}_launches.csv python bench.py $rest >> "$log" 2>&1 ;;
    ncu) k=${rest%%:*}; r2=${rest#*:}; skip=${r2%%:*}; args=${r2#*:}
         timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/${name}_$k $SB $args >> "$log" 2>&1 ;;
    *) echo "unknown stage $st" >> "$log" ;;
  esac
  echo "--- rc=$?" >> "$log"
done
tail -c 6000 "$log"
