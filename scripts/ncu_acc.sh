#!/bin/bash
# usage: scripts/ncu_acc.sh <acc_mode> <logn> <tag>   -- one ncu --set full capture of the accumulate kernel
MODE=${1:-3}; LOGN=${2:-22}; TAG=${3:-acc}
export TB200_ACC_MODE=$MODE
python bench.py --steps 1 --warmup 1 --logn $LOGN --no-cpu-baseline --no-e2e --commit-nv 0 > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_accumulate_s -c 1 -f -o gpurun_out/${TAG} \
  python bench.py --steps 1 --warmup 0 --logn $LOGN --no-cpu-baseline --no-e2e --commit-nv 0 > gpurun_out/${TAG}_ncu.log 2>&1
