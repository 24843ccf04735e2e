#!/bin/bash
# accumulate-kernel variants side by side (TB200_ACC_MODE): step and accumulate-stage time at 2^LOGN, bit-exactness checked
LOGN=${1:-24}; shift
for m in "$@"; do
  TB200_ACC_MODE=$m python bench.py --steps 3 --warmup 2 --logn $LOGN --no-cpu-baseline --no-e2e --commit-nv 0 > gpurun_out/bench_mode$m.json 2> gpurun_out/bench_mode$m.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_mode$m.json'));print('mode',$m, round(d['ms_per_step'],2), round(d['stages_ms']['accumulate'],2), d['verified_bit_exact'])"
done
