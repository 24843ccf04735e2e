#!/bin/bash
# One `ncu --set full` capture of the dominant kernel at the bench size (2^24 points, c = 20), after the same command has
# exited 0 without ncu; leaves the raw page as CSV (the .ncu-rep itself is scratch) -- profiles/r02_accumulate_raw.csv.
set -e
python scripts/profile_target.py 24 0 > gpurun_out/acc_r2_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_accumulate_s -c 1 -f -o gpurun_out/acc_r2 \
  python scripts/profile_target.py 24 0 > gpurun_out/acc_r2_ncu.log 2>&1
ncu -i gpurun_out/acc_r2.ncu-rep --page raw --csv > gpurun_out/acc_r2_raw.csv 2> gpurun_out/acc_r2_raw.err
ncu -i gpurun_out/acc_r2.ncu-rep --page details --csv > gpurun_out/acc_r2_details.csv 2>> gpurun_out/acc_r2_raw.err
