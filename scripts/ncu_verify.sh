#!/bin/bash
# Launch list of the verifier's device steps and one `ncu --set full` capture of the cooperative GT power kernel, after the
# same command has exited 0 without ncu; leaves CSV pages (the .ncu-rep itself is scratch).
set -e
python scripts/profile_verify_target.py > gpurun_out/ver_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ver_launches.csv \
  python scripts/profile_verify_target.py > gpurun_out/ver_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fq12_pow_coop -c 1 -f -o gpurun_out/ver_pow \
  python scripts/profile_verify_target.py > gpurun_out/ver_pow_ncu.log 2>&1
ncu -i gpurun_out/ver_pow.ncu-rep --page raw --csv > gpurun_out/ver_pow_raw.csv 2> gpurun_out/ver_pow_raw.err
