#!/bin/bash
# ncu launch list (2 frames) + one --set full capture of one frame (C3).  Usage: bash scripts/gpu_profile.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
# upload: repack + cutoff table = 2 launches; 5 warm frames x 6 = 30 -> skip 32
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 32 -c 12 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/ncu_frame.py 8 > gpurun_out/ncu_list_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -s 32 -c 6 -f -o gpurun_out/prof_$TAG python scripts/ncu_frame.py 7 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep gpurun_out/launches_$TAG.csv
tail -3 gpurun_out/ncu_list_$TAG.log
