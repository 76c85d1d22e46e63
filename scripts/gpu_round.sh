#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list and one --set full capture.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag> [skip_tests]
TAG=${1:-r1}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
fi
timeout 900 python bench.py 2> gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json
# launch list: 3 warm-up frames (12 launches each) skipped, then 2 frames; cold-cache serialised times
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 36 -c 24 --csv \
   --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
# full capture of one frame's kernels
timeout 1200 ncu --set full --clock-control none --import-source on -s 36 -c 12 -f -o gpurun_out/prof_$TAG \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out
