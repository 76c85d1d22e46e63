#!/bin/bash
# One GPU-box visit: parity tests, the bench line (default flags), the reference arm, ncu launch list + --set full capture.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag> [skip_tests]
TAG=${1:-r2}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
fi
timeout 600 python bench.py 2> gpurun_out/bench_$TAG.err > gpurun_out/bench_$TAG.json; tail -c 600 gpurun_out/bench_$TAG.json; echo
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/bench_ref_$TAG.err > gpurun_out/bench_ref_$TAG.json; tail -c 400 gpurun_out/bench_ref_$TAG.json; echo
bash scripts/gpu_profile.sh $TAG
