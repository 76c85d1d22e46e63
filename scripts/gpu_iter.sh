#!/bin/bash
# one quick GPU iteration: parity tests (fail fast), sort/bin timelines, bench line.  Usage: bash scripts/gpu_iter.sh <tag> [pytest -k expr]
TAG=${1:-it}
mkdir -p gpurun_out
if [ -n "$2" ]; then
  timeout 400 python -m pytest tests -x -q -m gpu -k "$2" > gpurun_out/pytest_gpu_$TAG.log 2>&1
else
  timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1
fi
tail -6 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python scripts/timeline_sort.py > gpurun_out/tl_sort_$TAG.log 2>&1
timeout 300 python scripts/timeline.py > gpurun_out/tl_bin_$TAG.log 2>&1
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$TAG.json").read())
    print({k: d[k] for k in ["value", "ms_per_step", "frame_ms_p50"]}, d["e2e"]["value"])
    print([(s["stage"], s["us"]) for s in d["stages"]])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_$TAG.err").read()[-2000:])
PY
