#!/bin/bash
# quick perf iteration: C3 timing + launch list (no tests)
TAG=${1:-q}
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/quick_$TAG.log
import sys; sys.path.insert(0,'.')
from scripts.gpu_probe import timing, parity
parity(100_000, 640, 360, 0.1)
parity(60_000, 333, 177, 1.0, f16=True)
timing(6_000_000, 0.02, True, frames=40)
timing(1_000_000, 1.0, False, frames=8)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 36 -c 12 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python scripts/launch_list.py gpurun_out/launches_$TAG.csv
