#!/bin/bash
for m in 8 7 6; do
  touch bevy_gaussian_splatting_b200/csrc/raster.cu
  make -C bevy_gaussian_splatting_b200/csrc -j8 EXTRA=-DRT_MIN_CTAS=$m > /dev/null 2>&1
  echo "== RT_MIN_CTAS=$m"
  timeout 240 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'p50', d['frame_ms_p50'], 'e2e', d['e2e']['value'], [s['us'] for s in d['stages']])"
done
