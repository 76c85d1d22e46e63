#!/bin/bash
# A/B two builds: compact C3 timing + sort_all timing + parity
for so in "$@"; do
  cp bevy_gaussian_splatting_b200/$so /tmp/cur.so; cp /tmp/cur.so bevy_gaussian_splatting_b200/libbgs.so
  echo "== $so"
  python - <<'PY' 2>&1 | grep -E "timing|parity"
import sys; sys.path.insert(0,'.')
from scripts.gpu_probe import timing, parity
parity(60_000, 333, 177, 0.3, f16=True, bits=16)
timing(6_000_000, 0.02, True, frames=30)
timing(6_000_000, 0.02, True, frames=20, sort_all=True)
PY
done
