#!/bin/bash
# A/B project-kernel builds: $@ = list of .so files under the package dir
for so in "$@"; do
  cp bevy_gaussian_splatting_b200/$so /tmp/cur.so
  cp /tmp/cur.so bevy_gaussian_splatting_b200/libbgs.so
  echo "== $so"
  python - <<'PY' 2>&1 | grep timing
import sys; sys.path.insert(0,'.')
from scripts.gpu_probe import timing
timing(6_000_000, 0.02, True, frames=40)
PY
done
