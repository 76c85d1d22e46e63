#!/bin/bash
# binning rounds vs one round across footprint sizes (tuning probe for the auto threshold and the schedule)
for SC in 1.0 0.5 0.3 0.2 0.1; do
  echo "== 6M f16 scale $SC"; timeout 120 python scripts/raw6m.py 6000000 1 $SC 2>&1 | grep "frame 3\|identical"
done
for SC in 1.0 0.3; do
  echo "== 1M f32 scale $SC"; timeout 120 python scripts/raw6m.py 1000000 0 $SC 2>&1 | grep "frame 3\|identical"
done
for F in "8,64,512,4096" "16,256,4096" "32,256,2048,16384"; do
  echo "== BGS_CHUNK_FRACS=$F"
  BGS_CHUNK_FRACS=$F timeout 120 python scripts/raw6m.py 6000000 1 2>&1 | grep "rounds frame 3"
  BGS_CHUNK_FRACS=$F timeout 120 python scripts/raw6m.py 6000000 1 0.3 2>&1 | grep "rounds frame 3"
  BGS_CHUNK_FRACS=$F timeout 120 python scripts/raw6m.py 1000000 0 2>&1 | grep "rounds frame 3"
done
