#!/bin/bash
# like gpu_retry.sh with --gpus N.  Usage: gpu_retry_n.sh <N> <log> <timeout> <command...>
N=$1; shift; LOG=$1; shift; TO=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus $N --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if ! grep -q "status=transient" $LOG; then exit $rc; fi
  sleep 90
done
