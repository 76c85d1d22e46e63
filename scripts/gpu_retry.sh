#!/bin/bash
# retry a gpurun call until a GPU slot is free (exit code 3 = nothing charged).  Usage: gpu_retry.sh <log> <timeout> <command...>
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if ! grep -q "status=transient" $LOG; then exit $rc; fi
  sleep 90
done
