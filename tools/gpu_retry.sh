#!/bin/bash
# usage: tools/gpu_retry.sh <log> <timeout> <command...>   -- retries while the pod answers "busy" (exit 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient" $log; then sleep 90; continue; fi
  exit $rc
done
