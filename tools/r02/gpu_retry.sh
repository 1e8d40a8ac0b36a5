#!/bin/bash
# usage: tools/r02/gpu_retry.sh <call-log> <timeout> <command...>   -- retries while the pod answers "transient"
log=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient" $log; then exit 0; fi
  sleep 90
done
