#!/bin/bash
# usage: [GPUS=n] tools/r02/gpu_retry.sh <call-log> <timeout> <command...>   -- retries while the pod answers "transient"
log=$1; shift; to=$1; shift
extra=""
if [ -n "$GPUS" ]; then extra="--gpus $GPUS"; fi
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun $extra --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient\|rc=3\|no box" $log; then exit 0; fi
  sleep 90
done
