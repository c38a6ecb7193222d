#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> <command...>  -- retries while the pod answers "transient" (no slot free)
log=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient" $log; then break; fi
  sleep 90
done
