#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   -- retries while gpurun answers 3 (no box / slot free, nothing charged)
log=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
