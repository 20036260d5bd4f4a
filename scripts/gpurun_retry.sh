#!/bin/bash
# usage: gpurun_retry.sh <log> <timeout_s> <command...>  -- retries while the pod answers busy (rc 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc" >> $log; exit $rc; fi
  sleep 120
done
