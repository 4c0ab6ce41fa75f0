#!/bin/bash
# usage: scripts/gpu/submit.sh <timeout_s> <log> <command...>   -- retries while the pod has no free slot (exit 3)
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
