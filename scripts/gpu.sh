#!/bin/bash
# usage: scripts/gpu.sh <logname> <timeout-seconds> '<command>'   -- gpurun with retries while the pod has no free slot (exit code 3)
log=gpurun_out/$1.log; shift
to=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then break; fi
  sleep 90
done
echo "gpu.sh done rc=$rc" >> "$log"
