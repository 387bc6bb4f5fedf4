#!/bin/bash
# usage: scripts/gpuN.sh <ngpus> <logname> <timeout-seconds> '<command>'
n=$1; shift
log=gpurun_out/$1.log; shift
to=$1; shift
for attempt in 1 2 3 4 5 6 7 8; do
  gpurun --gpus "$n" --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then break; fi
  sleep 120
done
echo "gpuN.sh done rc=$rc" >> "$log"
