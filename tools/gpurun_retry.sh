#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers "transient" (nothing charged)
T=$1; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out"
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  break
done
