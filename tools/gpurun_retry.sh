#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'  -- retries while the pod answers busy/transient (nothing is charged for those)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out"
  if echo "$out" | grep -q "status=transient\|status=busy\|retry in a few minutes"; then sleep 45; continue; fi
  break
done
