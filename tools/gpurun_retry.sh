#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing charged): tools/gpurun_retry.sh <timeout> '<command>'
T=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  echo "$out" | tail -14; exit 0
done
echo "$out" | tail -5; exit 3
