#!/bin/bash
# gpurun with retries while the pod answers "transient / busy" (nothing is charged for those).  Usage: gpurun_retry.sh <timeout> '<cmd>'
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun $GPURUN_FLAGS --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=3"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
