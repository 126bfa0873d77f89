#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'  — retries while the pod answers "busy" (rc 3 / transient), up to ~40 min
T=$1; shift
for i in $(seq 1 16); do
  out=$(gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|no box\|busy"; then sleep 150; continue; fi
  break
done
