#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'  — retries while the pod answers "busy", keeps the full log in gpurun_out/last_stdout.txt
T=$1; shift
mkdir -p gpurun_out
for i in $(seq 1 16); do
  gpurun --timeout "$T" -- "$@" > gpurun_out/last_stdout.txt 2>&1
  tail -45 gpurun_out/last_stdout.txt
  if grep -q "status=transient\|no box\|busy" gpurun_out/last_stdout.txt; then sleep 150; continue; fi
  break
done
