#!/bin/bash
# compute-sanitizer memcheck over the small-shape kernel tests of the decoder (one-hot) and generic tensor-core kernels
set -u
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_vattn.py -x -q -k "(test_vattn_forward or test_vattn_backward) and auto and (shape_queryTrue or K16-D256 or M40-N300)" \
  > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|passed|failed|error" gpurun_out/sanitize_memcheck.log | head -20
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_vattn.py -x -q -k "test_vattn_backward and auto and shape_queryTrue and M333" \
  > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed|rror" gpurun_out/sanitize_racecheck.log | head -20
