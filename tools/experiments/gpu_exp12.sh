#!/bin/bash
set -u
echo "--- forced hi-only staging, small-shape backward tests"
NSDP_STAGE_LO=0 timeout 900 python -m pytest tests/test_gpu_vattn.py -x -q -k "backward" 2>&1 | tail -12
