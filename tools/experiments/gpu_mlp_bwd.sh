#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp.py -q -s -k "backward" 2>&1 | grep "rel\|W=\|passed\|failed\|Error" | tail -60 > gpurun_out/mlp_bwd_test.log
tail -50 gpurun_out/mlp_bwd_test.log | cut -c1-200
