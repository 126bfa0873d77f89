#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_one_step.csv python tools/one_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
