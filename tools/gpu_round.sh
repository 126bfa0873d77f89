#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the per-launch list of one step and full ncu captures of the
# decoder kernels. Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches_one_step.csv python tools/one_step.py > gpurun_out/ncu_launches.log 2>&1
  echo "launch list rc=$?"
  REPS=2 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'vattn_bwd_tc_kernel|vattn_fwd_tc_kernel|dw_tc_kernel|resnet_tail' -f -o gpurun_out/prof_decoder \
      python tools/run_decoder_bwd.py > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
fi
