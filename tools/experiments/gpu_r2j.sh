#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vattn.py tests/test_gpu_tdnet.py -m gpu -q -x > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E " gpurun_out/pytest_j.log | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_j.json').read().strip().splitlines()[-1])
k = d['roofline']['kernel_ms_per_step']
print(d['ms_per_step'], d['e2e']['ms_per_step'], 'vbwd', k['vattn_bwd_D200_K7_M50000'], 'tailbwd', k['resnet_tail_bwd'], d['config']['step_execution'][:30])
PY
tail -3 gpurun_out/bench_j.err
