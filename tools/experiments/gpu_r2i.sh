#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tdnet.py tests/test_gpu_graph.py tests/test_ablation.py -m gpu -q > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E " gpurun_out/pytest_i.log | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_i.json').read().strip().splitlines()[-1])
k = d['roofline']['kernel_ms_per_step']
print(d['ms_per_step'], d['e2e']['ms_per_step'], 'vbwd', k['vattn_bwd_D200_K7_M50000'], 'tailbwd', k['resnet_tail_bwd'])
PY
