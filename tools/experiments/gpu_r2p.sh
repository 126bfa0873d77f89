#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_adam.py tests/test_gpu_graph.py tests/test_gpu_tdnet.py -m gpu -q -x > gpurun_out/pytest_p.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E " gpurun_out/pytest_p.log | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_p.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['step_execution'][:40], d['gpu_launches'])
PY
tail -3 gpurun_out/bench_p.err
