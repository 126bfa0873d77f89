#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_emlp.py tests/test_gpu_pointnet2_modules.py tests/test_gpu_tdnet.py -m gpu -q -x > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E " gpurun_out/pytest_c.log | tail -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_c.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['config']['step_execution'])
print({k: v for k, v in d['roofline']['kernel_ms_per_step'].items() if 'emlp' in k})
PY
tail -3 gpurun_out/bench_c.err
