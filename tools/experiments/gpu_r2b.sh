#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tdnet.py tests/test_gpu_pointnet2_modules.py tests/test_reference_scripts.py -m gpu -q -s > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|worst|d/dq|stage 2|graph" gpurun_out/pytest_b.log | tail -20
for g in 1 0; do
NSDP_B200_GRAPH=$g timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_graph$g.json 2> gpurun_out/bench_graph$g.err; echo "bench graph=$g rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_graph$g.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['config']['step_execution'])
PY
grep -i "graph" gpurun_out/bench_graph$g.err | tail -3
done
