#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vattn.py tests/test_gpu_tdnet.py -m gpu -q -x > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E " gpurun_out/pytest_k.log | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_k.json').read().strip().splitlines()[-1])
k = d['roofline']['kernel_ms_per_step']
print(d['ms_per_step'], d['e2e']['ms_per_step'], {n: v for n, v in k.items() if 'vattn' in n or 'tail' in n})
PY
NSDP_B200_LIB=nsdp_b200/lib/libnsdp_b200_trace.so timeout 300 python tools/trace_fwd.py > gpurun_out/trace_fwd.txt 2>&1
