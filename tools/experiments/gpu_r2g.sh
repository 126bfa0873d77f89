#!/bin/bash
set -u
mkdir -p gpurun_out
for s in 0 1 2 3 8; do NSDP_KNN_SPLITS=$s timeout 120 python tools/microbench_knn.py 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_graph.py tests/test_gpu_vattn.py tests/test_gpu_mlp.py tests/test_gpu_index_kernels.py tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E " gpurun_out/pytest_g.log | tail -10
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_g.json').read().strip().splitlines()[-1])
k = d['roofline']['kernel_ms_per_step']
print(d['ms_per_step'], 'vbwd', k['vattn_bwd_D200_K7_M50000'], 'tailbwd', k['resnet_tail_bwd'])
PY
