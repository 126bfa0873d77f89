#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/microbench_c4.py > gpurun_out/microbench_c4.json 2> gpurun_out/microbench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/microbench_c4.json'))
for r in d['sweep']:
    print(r['W'], round(r['ms'], 4), 'exec', round(r['frac_tensor_peak_executed'], 3))
PY
timeout 300 python bench.py --forward-only --steps 10 --warmup 3 > gpurun_out/bench_forward_only.json 2> gpurun_out/bench_forward_only.err; echo "fwd rc=$?"
timeout 300 python bench.py --impl reference --ref-device cuda --forward-only --steps 5 --warmup 2 > gpurun_out/bench_reference_gpu_eager_forward_only.json 2>> gpurun_out/bench_forward_only.err; echo "ref fwd rc=$?"
cut -c1-330 gpurun_out/bench_forward_only.json gpurun_out/bench_reference_gpu_eager_forward_only.json
